/* kaldi_oracle.c -- CPU oracle (see kaldi_oracle.h for scope and provenance).
 *
 * TEST INFRASTRUCTURE ONLY: never imported by the product path.
 *
 * Arithmetic conventions: Kaldi's BaseFloat is float32.  Where Kaldi runs an
 * explicit scalar float loop we do the same in float (compiled with
 * -ffp-contract=off so no FMA is introduced).  Where Kaldi calls BLAS
 * (VecVec / AddMatVec: summation order unspecified) or its float32
 * split-radix FFT, the oracle accumulates in double and rounds once to float:
 * that is at least as close to the exact value as any float32 evaluation, so
 * both Kaldi and the CUDA path sit within float32 round-off of it.
 */
#include "kaldi_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.1415926535897932384626433832795
#endif
#define M_2PI 6.283185307179586476925286766559005
#ifndef M_SQRT2
#define M_SQRT2 1.4142135623730950488016887
#endif

/* ------------------------------------------------------------------------ */
/* framing: kaldi feature-window.cc, reached from shennong/frames.py:137,
 * shennong/processor/base.py:130                                            */

int32_t orc_window_shift(const orc_frame_opts *o) {
  /* FrameExtractionOptions::WindowShift: float * double * float -> int32 */
  return (int32_t)((double)o->samp_freq * 0.001 * (double)o->frame_shift_ms);
}

int32_t orc_window_size(const orc_frame_opts *o) {
  return (int32_t)((double)o->samp_freq * 0.001 * (double)o->frame_length_ms);
}

static int32_t round_up_pow2(int32_t n) {
  int32_t p = 1;
  while (p < n) p <<= 1;
  return p;
}

int32_t orc_padded_window_size(const orc_frame_opts *o) {
  int32_t w = orc_window_size(o);
  return o->round_to_power_of_two ? round_up_pow2(w) : w;
}

int64_t orc_num_frames(int64_t nsamples, const orc_frame_opts *o) {
  /* NumFrames(..., flush=true) */
  int64_t shift = orc_window_shift(o), length = orc_window_size(o);
  if (shift <= 0) return -1;
  if (o->snip_edges) {
    if (nsamples < length) return 0;
    return 1 + (nsamples - length) / shift;
  }
  return (nsamples + shift / 2) / shift;
}

int64_t orc_first_sample_of_frame(int32_t frame, const orc_frame_opts *o) {
  int64_t shift = orc_window_shift(o);
  if (o->snip_edges) return (int64_t)frame * shift;
  int64_t midpoint = shift * frame + shift / 2;
  return midpoint - orc_window_size(o) / 2;
}

void orc_window_function(const orc_frame_opts *o, float *out) {
  /* FeatureWindowFunction: double math, float storage
   * (formulas also documented at shennong/window.py:9-38) */
  int32_t n = orc_window_size(o);
  double a = M_2PI / (n - 1);
  for (int32_t i = 0; i < n; i++) {
    double x = (double)i, w;
    switch (o->window_type) {
      case ORC_WIN_HANNING: w = 0.5 - 0.5 * cos(a * x); break;
      case ORC_WIN_HAMMING: w = 0.54 - 0.46 * cos(a * x); break;
      case ORC_WIN_POVEY: w = pow(0.5 - 0.5 * cos(a * x), 0.85); break;
      case ORC_WIN_RECTANGULAR: w = 1.0; break;
      default:
        w = o->blackman_coeff - 0.5 * cos(a * x) +
            (0.5 - o->blackman_coeff) * cos(2 * a * x);
    }
    out[i] = (float)w;
  }
}

/* ExtractWindow + ProcessWindow (restated in shennong/processor/plp.py:171-260).
 * window: [padded] output.  Returns log-energy (pre-window) if requested.
 * eps is FLT_EPSILON for Kaldi's C++ computers and DBL_EPSILON for the PLP
 * python restatement (plp.py:191-193 uses np.finfo(float).eps).             */
static float extract_window(const float *wave, int64_t nsamples, int32_t frame,
                            const orc_frame_opts *o, const float *window_fn,
                            float *window, int need_log_energy, double eps) {
  int32_t len = orc_window_size(o), padded = orc_padded_window_size(o);
  int64_t start = orc_first_sample_of_frame(frame, o);
  if (start >= 0 && start + len <= nsamples) {
    for (int32_t s = 0; s < len; s++) window[s] = wave[start + s];
  } else {
    for (int32_t s = 0; s < len; s++) {
      int64_t k = start + s;
      while (k < 0 || k >= nsamples) {
        if (k < 0) k = -k - 1;
        else k = 2 * nsamples - 1 - k;
      }
      window[s] = wave[k];
    }
  }
  for (int32_t s = len; s < padded; s++) window[s] = 0.0f;
  /* dither is stochastic in the reference (RandGauss on libc rand()): parity
   * is only defined for dither == 0.  For the CPU-baseline timing the oracle
   * draws Box-Muller normals with the same arithmetic cost as Kaldi's
   * RandGauss (one log, sqrt and cos per sample) from a counter-based hash. */
  if (o->dither != 0.0f) {
    for (int32_t s = 0; s < len; s++) {
      uint64_t z = ((uint64_t)frame * 8192u + (uint64_t)s + 1u) * 0x9e3779b97f4a7c15ull;
      z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
      z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
      z ^= z >> 31;
      float u1 = ((float)((uint32_t)z >> 8) + 1.0f) * (1.0f / 16777216.0f);
      float u2 = (float)((uint32_t)(z >> 40)) * (1.0f / 16777216.0f);
      window[s] += o->dither * (sqrtf(-2.0f * logf(u1)) * cosf(6.2831853f * u2));
    }
  }
  if (o->remove_dc_offset) {
    double sum = 0.0; /* VectorBase<float>::Sum() accumulates in double */
    for (int32_t s = 0; s < len; s++) sum += window[s];
    float mean = (float)(sum / len);
    for (int32_t s = 0; s < len; s++) window[s] -= mean;
  }
  float log_energy = 0.0f;
  if (need_log_energy) {
    double e = 0.0;
    for (int32_t s = 0; s < len; s++) e += (double)window[s] * window[s];
    float ef = (float)e;
    double floored = (double)ef > eps ? (double)ef : eps;
    log_energy = (float)log(floored);
  }
  if (o->preemph_coeff != 0.0f) {
    float c = o->preemph_coeff;
    for (int32_t s = len - 1; s > 0; s--) window[s] -= c * window[s - 1];
    window[0] -= c * window[0];
  }
  for (int32_t s = 0; s < len; s++) window[s] *= window_fn[s];
  return log_energy;
}

/* ------------------------------------------------------------------------ */
/* real FFT -> power spectrum (srfft.cc / ComputePowerSpectrum; plp.py:571-576).
 * FFT evaluated in double; re/im rounded to float; power formed in float like
 * ComputePowerSpectrum does.                                                */

/* twiddle/bit-reversal tables of a plan (double precision) */
typedef struct {
  int32_t n;          /* real FFT size */
  int32_t pow2;
  double *wr, *wi;    /* e^{-2 pi i k / n}, k < n   */
  int32_t *rev;       /* bit reversal for the n/2-point complex FFT */
} orc_fft;

static void fft_init(orc_fft *f, int32_t n) {
  f->n = n;
  f->pow2 = (n & (n - 1)) == 0 && n >= 4;
  f->wr = (double *)malloc(sizeof(double) * n);
  f->wi = (double *)malloc(sizeof(double) * n);
  for (int32_t k = 0; k < n; k++) {
    f->wr[k] = cos(-M_2PI * k / n);
    f->wi[k] = sin(-M_2PI * k / n);
  }
  f->rev = NULL;
  if (f->pow2) {
    int32_t m = n / 2;
    f->rev = (int32_t *)malloc(sizeof(int32_t) * m);
    for (int32_t i = 0, j = 0; i < m; i++) {
      f->rev[i] = j;
      int32_t bit = m >> 1;
      for (; bit && (j & bit); bit >>= 1) j ^= bit;
      j ^= bit;
    }
  }
}
static void fft_free(orc_fft *f) { free(f->wr); free(f->wi); free(f->rev); }

/* frame: [n] in, power: [n/2+1] out; zr/zi: scratch [n] */
static void power_spectrum(const orc_fft *f, const float *frame, float *power,
                           double *zr, double *zi) {
  const int32_t n = f->n, half = n / 2;
  double *xr = zr + half, *xi = zi + half; /* second halves hold X */
  if (f->pow2) {
    /* packed real FFT: z[j] = x[2j] + i x[2j+1], m-point complex FFT */
    const int32_t m = half;
    for (int32_t j = 0; j < m; j++) {
      zr[f->rev[j]] = frame[2 * j];
      zi[f->rev[j]] = frame[2 * j + 1];
    }
    for (int32_t len = 2; len <= m; len <<= 1) {
      const int32_t h = len >> 1, step = n / len; /* e^{-2 pi i k/len} = w[k*step] */
      for (int32_t i = 0; i < m; i += len)
        for (int32_t k = 0; k < h; k++) {
          const double wr = f->wr[k * step], wi = f->wi[k * step];
          const double vr = zr[i + k + h] * wr - zi[i + k + h] * wi;
          const double vi = zr[i + k + h] * wi + zi[i + k + h] * wr;
          const double ur = zr[i + k], ui = zi[i + k];
          zr[i + k] = ur + vr; zi[i + k] = ui + vi;
          zr[i + k + h] = ur - vr; zi[i + k + h] = ui - vi;
        }
    }
    for (int32_t k = 0; k <= m; k++) {
      const int32_t a = k % m, b = (m - k) % m;
      const double er = 0.5 * (zr[a] + zr[b]), ei = 0.5 * (zi[a] - zi[b]);
      const double orr = 0.5 * (zi[a] + zi[b]), oi = -0.5 * (zr[a] - zr[b]);
      const double wr = (k == m) ? -1.0 : f->wr[k], wi = (k == m) ? 0.0 : f->wi[k];
      xr[k] = er + orr * wr - oi * wi;
      xi[k] = ei + orr * wi + oi * wr;
    }
  } else {
    for (int32_t k = 0; k <= half; k++) {
      double sr = 0.0, si = 0.0;
      int32_t idx = 0;
      for (int32_t t = 0; t < n; t++) {
        sr += frame[t] * f->wr[idx];
        si += frame[t] * f->wi[idx];
        idx += k; if (idx >= n) idx -= n;
      }
      xr[k] = sr; xi[k] = si;
    }
  }
  for (int32_t k = 0; k <= half; k++) {
    float r = (float)xr[k], i = (float)xi[k];
    if (k == 0 || 2 * k == n) power[k] = r * r; /* purely real bins */
    else power[k] = r * r + i * i;
  }
}

/* ------------------------------------------------------------------------ */
/* mel banks: mel-computations.cc MelBanks::MelBanks (shennong call sites:
 * processor/base.py:308, plp.py:491-494)                                    */

static float mel_scale(float f) { return 1127.0f * logf(1.0f + f / 700.0f); }
static float inv_mel_scale(float m) {
  return 700.0f * (expf(m / 1127.0f) - 1.0f);
}

static float vtln_warp_freq(float vtln_low_cutoff, float vtln_high_cutoff,
                            float low_freq, float high_freq, float warp,
                            float freq) {
  if (freq < low_freq || freq > high_freq) return freq;
  float one = 1.0f;
  float l = vtln_low_cutoff * (warp > one ? warp : one);
  float h = vtln_high_cutoff * (warp < one ? warp : one);
  float scale = 1.0f / warp;
  float Fl = scale * l, Fh = scale * h;
  float scale_left = (Fl - low_freq) / (l - low_freq);
  float scale_right = (high_freq - Fh) / (high_freq - h);
  if (freq < l) return low_freq + scale_left * (freq - low_freq);
  else if (freq < h) return scale * freq;
  else return high_freq + scale_right * (freq - high_freq);
}

static float vtln_warp_mel_freq(float vl, float vh, float lo, float hi,
                                float warp, float mel) {
  return mel_scale(vtln_warp_freq(vl, vh, lo, hi, warp, inv_mel_scale(mel)));
}

int32_t orc_mel_banks(const orc_frame_opts *fo, const orc_mel_opts *mo,
                      float vtln_warp, float *weights, float *center_freqs) {
  int32_t num_bins = mo->num_bins;
  if (num_bins < 3) return -1;
  float sample_freq = fo->samp_freq;
  int32_t padded = orc_padded_window_size(fo);
  if (padded % 2 != 0) return -1;
  int32_t num_fft_bins = padded / 2;
  float nyquist = 0.5f * sample_freq;
  float low_freq = mo->low_freq, high_freq;
  if (mo->high_freq > 0.0f) high_freq = mo->high_freq;
  else high_freq = nyquist + mo->high_freq;
  if (low_freq < 0.0f || low_freq >= nyquist || high_freq <= 0.0f ||
      high_freq > nyquist || high_freq <= low_freq)
    return -1;
  float fft_bin_width = sample_freq / padded;
  float mel_low = mel_scale(low_freq), mel_high = mel_scale(high_freq);
  float mel_delta = (mel_high - mel_low) / (num_bins + 1);
  float vtln_low = mo->vtln_low, vtln_high = mo->vtln_high;
  if (vtln_high < 0.0f) vtln_high += nyquist;
  if (vtln_warp != 1.0f &&
      (vtln_low < 0.0f || vtln_low <= low_freq || vtln_low >= high_freq ||
       vtln_high <= 0.0f || vtln_high >= high_freq || vtln_high <= vtln_low))
    return -1;
  memset(weights, 0, sizeof(float) * (size_t)num_bins * num_fft_bins);
  for (int32_t bin = 0; bin < num_bins; bin++) {
    float left = mel_low + bin * mel_delta,
          center = mel_low + (bin + 1) * mel_delta,
          right = mel_low + (bin + 2) * mel_delta;
    if (vtln_warp != 1.0f) {
      left = vtln_warp_mel_freq(vtln_low, vtln_high, low_freq, high_freq,
                                vtln_warp, left);
      center = vtln_warp_mel_freq(vtln_low, vtln_high, low_freq, high_freq,
                                  vtln_warp, center);
      right = vtln_warp_mel_freq(vtln_low, vtln_high, low_freq, high_freq,
                                 vtln_warp, right);
    }
    if (center_freqs) center_freqs[bin] = inv_mel_scale(center);
    int32_t first = -1;
    for (int32_t i = 0; i < num_fft_bins; i++) {
      float freq = fft_bin_width * i;
      float mel = mel_scale(freq);
      if (mel > left && mel < right) {
        float w;
        if (mel <= center) w = (mel - left) / (center - left);
        else w = (right - mel) / (right - center);
        weights[(size_t)bin * num_fft_bins + i] = w;
        if (first == -1) first = i;
      }
    }
    if (first == -1) return -1; /* KALDI_ASSERT: num-mel-bins too large */
  }
  return 0;
}

/* ------------------------------------------------------------------------ */
/* per-kind setup                                                           */

typedef struct {
  orc_frame_opts fo;
  orc_mel_opts mo;
  orc_feat_opts xo;
  int32_t len, padded, nfft_bins, dim;
  float *window_fn;
  float *mel_w;      /* dense [num_bins, nfft_bins] */
  int32_t *mel_first, *mel_size;
  float *center_freqs;
  float *dct;        /* [num_ceps, num_bins] */
  float *lifter;     /* [num_ceps] */
  float *loudness;   /* [num_bins] */
  float *idft;       /* [lpc_order+1, num_bins+2] */
  float log_energy_floor;
  double eps;
  orc_fft fft;
} orc_plan;

static void plan_free(orc_plan *p) {
  free(p->window_fn); free(p->mel_w); free(p->mel_first); free(p->mel_size);
  free(p->center_freqs); free(p->dct); free(p->lifter); free(p->loudness);
  free(p->idft);
  if (p->fft.wr) fft_free(&p->fft);
  memset(p, 0, sizeof(*p));
}

int32_t orc_feat_dim(const orc_frame_opts *fo, const orc_mel_opts *mo,
                     const orc_feat_opts *xo) {
  switch (xo->kind) {
    case ORC_FEAT_SPECTROGRAM: return orc_padded_window_size(fo) / 2 + 1;
    case ORC_FEAT_FBANK: return mo->num_bins + (xo->use_energy ? 1 : 0);
    case ORC_FEAT_MFCC:
      if (xo->num_ceps <= 0 || xo->num_ceps > mo->num_bins) return -1;
      return xo->num_ceps;
    case ORC_FEAT_PLP:
      if (xo->num_ceps <= 0 || xo->num_ceps > xo->lpc_order + 1) return -1;
      return xo->num_ceps;
    case ORC_FEAT_ENERGY: return 1;
  }
  return -1;
}

static int32_t plan_init(orc_plan *p, const orc_frame_opts *fo,
                         const orc_mel_opts *mo, const orc_feat_opts *xo,
                         float vtln_warp) {
  memset(p, 0, sizeof(*p));
  p->fo = *fo; p->xo = *xo;
  if (mo) p->mo = *mo;
  if (xo->kind == ORC_FEAT_ENERGY && xo->raw_energy) {
    /* shennong/processor/energy.py:148-151 */
    p->fo.preemph_coeff = 0.0f;
    p->fo.window_type = ORC_WIN_RECTANGULAR;
  }
  p->len = orc_window_size(&p->fo);
  p->padded = orc_padded_window_size(&p->fo);
  p->nfft_bins = p->padded / 2;
  if (p->len <= 0 || orc_window_shift(&p->fo) <= 0) return -1;
  p->dim = orc_feat_dim(&p->fo, &p->mo, xo);
  if (p->dim <= 0) return -1;
  p->eps = (xo->kind == ORC_FEAT_PLP) ? DBL_EPSILON : (double)FLT_EPSILON;
  if (xo->kind != ORC_FEAT_ENERGY) fft_init(&p->fft, p->padded);
  p->window_fn = (float *)malloc(sizeof(float) * p->len);
  orc_window_function(&p->fo, p->window_fn);
  if (xo->energy_floor > 0.0f)
    p->log_energy_floor = (float)log((double)xo->energy_floor);
  int needs_mel = xo->kind == ORC_FEAT_FBANK || xo->kind == ORC_FEAT_MFCC ||
                  xo->kind == ORC_FEAT_PLP;
  if (needs_mel) {
    int32_t nb = mo->num_bins;
    if (nb < 3) return -1;
    p->mel_w = (float *)malloc(sizeof(float) * (size_t)nb * p->nfft_bins);
    p->center_freqs = (float *)malloc(sizeof(float) * nb);
    if (orc_mel_banks(&p->fo, mo, vtln_warp, p->mel_w, p->center_freqs) != 0)
      return -1;
    p->mel_first = (int32_t *)malloc(sizeof(int32_t) * nb);
    p->mel_size = (int32_t *)malloc(sizeof(int32_t) * nb);
    for (int32_t b = 0; b < nb; b++) {
      int32_t first = -1, last = -1;
      /* bins_[bin] keeps the range [first_index, last_index] where the
       * membership test held, including weights that evaluate to 0 */
      for (int32_t i = 0; i < p->nfft_bins; i++) {
        if (p->mel_w[(size_t)b * p->nfft_bins + i] != 0.0f) {
          if (first < 0) first = i;
          last = i;
        }
      }
      if (first < 0) { first = 0; last = -1; }
      p->mel_first[b] = first; p->mel_size[b] = last + 1 - first;
    }
  }
  if (xo->kind == ORC_FEAT_MFCC) {
    int32_t nb = mo->num_bins, nc = xo->num_ceps;
    p->dct = (float *)malloc(sizeof(float) * (size_t)nc * nb);
    /* ComputeDctMatrix */
    float norm0 = (float)sqrt(1.0 / (double)(float)nb);
    float norm = (float)sqrt(2.0 / (double)(float)nb);
    for (int32_t n = 0; n < nb; n++) p->dct[n] = norm0;
    for (int32_t k = 1; k < nc; k++)
      for (int32_t n = 0; n < nb; n++)
        p->dct[(size_t)k * nb + n] =
            (float)((double)norm * cos(M_PI / nb * (n + 0.5) * k));
  }
  if ((xo->kind == ORC_FEAT_MFCC || xo->kind == ORC_FEAT_PLP) &&
      xo->cepstral_lifter != 0.0f) {
    int32_t nc = xo->num_ceps;
    p->lifter = (float *)malloc(sizeof(float) * nc);
    /* ComputeLifterCoeffs: 1.0 + 0.5 * Q * sin(M_PI * i / Q) */
    for (int32_t i = 0; i < nc; i++)
      p->lifter[i] = (float)(1.0 + 0.5 * (double)xo->cepstral_lifter *
                                       sin(M_PI * i / (double)xo->cepstral_lifter));
  }
  if (xo->kind == ORC_FEAT_PLP) {
    int32_t nb = mo->num_bins, nl = xo->lpc_order + 1, d = nb + 2;
    p->loudness = (float *)malloc(sizeof(float) * nb);
    /* GetEqualLoudnessVector */
    for (int32_t i = 0; i < nb; i++) {
      float fsq = p->center_freqs[i] * p->center_freqs[i];
      float fsub = fsq / (fsq + 1.6e5f);
      p->loudness[i] = fsub * fsub * ((fsq + 1.44e6f) / (fsq + 9.61e6f));
    }
    /* InitIdftBases(lpc_order + 1, num_bins + 2) */
    p->idft = (float *)malloc(sizeof(float) * (size_t)nl * d);
    float angle = (float)(M_PI / (double)(float)(d - 1));
    float scale = (float)(1.0f / (2.0 * (double)(float)(d - 1)));
    for (int32_t i = 0; i < nl; i++) {
      p->idft[(size_t)i * d] = (float)(1.0 * scale);
      float i_fl = (float)i;
      for (int32_t j = 1; j < d - 1; j++) {
        float j_fl = (float)j;
        p->idft[(size_t)i * d + j] =
            (float)(2.0 * scale * cos((double)(angle * i_fl * j_fl)));
      }
      p->idft[(size_t)i * d + d - 1] =
          (float)(scale * cos((double)(angle * i_fl * (float)(d - 1))));
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------ */
/* RASTA filter state (shennong/processor/plp.py:64-146): float64 IIR per mel
 * bin, inputs log(x + float32 eps) in float32                               */
typedef struct {
  int32_t size, count;
  double *delay;        /* [4, size] */
  float *first_frames;  /* [4, size] */
} rasta_state;

static const double RASTA_NUM[5] = {0.2, 0.1, -0.0, -0.1, -0.2};

static void rasta_init(rasta_state *r, int32_t size) {
  r->size = size; r->count = 0;
  r->delay = (double *)calloc((size_t)4 * size, sizeof(double));
  r->first_frames = (float *)calloc((size_t)4 * size, sizeof(float));
}
static void rasta_free(rasta_state *r) { free(r->delay); free(r->first_frames); }

static void rasta_filter(rasta_state *r, float *mel /* in/out [size] */) {
  int32_t n = r->size;
  float *x = (float *)malloc(sizeof(float) * n);
  for (int32_t b = 0; b < n; b++) x[b] = logf(mel[b] + FLT_EPSILON);
  if (r->count < 4) {
    memcpy(r->first_frames + (size_t)r->count * n, x, sizeof(float) * n);
    if (r->count == 3) {
      /* lfilter(num, 1, first_frames, zi=lfilter_zi(num, 1) * x[0]) */
      double zi[4];
      zi[3] = RASTA_NUM[4];
      zi[2] = RASTA_NUM[3] + zi[3];
      zi[1] = RASTA_NUM[2] + zi[2];
      zi[0] = RASTA_NUM[1] + zi[1];
      for (int32_t b = 0; b < n; b++) {
        double z[4];
        double x0 = r->first_frames[b];
        for (int k = 0; k < 4; k++) z[k] = zi[k] * x0;
        for (int t = 0; t < 4; t++) {
          double xt = r->first_frames[(size_t)t * n + b];
          z[0] = RASTA_NUM[1] * xt + z[1];
          z[1] = RASTA_NUM[2] * xt + z[2];
          z[2] = RASTA_NUM[3] * xt + z[3];
          z[3] = RASTA_NUM[4] * xt;
        }
        for (int k = 0; k < 4; k++) r->delay[(size_t)k * n + b] = z[k];
      }
    }
    for (int32_t b = 0; b < n; b++) mel[b] = expf(0.0f);
  } else {
    for (int32_t b = 0; b < n; b++) {
      double *z0 = &r->delay[b], *z1 = &r->delay[(size_t)n + b],
             *z2 = &r->delay[(size_t)2 * n + b], *z3 = &r->delay[(size_t)3 * n + b];
      double xt = x[b];
      double y = RASTA_NUM[0] * xt + *z0;
      *z0 = RASTA_NUM[1] * xt + *z1 + 0.94 * y;
      *z1 = RASTA_NUM[2] * xt + *z2;
      *z2 = RASTA_NUM[3] * xt + *z3;
      *z3 = RASTA_NUM[4] * xt;
      mel[b] = expf((float)y);
    }
  }
  r->count++;
  free(x);
}

/* ------------------------------------------------------------------------ */
/* per-frame computers                                                       */

static float dotf(const float *a, const float *b, int32_t n) {
  double s = 0.0;
  for (int32_t i = 0; i < n; i++) s += (double)a[i] * b[i];
  return (float)s;
}

static float log_energy_of(const float *frame, int32_t n, double eps) {
  float e = dotf(frame, frame, n);
  double f = (double)e > eps ? (double)e : eps;
  return (float)log(f);
}

static void mel_compute(const orc_plan *p, const float *power, float *mel) {
  for (int32_t b = 0; b < p->mo.num_bins; b++)
    mel[b] = dotf(p->mel_w + (size_t)b * p->nfft_bins + p->mel_first[b],
                  power + p->mel_first[b], p->mel_size[b]);
}

/* Durbin + ComputeLpc (mel-computations.cc; used at plp.py:601) */
static float compute_lpc(const float *ac, int32_t n, float *lpc, float *tmp) {
  float E = ac[0];
  for (int32_t i = 0; i < n; i++) {
    float ki = ac[i + 1];
    for (int32_t j = 0; j < i; j++) ki += lpc[j] * ac[i - j];
    ki = ki / E;
    float c = 1 - ki * ki;
    if (c < 1.0e-5f) c = 1.0e-5f;
    E *= c;
    tmp[i] = -ki;
    for (int32_t j = 0; j < i; j++) tmp[j] = lpc[j] - ki * lpc[i - j - 1];
    for (int32_t j = 0; j <= i; j++) lpc[j] = tmp[j];
  }
  return (float)-log(1.0 / (double)E);
}

static void compute_frame(const orc_plan *p, float raw_log_energy,
                          float *frame /*[padded], destroyed*/, float *out,
                          double *wr, double *wi, float *power, float *mel,
                          rasta_state *rasta) {
  const orc_feat_opts *xo = &p->xo;
  float log_energy = raw_log_energy;
  int post_window_energy = 0;
  switch (xo->kind) {
    case ORC_FEAT_SPECTROGRAM: post_window_energy = !xo->raw_energy; break;
    default: post_window_energy = xo->use_energy && !xo->raw_energy;
  }
  if (post_window_energy) log_energy = log_energy_of(frame, p->padded, p->eps);
  power_spectrum(&p->fft, frame, power, wr, wi);
  int32_t nb = p->mo.num_bins;
  if (xo->kind == ORC_FEAT_SPECTROGRAM) {
    for (int32_t k = 0; k <= p->nfft_bins; k++) {
      float v = power[k] < FLT_EPSILON ? FLT_EPSILON : power[k];
      out[k] = logf(v);
    }
    if (xo->energy_floor > 0.0f && log_energy < p->log_energy_floor)
      log_energy = p->log_energy_floor;
    out[0] = log_energy;
    return;
  }
  if (xo->kind == ORC_FEAT_FBANK) {
    if (!xo->use_power)
      for (int32_t k = 0; k <= p->nfft_bins; k++) power[k] = powf(power[k], 0.5f);
    int32_t off = (xo->use_energy && !xo->htk_compat) ? 1 : 0;
    mel_compute(p, power, out + off);
    if (xo->use_log_fbank)
      for (int32_t b = 0; b < nb; b++) {
        float v = out[off + b] < FLT_EPSILON ? FLT_EPSILON : out[off + b];
        out[off + b] = logf(v);
      }
    if (xo->use_energy) {
      if (xo->energy_floor > 0.0f && log_energy < p->log_energy_floor)
        log_energy = p->log_energy_floor;
      out[xo->htk_compat ? nb : 0] = log_energy;
    }
    return;
  }
  if (xo->kind == ORC_FEAT_MFCC) {
    mel_compute(p, power, mel);
    for (int32_t b = 0; b < nb; b++) {
      float v = mel[b] < FLT_EPSILON ? FLT_EPSILON : mel[b];
      mel[b] = logf(v);
    }
    int32_t nc = xo->num_ceps;
    for (int32_t k = 0; k < nc; k++) out[k] = dotf(p->dct + (size_t)k * nb, mel, nb);
    if (xo->cepstral_lifter != 0.0f)
      for (int32_t k = 0; k < nc; k++) out[k] *= p->lifter[k];
    if (xo->use_energy) {
      if (xo->energy_floor > 0.0f && log_energy < p->log_energy_floor)
        log_energy = p->log_energy_floor;
      out[0] = log_energy;
    }
    if (xo->htk_compat) {
      float energy = out[0];
      for (int32_t i = 0; i < nc - 1; i++) out[i] = out[i + 1];
      if (!xo->use_energy) energy *= (float)M_SQRT2;
      out[nc - 1] = energy;
    }
    return;
  }
  /* PLP: shennong/processor/plp.py:548-626 */
  {
    int32_t nl = xo->lpc_order, nc = xo->num_ceps, d = nb + 2;
    float *dup = mel; /* [nb+2] */
    mel_compute(p, power, dup + 1);
    if (xo->rasta && rasta) rasta_filter(rasta, dup + 1);
    for (int32_t b = 0; b < nb; b++) dup[1 + b] *= p->loudness[b];
    for (int32_t b = 0; b < nb; b++) dup[1 + b] = powf(dup[1 + b], xo->compress_factor);
    dup[0] = dup[1];
    dup[nb + 1] = dup[nb];
    float ac[64] = {0}, lpc[64] = {0}, tmp[64] = {0}, cep[64] = {0};
    for (int32_t i = 0; i <= nl; i++) ac[i] = dotf(p->idft + (size_t)i * d, dup, d);
    float residual = compute_lpc(ac, nl, lpc, tmp);
    /* plp.py:603 max(residual_log_energy, np.finfo(float).eps) */
    double res = (double)residual > DBL_EPSILON ? (double)residual : DBL_EPSILON;
    /* _lpc2cepstrum plp.py:164-168: python double accumulation, float store */
    for (int32_t i = 0; i < nl; i++) {
      double sum = 0.0;
      for (int32_t j = 0; j < i; j++)
        sum += (double)(i - j) * (double)lpc[j] * (double)cep[i - j - 1];
      cep[i] = (float)(-(double)lpc[i] - sum / (double)(i + 1));
    }
    for (int32_t i = 1; i < nc; i++) out[i] = cep[i - 1];
    out[0] = (float)res;
    if (xo->cepstral_lifter != 0.0f)
      for (int32_t i = 0; i < nc; i++) out[i] *= p->lifter[i];
    if (xo->cepstral_scale != 1.0f)
      for (int32_t i = 0; i < nc; i++) out[i] *= xo->cepstral_scale;
    if (xo->use_energy) {
      if (xo->energy_floor > 0.0f && log_energy < p->log_energy_floor)
        log_energy = p->log_energy_floor;
      out[0] = log_energy;
    }
    if (xo->htk_compat) {
      float e = out[0];
      for (int32_t i = 0; i < nc - 1; i++) out[i] = out[i + 1];
      out[nc - 1] = e;
    }
  }
}

static int64_t compute_with_plan(const orc_plan *p, const float *wave,
                                 int64_t nsamples, float *out, double *out64) {
  int64_t nframes = orc_num_frames(nsamples, &p->fo);
  if (nframes <= 0) return nframes < 0 ? -1 : 0;
  const orc_feat_opts *xo = &p->xo;
  float *frame = (float *)malloc(sizeof(float) * p->padded);
  double *wr = (double *)malloc(sizeof(double) * (p->padded + 2));
  double *wi = (double *)malloc(sizeof(double) * (p->padded + 2));
  float *power = (float *)malloc(sizeof(float) * (p->nfft_bins + 1));
  float *mel = (float *)malloc(sizeof(float) * (p->mo.num_bins + 2 + 1));
  rasta_state rasta; int have_rasta = 0;
  if (xo->kind == ORC_FEAT_PLP && xo->rasta) {
    rasta_init(&rasta, p->mo.num_bins); have_rasta = 1;
  }
  int need_raw;
  switch (xo->kind) {
    case ORC_FEAT_SPECTROGRAM: need_raw = xo->raw_energy; break;
    case ORC_FEAT_ENERGY: need_raw = 0; break;
    default: need_raw = xo->use_energy && xo->raw_energy;
  }
  for (int64_t f = 0; f < nframes; f++) {
    float raw = extract_window(wave, nsamples, (int32_t)f, &p->fo, p->window_fn,
                               frame, need_raw, p->eps);
    if (xo->kind == ORC_FEAT_ENERGY) {
      /* shennong/processor/energy.py:171-183: float64 sum of squares */
      double e = 0.0;
      for (int32_t s = 0; s < p->len; s++) e += (double)frame[s] * (double)frame[s];
      if (e < DBL_MIN) e = DBL_MIN;
      if (xo->energy_compression == 1) e = log(e);
      else if (xo->energy_compression == 2) e = sqrt(e);
      if (out64) out64[f] = e;
      if (out) out[f] = (float)e;
      continue;
    }
    compute_frame(p, raw, frame, out + (size_t)f * p->dim, wr, wi, power, mel,
                  have_rasta ? &rasta : NULL);
  }
  if (have_rasta) rasta_free(&rasta);
  free(frame); free(wr); free(wi); free(power); free(mel);
  return nframes;
}

int64_t orc_compute_features(const float *wave, int64_t nsamples,
                             const orc_frame_opts *fo, const orc_mel_opts *mo,
                             const orc_feat_opts *xo, float vtln_warp,
                             float *out, double *out64) {
  orc_plan p;
  if (plan_init(&p, fo, mo, xo, vtln_warp) != 0) { plan_free(&p); return -1; }
  int64_t n = compute_with_plan(&p, wave, nsamples, out, out64);
  plan_free(&p);
  return n;
}

int64_t orc_compute_features_batch(const int16_t *pcm,
                                   const int64_t *sample_offsets,
                                   const int64_t *frame_offsets, int64_t nutts,
                                   const orc_frame_opts *fo,
                                   const orc_mel_opts *mo,
                                   const orc_feat_opts *xo, float *out,
                                   int32_t nthreads) {
  orc_plan p;
  if (plan_init(&p, fo, mo, xo, 1.0f) != 0) { plan_free(&p); return -1; }
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  int64_t total = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
  for (int64_t u = 0; u < nutts; u++) {
    int64_t n = sample_offsets[u + 1] - sample_offsets[u];
    float *wave = (float *)malloc(sizeof(float) * (n > 0 ? n : 1));
    for (int64_t i = 0; i < n; i++) wave[i] = (float)pcm[sample_offsets[u] + i];
    total += compute_with_plan(&p, wave, n, out + (size_t)frame_offsets[u] * p.dim, NULL);
    free(wave);
  }
  plan_free(&p);
  return total;
}

/* ------------------------------------------------------------------------ */
/* deltas: feature-functions.cc DeltaFeatures (shennong/postprocessor/delta.py:130) */

void orc_compute_deltas(const float *in, int64_t nframes, int32_t dim,
                        int32_t order, int32_t window, float *out) {
  float **scales = (float **)malloc(sizeof(float *) * (order + 1));
  int32_t *sdim = (int32_t *)malloc(sizeof(int32_t) * (order + 1));
  scales[0] = (float *)malloc(sizeof(float));
  scales[0][0] = 1.0f; sdim[0] = 1;
  for (int32_t i = 1; i <= order; i++) {
    int32_t prev_offset = (sdim[i - 1] - 1) / 2, cur_offset = prev_offset + window;
    sdim[i] = sdim[i - 1] + 2 * window;
    scales[i] = (float *)calloc(sdim[i], sizeof(float));
    float normalizer = 0.0f;
    for (int32_t j = -window; j <= window; j++) {
      normalizer += (float)(j * j);
      for (int32_t k = -prev_offset; k <= prev_offset; k++)
        scales[i][j + k + cur_offset] += (float)j * scales[i - 1][k + prev_offset];
    }
    float inv = (float)(1.0 / (double)normalizer);
    for (int32_t k = 0; k < sdim[i]; k++) scales[i][k] *= inv;
  }
  int32_t odim = dim * (order + 1);
  for (int64_t t = 0; t < nframes; t++) {
    float *o = out + (size_t)t * odim;
    for (int32_t k = 0; k < odim; k++) o[k] = 0.0f;
    for (int32_t i = 0; i <= order; i++) {
      int32_t max_offset = (sdim[i] - 1) / 2;
      for (int32_t j = -max_offset; j <= max_offset; j++) {
        int64_t tt = t + j;
        if (tt < 0) tt = 0;
        else if (tt >= nframes) tt = nframes - 1;
        float s = scales[i][j + max_offset];
        if (s != 0.0f)
          for (int32_t k = 0; k < dim; k++)
            o[i * dim + k] += s * in[(size_t)tt * dim + k];
      }
    }
  }
  for (int32_t i = 0; i <= order; i++) free(scales[i]);
  free(scales); free(sdim);
}

/* ------------------------------------------------------------------------ */
/* CMVN: transform/cmvn.cc (shennong/postprocessor/cmvn.py:217-219, 273-278) */

void orc_cmvn_accumulate(const float *feats, int64_t nframes, int32_t dim,
                         const float *weights, double *stats) {
  double *mean = stats, *var = stats + (dim + 1);
  for (int64_t t = 0; t < nframes; t++) {
    float w = weights ? weights[t] : 1.0f;
    if (weights && w == 0.0f) continue;
    mean[dim] += w;
    for (int32_t d = 0; d < dim; d++) {
      float x = feats[(size_t)t * dim + d];
      float xw = x * w;        /* float product, then promoted */
      float xxw = x * x * w;
      mean[d] += xw;
      var[d] += xxw;
    }
  }
}

int32_t orc_cmvn_apply(const double *stats, int32_t dim, int32_t norm_vars,
                       int32_t reverse, float *feats, int64_t nframes) {
  double count = stats[dim];
  if (count < 1.0) return -1;
  const double *s0 = stats, *s1 = stats + (dim + 1);
  if (!reverse && !norm_vars) {
    for (int32_t d = 0; d < dim; d++) {
      float offset = (float)((double)(float)(-1.0 / count) * s0[d]);
      for (int64_t t = 0; t < nframes; t++) feats[(size_t)t * dim + d] += offset;
    }
    return 0;
  }
  for (int32_t d = 0; d < dim; d++) {
    double mean = s0[d] / count;
    double var = s1[d] / count - mean * mean;
    if (var < 1.0e-20) var = 1.0e-20;
    float scale, offset;
    if (!reverse) {
      double sc = 1.0 / sqrt(var);
      scale = (float)sc; offset = (float)(-(mean * sc));
    } else {
      scale = (float)sqrt(var); offset = (float)mean;
    }
    for (int64_t t = 0; t < nframes; t++) {
      float x = feats[(size_t)t * dim + d];
      if (!reverse || norm_vars) x = x * scale;
      x = x + offset;
      feats[(size_t)t * dim + d] = x;
    }
  }
  return 0;
}

/* SlidingWindowCmn: feature-functions.cc (shennong/postprocessor/cmvn.py:492) */
void orc_sliding_window_cmn(const float *in, int64_t nframes, int32_t dim,
                            int32_t center, int32_t cmn_window,
                            int32_t min_window, int32_t normalize_variance,
                            float *out) {
  double *cur_sum = (double *)calloc(dim, sizeof(double));
  double *cur_sumsq = (double *)calloc(dim, sizeof(double));
  int64_t last_start = -1, last_end = -1;
  for (int64_t t = 0; t < nframes; t++) {
    int64_t ws, we;
    if (center) { ws = t - cmn_window / 2; we = ws + cmn_window; }
    else { ws = t - cmn_window; we = t + 1; }
    if (ws < 0) { we -= ws; ws = 0; }
    if (!center) {
      if (we > t) we = (t + 1 > min_window) ? t + 1 : min_window;
    }
    if (we > nframes) {
      ws -= (we - nframes); we = nframes;
      if (ws < 0) ws = 0;
    }
    if (last_start == -1) {
      for (int64_t f = ws; f < we; f++)
        for (int32_t d = 0; d < dim; d++) {
          double x = in[(size_t)f * dim + d];
          cur_sum[d] += x; cur_sumsq[d] += x * x;
        }
    } else {
      if (ws > last_start)
        for (int32_t d = 0; d < dim; d++) {
          double x = in[(size_t)last_start * dim + d];
          cur_sum[d] -= x; cur_sumsq[d] -= x * x;
        }
      if (we > last_end)
        for (int32_t d = 0; d < dim; d++) {
          double x = in[(size_t)last_end * dim + d];
          cur_sum[d] += x; cur_sumsq[d] += x * x;
        }
    }
    int64_t wf = we - ws;
    last_start = ws; last_end = we;
    for (int32_t d = 0; d < dim; d++) {
      double y = (double)in[(size_t)t * dim + d] - cur_sum[d] / (double)wf;
      if (normalize_variance) {
        if (wf == 1) y = 0.0;
        else {
          double v = cur_sumsq[d] / (double)wf -
                     cur_sum[d] * cur_sum[d] / ((double)wf * (double)wf);
          if (v < 1.0e-10) v = 1.0e-10;
          y *= pow(v, -0.5);
        }
      }
      out[(size_t)t * dim + d] = (float)y;
    }
  }
  free(cur_sum); free(cur_sumsq);
}

/* ComputeVadEnergy: ivector/voice-activity-detection.cc (vad.py:183) */
void orc_vad_energy(const float *feats, int64_t nframes, int32_t dim,
                    float energy_threshold, float energy_mean_scale,
                    int32_t frames_context, float proportion_threshold,
                    float *out) {
  if (nframes == 0) return;
  float thr = energy_threshold;
  if (energy_mean_scale != 0.0f) {
    double s = 0.0;
    for (int64_t t = 0; t < nframes; t++) s += feats[(size_t)t * dim];
    float sum = (float)s;
    thr += energy_mean_scale * sum / (float)nframes;
  }
  for (int64_t t = 0; t < nframes; t++) {
    int32_t num = 0, den = 0;
    for (int64_t t2 = t - frames_context; t2 <= t + frames_context; t2++)
      if (t2 >= 0 && t2 < nframes) {
        den++;
        if (feats[(size_t)t2 * dim] > thr) num++;
      }
    out[t] = ((float)num >= (float)den * proportion_threshold) ? 1.0f : 0.0f;
  }
}

/* ------------------------------------------------------------------------ */
/* Kaldi pitch: pitch-functions.cc + resample.cc, reached through
 * compute_kaldi_pitch (shennong/processor/pitch_kaldi.py:298-299) in offline
 * mode: one AcceptWaveform (resampler not flushed) + InputFinished().        */

static int32_t gcd_i(int32_t a, int32_t b) { while (b) { int32_t t = a % b; a = b; b = t; } return a; }

/* LinearResample::FilterFunc (BaseFloat argument, double math inside) */
static float resample_filter_func(float t, float cutoff, int32_t num_zeros) {
  float window, filter;
  if (fabs((double)t) < num_zeros / (2.0 * cutoff))
    window = (float)(0.5 * (1 + cos(M_2PI * cutoff / num_zeros * t)));
  else window = 0.0f;
  if (t != 0.0f) filter = (float)(sin(M_2PI * cutoff * t) / (M_PI * t));
  else filter = (float)(2.0 * cutoff);
  return filter * window;
}

static int64_t linear_resample_num_out(int64_t n_in, int32_t rate_in, int32_t rate_out,
                                       float cutoff, int32_t num_zeros, int flush) {
  int32_t base = gcd_i(rate_in, rate_out);
  int64_t tick_freq = (int64_t)rate_in / base * rate_out; /* lcm */
  int64_t ticks_per_in = tick_freq / rate_in;
  int64_t interval = n_in * ticks_per_in;
  if (!flush) {
    float window_width = (float)(num_zeros / (2.0 * cutoff));
    int32_t ww_ticks = (int32_t)floor((double)(window_width * (float)tick_freq));
    interval -= ww_ticks;
  }
  if (interval <= 0) return 0;
  int64_t ticks_per_out = tick_freq / rate_out;
  int64_t last = interval / ticks_per_out;
  if (last * ticks_per_out == interval) last--;
  return last + 1;
}

/* produce all `n_out` outputs from the whole input (zeros beyond the ends):
 * identical to the streamed Resample() calls because each output is a fixed
 * dot product with input samples wherever they exist. */
static void linear_resample(const float *in, int64_t n_in, int32_t rate_in,
                            int32_t rate_out, float cutoff, int32_t num_zeros,
                            float *out, int64_t n_out) {
  int32_t base = gcd_i(rate_in, rate_out);
  int32_t in_unit = rate_in / base, out_unit = rate_out / base;
  double window_width = num_zeros / (2.0 * cutoff);
  int32_t *first_index = (int32_t *)malloc(sizeof(int32_t) * out_unit);
  int32_t *nw = (int32_t *)malloc(sizeof(int32_t) * out_unit);
  float **weights = (float **)malloc(sizeof(float *) * out_unit);
  for (int32_t i = 0; i < out_unit; i++) {
    double output_t = i / (double)rate_out;
    double min_t = output_t - window_width, max_t = output_t + window_width;
    int32_t min_idx = (int32_t)ceil(min_t * rate_in), max_idx = (int32_t)floor(max_t * rate_in);
    first_index[i] = min_idx; nw[i] = max_idx - min_idx + 1;
    weights[i] = (float *)malloc(sizeof(float) * nw[i]);
    for (int32_t j = 0; j < nw[i]; j++) {
      double input_t = (min_idx + j) / (double)rate_in, delta_t = input_t - output_t;
      weights[i][j] = resample_filter_func((float)delta_t, cutoff, num_zeros) / (float)rate_in;
    }
  }
  for (int64_t so = 0; so < n_out; so++) {
    int64_t unit = so / out_unit;
    int32_t wrapped = (int32_t)(so - unit * out_unit);
    int64_t first_in = first_index[wrapped] + unit * in_unit;
    double acc = 0.0;
    for (int32_t j = 0; j < nw[wrapped]; j++) {
      int64_t idx = first_in + j;
      if (idx >= 0 && idx < n_in) acc += (double)weights[wrapped][j] * in[idx];
    }
    out[so] = (float)acc;
  }
  for (int32_t i = 0; i < out_unit; i++) free(weights[i]);
  free(weights); free(first_index); free(nw);
}

/* Whole-signal resampling with Kaldi's LinearResample, flushed (what
 * kaldi::ResampleWaveform does: resample.cc; cutoff <= 0 selects its default
 * 0.99 * 0.5 * min(rate_in, rate_out), num_zeros <= 0 its default 6).  The
 * step before the path (reference: shennong/audio.py:358-423 calls
 * sox or scipy for the same purpose). */
int64_t orc_resample_num_out(int64_t n_in, int32_t rate_in, int32_t rate_out) {
  return linear_resample_num_out(n_in, rate_in, rate_out, 1.0f, 1, 1);
}
void orc_resample(const float *in, int64_t n_in, int32_t rate_in, int32_t rate_out,
                  float cutoff, int32_t num_zeros, float *out /*[orc_resample_num_out]*/) {
  if (cutoff <= 0.0f) cutoff = 0.99f * 0.5f * (float)(rate_in < rate_out ? rate_in : rate_out);
  if (num_zeros <= 0) num_zeros = 6;
  linear_resample(in, n_in, rate_in, rate_out, cutoff, num_zeros, out,
                  orc_resample_num_out(n_in, rate_in, rate_out));
}

static int32_t nccf_window_size(const orc_pitch_opts *o) {
  return (int32_t)((double)o->resample_freq * (double)o->frame_length_ms / 1000.0);
}
static int32_t nccf_window_shift(const orc_pitch_opts *o) {
  return (int32_t)((double)o->resample_freq * (double)o->frame_shift_ms / 1000.0);
}

static void pitch_lag_range(const orc_pitch_opts *o, int32_t *first, int32_t *last) {
  double outer_min_lag = 1.0 / o->max_f0 - (o->upsample_filter_width / (2.0 * o->resample_freq));
  double outer_max_lag = 1.0 / o->min_f0 + (o->upsample_filter_width / (2.0 * o->resample_freq));
  *first = (int32_t)ceil(o->resample_freq * outer_min_lag);
  *last = (int32_t)floor(o->resample_freq * outer_max_lag);
}

/* SelectLags: float loop `lag *= 1.0 + delta_pitch` */
static int32_t select_lags(const orc_pitch_opts *o, float *lags) {
  float min_lag = (float)(1.0 / o->max_f0), max_lag = (float)(1.0 / o->min_f0);
  int32_t n = 0;
  for (float lag = min_lag; lag <= max_lag; lag = (float)(lag * (1.0 + o->delta_pitch))) {
    if (lags) lags[n] = lag;
    n++;
  }
  return n;
}

int32_t orc_pitch_num_lags(const orc_pitch_opts *o) { return select_lags(o, NULL); }

static int64_t pitch_frames_available(int64_t n_down, const orc_pitch_opts *o,
                                      int32_t last_lag, int input_finished) {
  int32_t shift = nccf_window_shift(o), length = nccf_window_size(o);
  if (!input_finished) length += last_lag;
  if (n_down < length) return 0;
  if (!o->snip_edges) {
    if (input_finished) return (int64_t)((float)n_down * 1.0f / (float)shift + 0.5f);
    return (int64_t)((float)(n_down - length / 2) * 1.0f / (float)shift + 0.5f);
  }
  return (n_down - length) / shift + 1;
}

int64_t orc_pitch_num_frames(int64_t nsamples, const orc_pitch_opts *o) {
  int32_t first, last;
  pitch_lag_range(o, &first, &last);
  int64_t m = linear_resample_num_out(nsamples, (int32_t)o->samp_freq, (int32_t)o->resample_freq,
                                      o->lowpass_cutoff, o->lowpass_filter_width, 1);
  return pitch_frames_available(m, o, last, 1);
}

typedef struct {
  int32_t nstates;
  int32_t *backpointer; /* [nstates] */
  float *pov_nccf;      /* [nstates] */
} pitch_frame_info;

/* PitchFrameInfo::ComputeBacktraces, verbatim search strategy */
static void compute_backtraces(const orc_pitch_opts *o, const float *nccf_pitch,
                               const float *lags, int32_t ns, const float *prev_fc,
                               int32_t *bounds_lo, int32_t *bounds_hi,
                               int32_t *backpointer, float *this_fc) {
  float *local_cost = (float *)malloc(sizeof(float) * ns);
  for (int32_t i = 0; i < ns; i++) {
    float lc = 1.0f;
    lc = lc + -1.0f * nccf_pitch[i];
    lc = o->soft_min_f0 * lags[i] * nccf_pitch[i] + 1.0f * lc;
    local_cost[i] = lc;
  }
  const float delta_pitch_sq = (float)pow(log(1.0 + (double)o->delta_pitch), 2.0);
  const float inter_frame_factor = delta_pitch_sq * o->penalty_factor;
  int32_t last_bp = 0;
  for (int32_t i = 0; i < ns; i++) {
    int32_t start_j = last_bp;
    float best_cost = (float)((start_j - i) * (start_j - i)) * inter_frame_factor + prev_fc[start_j];
    int32_t best_j = start_j;
    for (int32_t j = start_j + 1; j < ns; j++) {
      float c = (float)((j - i) * (j - i)) * inter_frame_factor + prev_fc[j];
      if (c < best_cost) { best_cost = c; best_j = j; }
      else break;
    }
    backpointer[i] = best_j; this_fc[i] = best_cost;
    bounds_lo[i] = best_j; bounds_hi[i] = ns - 1;
    last_bp = best_j;
  }
  for (int32_t iter = 0; iter < ns; iter++) {
    int changed = 0;
    if (iter % 2 == 0) {
      last_bp = ns - 1;
      for (int32_t i = ns - 1; i >= 0; i--) {
        int32_t lower = bounds_lo[i];
        int32_t upper = last_bp < bounds_hi[i] ? last_bp : bounds_hi[i];
        if (upper == lower) { last_bp = lower; continue; }
        float best_cost = this_fc[i];
        int32_t best_j = backpointer[i], initial = best_j;
        if (best_j == upper) { last_bp = best_j; continue; }
        for (int32_t j = upper; j > lower + 1; j--) {
          float c = (float)((j - i) * (j - i)) * inter_frame_factor + prev_fc[j];
          if (c < best_cost) { best_cost = c; best_j = j; }
          else if (best_j > j) break;
        }
        bounds_hi[i] = best_j;
        if (best_j != initial) { this_fc[i] = best_cost; backpointer[i] = best_j; changed = 1; }
        last_bp = best_j;
      }
    } else {
      last_bp = 0;
      for (int32_t i = 0; i < ns; i++) {
        int32_t lower = last_bp > bounds_lo[i] ? last_bp : bounds_lo[i];
        int32_t upper = bounds_hi[i];
        if (upper == lower) { last_bp = lower; continue; }
        float best_cost = this_fc[i];
        int32_t best_j = backpointer[i], initial = best_j;
        if (best_j == lower) { last_bp = best_j; continue; }
        for (int32_t j = lower; j < upper - 1; j++) {
          float c = (float)((j - i) * (j - i)) * inter_frame_factor + prev_fc[j];
          if (c < best_cost) { best_cost = c; best_j = j; }
          else if (best_j < j) break;
        }
        bounds_lo[i] = best_j;
        if (best_j != initial) { this_fc[i] = best_cost; backpointer[i] = best_j; changed = 1; }
        last_bp = best_j;
      }
    }
    if (!changed) break;
  }
  for (int32_t i = 0; i < ns; i++) this_fc[i] += local_cost[i];
  free(local_cost);
}

int64_t orc_compute_kaldi_pitch(const float *wave, int64_t nsamples,
                                const orc_pitch_opts *o, float *out) {
  int32_t rate_in = (int32_t)o->samp_freq, rate_out = (int32_t)o->resample_freq;
  int32_t first_lag, last_lag;
  pitch_lag_range(o, &first_lag, &last_lag);
  int32_t nmeas = last_lag + 1 - first_lag;
  int32_t ns = select_lags(o, NULL);
  float *lags = (float *)malloc(sizeof(float) * ns);
  select_lags(o, lags);
  int32_t shift = nccf_window_shift(o), basic_len = nccf_window_size(o);
  int32_t full_len = basic_len + last_lag;

  /* phase 1: AcceptWaveform(whole wave), resampler not flushed;
   * phase 2: InputFinished() -> AcceptWaveform(empty) with flush. */
  int64_t m1 = linear_resample_num_out(nsamples, rate_in, rate_out, o->lowpass_cutoff,
                                       o->lowpass_filter_width, 0);
  int64_t m2 = linear_resample_num_out(nsamples, rate_in, rate_out, o->lowpass_cutoff,
                                       o->lowpass_filter_width, 1);
  float *down = (float *)malloc(sizeof(float) * (m2 > 0 ? m2 : 1));
  linear_resample(wave, nsamples, rate_in, rate_out, o->lowpass_cutoff,
                  o->lowpass_filter_width, down, m2);
  int64_t end1 = pitch_frames_available(m1, o, last_lag, 0);
  int64_t end2 = pitch_frames_available(m2, o, last_lag, 1);
  if (end2 < end1) end2 = end1;
  int64_t nframes = end2;
  if (nframes == 0) { free(lags); free(down); return 0; }

  /* ArbitraryResample weights (float arithmetic as in resample.cc) */
  float upsample_cutoff = o->resample_freq * 0.5f;
  int32_t *up_first = (int32_t *)malloc(sizeof(int32_t) * ns);
  int32_t *up_n = (int32_t *)malloc(sizeof(int32_t) * ns);
  float **up_w = (float **)malloc(sizeof(float *) * ns);
  {
    float filter_width = (float)(o->upsample_filter_width / (2.0 * upsample_cutoff));
    for (int32_t i = 0; i < ns; i++) {
      float t = lags[i] + (-(float)first_lag / o->resample_freq);
      float t_min = t - filter_width, t_max = t + filter_width;
      int32_t imin = (int32_t)ceil((double)(o->resample_freq * t_min));
      int32_t imax = (int32_t)floor((double)(o->resample_freq * t_max));
      if (imin < 0) imin = 0;
      if (imax >= nmeas) imax = nmeas - 1;
      up_first[i] = imin; up_n[i] = imax - imin + 1;
      up_w[i] = (float *)malloc(sizeof(float) * (up_n[i] > 0 ? up_n[i] : 1));
      for (int32_t j = 0; j < up_n[i]; j++) {
        float delta_t = t - (float)(imin + j) / o->resample_freq;
        up_w[i][j] = resample_filter_func(delta_t, upsample_cutoff, o->upsample_filter_width) /
                     o->resample_freq;
      }
    }
  }

  float *nccf_pitch_rs = (float *)malloc(sizeof(float) * (size_t)nframes * ns);
  float *nccf_pov_rs = (float *)malloc(sizeof(float) * (size_t)nframes * ns);
  float *window = (float *)malloc(sizeof(float) * full_len);
  float *zm = (float *)malloc(sizeof(float) * full_len);
  float *inner = (float *)malloc(sizeof(float) * nmeas);
  float *normp = (float *)malloc(sizeof(float) * nmeas);
  float *nccf_pitch = (float *)malloc(sizeof(float) * nmeas);
  float *nccf_pov = (float *)malloc(sizeof(float) * nmeas);

  double sumsq1 = 0.0, sum1 = 0.0, sumsq2 = 0.0, sum2 = 0.0;
  for (int64_t i = 0; i < m1; i++) { sumsq1 += (double)down[i] * down[i]; sum1 += down[i]; }
  sumsq2 = sumsq1; sum2 = sum1;
  for (int64_t i = m1; i < m2; i++) { sumsq2 += (double)down[i] * down[i]; sum2 += down[i]; }

  for (int64_t f = 0; f < nframes; f++) {
    int phase2 = f >= end1;
    int64_t avail = phase2 ? m2 : m1;
    double cur_sumsq = phase2 ? sumsq2 : sumsq1, cur_sum = phase2 ? sum2 : sum1;
    double cur_n = (double)avail;
    int64_t start;
    if (o->snip_edges) start = f * shift;
    else start = (int64_t)(((double)f + 0.5) * shift) - full_len / 2;
    for (int32_t s = 0; s < full_len; s++) {
      int64_t k = start + s;
      window[s] = (k >= 0 && k < avail) ? down[k] : 0.0f;
    }
    if (o->preemph_coeff != 0.0f) {
      for (int32_t i = full_len - 1; i > 0; i--) window[i] -= o->preemph_coeff * window[i - 1];
      window[0] *= (float)(1.0 - o->preemph_coeff);
    }
    double mean_square = cur_sumsq / cur_n - pow(cur_sum / cur_n, 2.0);
    /* ComputeCorrelation */
    {
      double s = 0.0;
      for (int32_t i = 0; i < basic_len; i++) s += window[i];
      float mean = (float)(s / basic_len);
      for (int32_t i = 0; i < full_len; i++) zm[i] = window[i] + -mean;
      float e1 = dotf(zm, zm, basic_len);
      for (int32_t lag = first_lag; lag <= last_lag; lag++) {
        float e2 = dotf(zm + lag, zm + lag, basic_len);
        float sum = dotf(zm, zm + lag, basic_len);
        inner[lag - first_lag] = sum;
        normp[lag - first_lag] = e1 * e2;
      }
    }
    double ballast_pitch = pow(mean_square * basic_len, 2) * o->nccf_ballast;
    for (int32_t l = 0; l < nmeas; l++) {
      float bp = (float)ballast_pitch;
      float den = (float)pow((double)(normp[l] + bp), 0.5);
      nccf_pitch[l] = den != 0.0f ? inner[l] / den : 0.0f;
      float den0 = (float)pow((double)(normp[l] + 0.0f), 0.5);
      nccf_pov[l] = den0 != 0.0f ? inner[l] / den0 : 0.0f;
    }
    for (int32_t i = 0; i < ns; i++) {
      nccf_pitch_rs[(size_t)f * ns + i] = dotf(nccf_pitch + up_first[i], up_w[i], up_n[i]);
      nccf_pov_rs[(size_t)f * ns + i] = dotf(nccf_pov + up_first[i], up_w[i], up_n[i]);
    }
  }

  /* Viterbi */
  int32_t *bp = (int32_t *)malloc(sizeof(int32_t) * (size_t)nframes * ns);
  float *fc = (float *)calloc(ns, sizeof(float));
  float *nfc = (float *)calloc(ns, sizeof(float));
  int32_t *blo = (int32_t *)malloc(sizeof(int32_t) * ns);
  int32_t *bhi = (int32_t *)malloc(sizeof(int32_t) * ns);
  for (int64_t f = 0; f < nframes; f++) {
    compute_backtraces(o, nccf_pitch_rs + (size_t)f * ns, lags, ns, fc, blo, bhi,
                       bp + (size_t)f * ns, nfc);
    float *t = fc; fc = nfc; nfc = t;
    float mn = fc[0];
    for (int32_t i = 1; i < ns; i++) if (fc[i] < mn) mn = fc[i];
    for (int32_t i = 0; i < ns; i++) fc[i] += -mn;
  }
  /* RecomputeBacktraces() is a no-op offline: every frame's stored
   * mean_square is within 1% of the final one (checked below). */
  int32_t best = 0;
  for (int32_t i = 1; i < ns; i++) if (fc[i] < fc[best]) best = i;
  for (int64_t f = nframes - 1; f >= 0; f--) {
    out[2 * f] = nccf_pov_rs[(size_t)f * ns + best];
    out[2 * f + 1] = (float)(1.0 / (double)lags[best]);
    best = bp[(size_t)f * ns + best];
  }
  free(bp); free(fc); free(nfc); free(blo); free(bhi);
  free(nccf_pitch_rs); free(nccf_pov_rs); free(window); free(zm); free(inner);
  free(normp); free(nccf_pitch); free(nccf_pov);
  for (int32_t i = 0; i < ns; i++) free(up_w[i]);
  free(up_w); free(up_first); free(up_n); free(lags); free(down);
  return nframes;
}

/* ProcessPitch (OnlineProcessPitch in offline use; pitch_kaldi.py:536-537) */
int32_t orc_process_pitch_dim(const orc_pitch_post_opts *o) {
  return (o->add_pov_feature != 0) + (o->add_normalized_log_pitch != 0) +
         (o->add_delta_pitch != 0) + (o->add_raw_log_pitch != 0);
}

static float nccf_to_pov_feature(float n) {
  if (n > 1.0f) n = 1.0f; else if (n < -1.0f) n = -1.0f;
  return (float)(pow((1.0001 - (double)n), 0.15) - 1.0);
}
static float nccf_to_pov(float n) {
  float ndash = fabsf(n);
  if (ndash > 1.0f) ndash = 1.0f;
  float r = (float)(-5.2 + 5.4 * exp(7.5 * ((double)ndash - 1.0)) + 4.8 * (double)ndash -
                    2.0 * exp(-10.0 * (double)ndash) + 4.2 * exp(20.0 * ((double)ndash - 1.0)));
  return (float)(1.0 / (1 + exp(-1.0 * (double)r)));
}

int64_t orc_process_pitch(const float *raw, int64_t nframes,
                          const orc_pitch_post_opts *o, float *out) {
  int32_t dim = orc_process_pitch_dim(o);
  if (nframes == 0) return 0;
  int64_t nout = nframes + o->delay;
  float *logp = (float *)malloc(sizeof(float) * nframes);
  float *pov = (float *)malloc(sizeof(float) * nframes);
  for (int64_t t = 0; t < nframes; t++) {
    logp[t] = (float)log((double)raw[2 * t + 1]);
    pov[t] = nccf_to_pov(raw[2 * t]);
  }
  for (int64_t fo = 0; fo < nout; fo++) {
    int64_t t = fo < o->delay ? 0 : fo - o->delay;
    int32_t idx = 0;
    float *row = out + (size_t)fo * dim;
    if (o->add_pov_feature)
      row[idx++] = o->pov_scale * nccf_to_pov_feature(raw[2 * t]) + o->pov_offset;
    if (o->add_normalized_log_pitch) {
      int64_t b = t - o->normalization_left_context; if (b < 0) b = 0;
      int64_t e = t + o->normalization_right_context + 1; if (e > nframes) e = nframes;
      double sp = 0.0, slp = 0.0;
      for (int64_t f = b; f < e; f++) { sp += pov[f]; slp += pov[f] * logp[f]; }
      float avg = (float)(slp / sp);
      row[idx++] = (logp[t] - avg) * o->pitch_scale;
    }
    if (o->add_delta_pitch) {
      /* order-1 delta with window delta_window on the clipped neighbourhood;
       * noise term omitted (stochastic in the reference): parity at stddev 0 */
      int32_t w = o->delta_window;
      float norm = 0.0f;
      for (int32_t j = -w; j <= w; j++) norm += (float)(j * j);
      float inv = (float)(1.0 / (double)norm);
      float d = 0.0f;
      for (int32_t j = -w; j <= w; j++) {
        int64_t tt = t + j;
        if (tt < 0) tt = 0; else if (tt >= nframes) tt = nframes - 1;
        float s = (float)j * inv;
        if (s != 0.0f) d += s * logp[tt];
      }
      row[idx++] = (d + 0.0f) * o->delta_pitch_scale;
    }
    if (o->add_raw_log_pitch) row[idx++] = logp[t];
  }
  free(logp); free(pov);
  return nout;
}

/* ------------------------------------------------------------------------ */
/* config-3 style pipeline for the CPU baseline                              */
int64_t orc_pipeline_batch(const int16_t *pcm, const int64_t *sample_offsets,
                           const int64_t *frame_offsets, int64_t nutts,
                           const orc_frame_opts *fo, const orc_mel_opts *mo,
                           const orc_feat_opts *xo, int32_t do_cmvn,
                           int32_t norm_vars, int32_t delta_order,
                           int32_t delta_window, float *out, int32_t nthreads) {
  orc_plan p;
  if (plan_init(&p, fo, mo, xo, 1.0f) != 0) { plan_free(&p); return -1; }
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  int32_t odim = p.dim * (delta_order + 1);
  int64_t total = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
  for (int64_t u = 0; u < nutts; u++) {
    int64_t n = sample_offsets[u + 1] - sample_offsets[u];
    int64_t nf = frame_offsets[u + 1] - frame_offsets[u];
    float *wave = (float *)malloc(sizeof(float) * (n > 0 ? n : 1));
    for (int64_t i = 0; i < n; i++) wave[i] = (float)pcm[sample_offsets[u] + i];
    float *base = (float *)malloc(sizeof(float) * (size_t)(nf > 0 ? nf : 1) * p.dim);
    int64_t got = compute_with_plan(&p, wave, n, base, NULL);
    if (got > 0 && do_cmvn) {
      double *stats = (double *)calloc((size_t)2 * (p.dim + 1), sizeof(double));
      orc_cmvn_accumulate(base, got, p.dim, NULL, stats);
      orc_cmvn_apply(stats, p.dim, norm_vars, 0, base, got);
      free(stats);
    }
    if (got > 0) {
      if (delta_order > 0)
        orc_compute_deltas(base, got, p.dim, delta_order, delta_window,
                           out + (size_t)frame_offsets[u] * odim);
      else
        memcpy(out + (size_t)frame_offsets[u] * odim, base, sizeof(float) * (size_t)got * p.dim);
    }
    total += got;
    free(base); free(wave);
  }
  plan_free(&p);
  return total;
}
