// pitch.cu -- Kaldi pitch extraction and post-processing for ragged batches.
//
// Replaces kaldi.feat.pitch.compute_kaldi_pitch / process_pitch
// (shennong/processor/pitch_kaldi.py:296-299, 535-537), i.e. Kaldi's
// pitch-functions.cc + resample.cc in OFFLINE use: one AcceptWaveform() with
// the resampler unflushed followed by InputFinished().
//
//   k1  resample_kernel      16k -> 4k windowed-sinc (LinearResample), one
//                            thread per output sample, all utterances at once
//   k2  pitch_track_kernel   one CTA per utterance, sequential over frames:
//                            NCCF at the integer lags, sinc-upsample to the
//                            log-spaced lags (ArbitraryResample), one exact
//                            Viterbi step (min-plus over all states), back-
//                            pointers to a per-CTA scratch; then backtrace and
//                            the (NCCF, pitch) rows
//   k3  process_pitch_kernel POV / normalised log-pitch / delta / raw log-pitch
#include <cfloat>
#include <cmath>
#include <cstring>

#include "device_utils.cuh"
#include "snb_internal.h"

namespace snb {

constexpr int kPitchThreads = 256;
constexpr int kMaxStatesPerThread = 8;

struct PitchTables {
  // host copies
  int32_t rate_in, rate_out, in_unit, out_unit;
  int32_t first_lag, last_lag, nmeas, nstates;
  int32_t shift, basic_len, full_len;
  int32_t down_nw_max, up_nw_max;
  std::vector<int32_t> down_first, down_nw;     // [out_unit]
  std::vector<float> down_w;                    // [out_unit, down_nw_max]
  std::vector<float> lags, pen;                 // [nstates]
  std::vector<int32_t> up_first, up_nw;         // [nstates]
  std::vector<float> up_w;                      // [nstates, up_nw_max]
  // device copies (one allocation)
  void *d_blob = nullptr;
  const int32_t *d_down_first, *d_down_nw, *d_up_first, *d_up_nw;
  const float *d_down_w, *d_lags, *d_pen, *d_up_w;
};

static int32_t gcd_i32(int32_t a, int32_t b) {
  while (b) { const int32_t t = a % b; a = b; b = t; }
  return a;
}

// LinearResample::FilterFunc / ArbitraryResample::FilterFunc
static float sinc_filter(float t, float cutoff, int32_t num_zeros) {
  const double two_pi = 6.283185307179586476925286766559005, pi = 3.1415926535897932384626433832795;
  float window = 0.0f, filter;
  if (std::fabs(static_cast<double>(t)) < num_zeros / (2.0 * cutoff))
    window = static_cast<float>(0.5 * (1 + std::cos(two_pi * cutoff / num_zeros * t)));
  if (t != 0.0f) filter = static_cast<float>(std::sin(two_pi * cutoff * t) / (pi * t));
  else filter = static_cast<float>(2.0 * cutoff);
  return filter * window;
}

static int64_t resample_num_out(int64_t n_in, const snb_pitch_opts &o, bool flush) {
  const int32_t rate_in = static_cast<int32_t>(o.samp_freq), rate_out = static_cast<int32_t>(o.resample_freq);
  const int32_t base = gcd_i32(rate_in, rate_out);
  const int64_t tick_freq = static_cast<int64_t>(rate_in) / base * rate_out;
  int64_t interval = n_in * (tick_freq / rate_in);
  if (!flush) {
    const float window_width = static_cast<float>(o.lowpass_filter_width / (2.0 * o.lowpass_cutoff));
    interval -= static_cast<int32_t>(std::floor(static_cast<double>(window_width * static_cast<float>(tick_freq))));
  }
  if (interval <= 0) return 0;
  const int64_t per_out = tick_freq / rate_out;
  int64_t last = interval / per_out;
  if (last * per_out == interval) --last;
  return last + 1;
}

static void lag_range(const snb_pitch_opts &o, int32_t *first, int32_t *last) {
  const double outer_min = 1.0 / o.max_f0 - o.upsample_filter_width / (2.0 * o.resample_freq);
  const double outer_max = 1.0 / o.min_f0 + o.upsample_filter_width / (2.0 * o.resample_freq);
  *first = static_cast<int32_t>(std::ceil(o.resample_freq * outer_min));
  *last = static_cast<int32_t>(std::floor(o.resample_freq * outer_max));
}

static int32_t nccf_win_size(const snb_pitch_opts &o) {
  return static_cast<int32_t>(static_cast<double>(o.resample_freq) * o.frame_length_ms / 1000.0);
}
static int32_t nccf_win_shift(const snb_pitch_opts &o) {
  return static_cast<int32_t>(static_cast<double>(o.resample_freq) * o.frame_shift_ms / 1000.0);
}

static int64_t frames_available(int64_t n_down, const snb_pitch_opts &o, int32_t last_lag, bool finished) {
  const int32_t shift = nccf_win_shift(o);
  int32_t length = nccf_win_size(o);
  if (!finished) length += last_lag;
  if (shift <= 0 || n_down < length) return 0;
  if (!o.snip_edges) {
    if (finished) return static_cast<int64_t>(static_cast<float>(n_down) * 1.0f / static_cast<float>(shift) + 0.5f);
    return static_cast<int64_t>(static_cast<float>(n_down - length / 2) * 1.0f / static_cast<float>(shift) + 0.5f);
  }
  return (n_down - length) / shift + 1;
}

static bool pitch_opts_valid(const snb_pitch_opts &o) {
  return o.samp_freq > 0 && o.resample_freq > 0 && o.min_f0 > 0 && o.max_f0 > o.min_f0 &&
         o.lowpass_cutoff > 0 && o.delta_pitch > 0 && o.frame_shift_ms > 0 && o.frame_length_ms > 0 &&
         o.lowpass_filter_width > 0 && o.upsample_filter_width > 0 &&
         o.resample_freq > 2 * o.lowpass_cutoff * 0.999f && o.samp_freq >= o.resample_freq;
}

int pitch_plan_init(snb_plan *plan) {
  const snb_pitch_opts &o = plan->po;
  if (!pitch_opts_valid(o)) return set_error(SNB_ERR_OPTION, "invalid pitch extraction options");
  PitchTables *t = new PitchTables();
  plan->pitch = t;
  t->rate_in = static_cast<int32_t>(o.samp_freq);
  t->rate_out = static_cast<int32_t>(o.resample_freq);
  const int32_t base = gcd_i32(t->rate_in, t->rate_out);
  t->in_unit = t->rate_in / base;
  t->out_unit = t->rate_out / base;
  // --- LinearResample::SetIndexesAndWeights ---
  const double window_width = o.lowpass_filter_width / (2.0 * o.lowpass_cutoff);
  t->down_first.resize(t->out_unit);
  t->down_nw.resize(t->out_unit);
  std::vector<std::vector<float>> dw(t->out_unit);
  t->down_nw_max = 0;
  for (int32_t i = 0; i < t->out_unit; ++i) {
    const double output_t = i / static_cast<double>(t->rate_out);
    const int32_t min_idx = static_cast<int32_t>(std::ceil((output_t - window_width) * t->rate_in));
    const int32_t max_idx = static_cast<int32_t>(std::floor((output_t + window_width) * t->rate_in));
    t->down_first[i] = min_idx;
    t->down_nw[i] = max_idx - min_idx + 1;
    for (int32_t j = 0; j < t->down_nw[i]; ++j) {
      const double delta_t = (min_idx + j) / static_cast<double>(t->rate_in) - output_t;
      dw[i].push_back(sinc_filter(static_cast<float>(delta_t), o.lowpass_cutoff, o.lowpass_filter_width) /
                      static_cast<float>(t->rate_in));
    }
    t->down_nw_max = std::max(t->down_nw_max, t->down_nw[i]);
  }
  t->down_w.assign(static_cast<size_t>(t->out_unit) * t->down_nw_max, 0.0f);
  for (int32_t i = 0; i < t->out_unit; ++i)
    std::copy(dw[i].begin(), dw[i].end(), t->down_w.begin() + static_cast<size_t>(i) * t->down_nw_max);
  // --- lags (SelectLags) ---
  lag_range(o, &t->first_lag, &t->last_lag);
  t->nmeas = t->last_lag + 1 - t->first_lag;
  {
    const float min_lag = static_cast<float>(1.0 / o.max_f0), max_lag = static_cast<float>(1.0 / o.min_f0);
    for (float lag = min_lag; lag <= max_lag; lag = static_cast<float>(lag * (1.0 + o.delta_pitch)))
      t->lags.push_back(lag);
  }
  t->nstates = static_cast<int32_t>(t->lags.size());
  t->shift = nccf_win_shift(o);
  t->basic_len = nccf_win_size(o);
  t->full_len = t->basic_len + t->last_lag;
  if (t->nmeas <= 0 || t->nstates <= 0 || t->shift <= 0 || t->basic_len <= 0)
    return set_error(SNB_ERR_OPTION, "invalid pitch extraction options");
  if (t->nstates > kPitchThreads * kMaxStatesPerThread || t->full_len > 4096 || t->nmeas > 1024)
    return set_error(SNB_ERR_UNSUPPORTED, "pitch options outside the GPU path limits");
  // --- ArbitraryResample (float arithmetic as in resample.cc) ---
  const float up_cutoff = o.resample_freq * 0.5f;
  const float filter_width = static_cast<float>(o.upsample_filter_width / (2.0 * up_cutoff));
  t->up_first.resize(t->nstates);
  t->up_nw.resize(t->nstates);
  std::vector<std::vector<float>> uw(t->nstates);
  t->up_nw_max = 1;
  for (int32_t i = 0; i < t->nstates; ++i) {
    const float tt = t->lags[i] + (-static_cast<float>(t->first_lag) / o.resample_freq);
    int32_t imin = static_cast<int32_t>(std::ceil(static_cast<double>(o.resample_freq * (tt - filter_width))));
    int32_t imax = static_cast<int32_t>(std::floor(static_cast<double>(o.resample_freq * (tt + filter_width))));
    if (imin < 0) imin = 0;
    if (imax >= t->nmeas) imax = t->nmeas - 1;
    t->up_first[i] = imin;
    t->up_nw[i] = std::max(0, imax - imin + 1);
    for (int32_t j = 0; j < t->up_nw[i]; ++j) {
      const float delta_t = tt - static_cast<float>(imin + j) / o.resample_freq;
      uw[i].push_back(sinc_filter(delta_t, up_cutoff, o.upsample_filter_width) / o.resample_freq);
    }
    t->up_nw_max = std::max(t->up_nw_max, t->up_nw[i]);
  }
  t->up_w.assign(static_cast<size_t>(t->nstates) * t->up_nw_max, 0.0f);
  for (int32_t i = 0; i < t->nstates; ++i)
    std::copy(uw[i].begin(), uw[i].end(), t->up_w.begin() + static_cast<size_t>(i) * t->up_nw_max);
  // --- Viterbi transition penalties: (i-j)^2 * inter_frame_factor ---
  const float delta_pitch_sq = static_cast<float>(std::pow(std::log(1.0 + static_cast<double>(o.delta_pitch)), 2.0));
  const float factor = delta_pitch_sq * o.penalty_factor;
  t->pen.resize(t->nstates);
  for (int32_t d = 0; d < t->nstates; ++d) t->pen[d] = static_cast<float>(d * d) * factor;
  // --- upload ---
  std::vector<int32_t> blob;
  auto push_i = [&](const std::vector<int32_t> &v) { size_t off = blob.size(); blob.insert(blob.end(), v.begin(), v.end()); while (blob.size() % 4) blob.push_back(0); return off; };
  auto push_f = [&](const std::vector<float> &v) {
    size_t off = blob.size();
    blob.resize(off + v.size());
    std::memcpy(blob.data() + off, v.data(), v.size() * 4);
    while (blob.size() % 4) blob.push_back(0);
    return off;
  };
  const size_t o1 = push_i(t->down_first), o2 = push_i(t->down_nw), o3 = push_f(t->down_w),
               o4 = push_f(t->lags), o5 = push_f(t->pen), o6 = push_i(t->up_first), o7 = push_i(t->up_nw),
               o8 = push_f(t->up_w);
  int32_t *d = nullptr;
  cudaError_t e = cudaMalloc(&d, blob.size() * 4);
  if (e == cudaSuccess) e = upload(d, blob.data(), blob.size() * 4);
  if (e != cudaSuccess) {
    cudaGetLastError();
    if (d) cudaFree(d);
    return set_error(SNB_ERR_CUDA, "cannot upload pitch tables: %s", cudaGetErrorString(e));
  }
  t->d_blob = d;
  t->d_down_first = d + o1; t->d_down_nw = d + o2;
  t->d_down_w = reinterpret_cast<const float *>(d + o3);
  t->d_lags = reinterpret_cast<const float *>(d + o4);
  t->d_pen = reinterpret_cast<const float *>(d + o5);
  t->d_up_first = d + o6; t->d_up_nw = d + o7;
  t->d_up_w = reinterpret_cast<const float *>(d + o8);
  return SNB_OK;
}

void pitch_plan_free(snb_plan *plan) {
  if (!plan->pitch) return;
  if (plan->pitch->d_blob) cudaFree(plan->pitch->d_blob);
  delete plan->pitch;
  plan->pitch = nullptr;
}

// per-utterance phase info packed as int64 x 4: down_offset, m1, m2, end1
int pitch_batch_init(const snb_plan *plan, snb_batch *b) {
  const snb_pitch_opts &o = plan->po;
  const PitchTables *t = plan->pitch;
  std::vector<int64_t> info(static_cast<size_t>(b->nutts) * 4 + 4, 0);
  int64_t off = 0;
  for (int64_t u = 0; u < b->nutts; ++u) {
    const int64_t n = b->sample_len[u];
    const int64_t m1 = resample_num_out(n, o, false), m2 = resample_num_out(n, o, true);
    int64_t end1 = frames_available(m1, o, t->last_lag, false);
    const int64_t end2 = b->frame_offsets[u + 1] - b->frame_offsets[u];
    if (end1 > end2) end1 = end2;
    info[4 * u] = off; info[4 * u + 1] = m1; info[4 * u + 2] = m2; info[4 * u + 3] = end1;
    off += m2;
  }
  info[4 * b->nutts] = off;
  b->total_down = off;
  b->down_offsets = info;       // uploaded by snb_batch_create with the other tables
  return SNB_OK;
}

// ---------------------------------------------------------------------------
// k1: LinearResample::Resample for the whole batch
// ---------------------------------------------------------------------------
struct ResampleArgs {
  const int16_t *pcm;
  const int64_t *sample_begin;
  const int64_t *sample_len;
  const int64_t *info;          // [nutts,4] down_offset, m1, m2, end1
  int64_t nutts, total_down;
  int32_t in_unit, out_unit, nw_max;
  const int32_t *first, *nw;
  const float *w;
  float *down;
};

__global__ void __launch_bounds__(256) resample_kernel(const ResampleArgs a) {
  const int64_t idx0 = static_cast<int64_t>(blockIdx.x) * blockDim.x;
  const int64_t idx = idx0 + threadIdx.x;
  // utterance of the CTA's first output sample (info[4u] is non-decreasing):
  // found ONCE per CTA -- the equal-length guess first, else a binary search
  // by thread 0 -- instead of 14 dependent loads in every thread; the other
  // threads walk forward from it (a CTA rarely spans more than two utterances)
  __shared__ int64_t s_u0;
  if (threadIdx.x == 0) {
    int64_t g = static_cast<int64_t>(static_cast<double>(idx0) * static_cast<double>(a.nutts) /
                                     static_cast<double>(a.total_down));
    g = min(max(g, static_cast<int64_t>(0)), a.nutts - 1);
    const bool ok = a.info[4 * g] <= idx0 && (g + 1 >= a.nutts || idx0 < a.info[4 * (g + 1)]);
    if (!ok) {
      int64_t lo = 0, hi = a.nutts;
      while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (a.info[4 * mid] <= idx0) lo = mid; else hi = mid;
      }
      g = lo;
    }
    s_u0 = g;
  }
  __syncthreads();
  if (idx >= a.total_down) return;
  int64_t u = s_u0;
  while (u + 1 < a.nutts && a.info[4 * (u + 1)] <= idx) ++u;
  const int64_t so = idx - a.info[4 * u];
  const int64_t in0 = a.sample_begin[u], n_in = a.sample_len[u];
  const int64_t unit = so / a.out_unit;
  const int32_t wrapped = static_cast<int32_t>(so - unit * a.out_unit);
  const int64_t first_in = a.first[wrapped] + unit * a.in_unit;
  const float *w = a.w + static_cast<int64_t>(wrapped) * a.nw_max;
  float acc = 0.0f;
  const int32_t nw = a.nw[wrapped];
  for (int32_t j = 0; j < nw; ++j) {
    const int64_t k = first_in + j;
    if (k >= 0 && k < n_in) acc = fmaf(w[j], static_cast<float>(a.pcm[in0 + k]), acc);
  }
  a.down[idx] = acc;
}

// ---------------------------------------------------------------------------
// k2: per-utterance NCCF + Viterbi
// ---------------------------------------------------------------------------
struct TrackArgs {
  const float *down;
  const int64_t *info;
  const int64_t *frame_offsets;
  int64_t nutts;
  int32_t first_lag, nmeas, nstates, shift, basic_len, full_len, up_nw_max;
  int32_t snip_edges;
  float preemph, soft_min_f0, nccf_ballast;
  const float *lags, *pen, *up_w;
  const int32_t *up_first, *up_nw;
  // per-CTA scratch
  int16_t *bp;            // [grid, max_frames, nstates]
  float *pov_raw;         // [grid, max_frames, nmeas]
  int32_t *states;        // [grid, max_frames]
  int64_t max_frames;
  float *out;
  int64_t ld_out;
  unsigned long long *queue;   // warp tracker: next utterance to hand out (zeroed before the launch)
};

__device__ __forceinline__ double block_sum_f64(double v, double *s_red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = group_sum_f64<32>(v);
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < kPitchThreads / 32; ++w) t += s_red[w];
  return t;
}

#ifndef SNB_PITCH_MINB
#define SNB_PITCH_MINB 4
#endif
__global__ void __launch_bounds__(kPitchThreads, SNB_PITCH_MINB) pitch_track_kernel(const TrackArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *s_win = reinterpret_cast<float *>(smem_raw);          // [full_len]
  float *s_np = s_win + a.full_len;                             // nccf_pitch [nmeas]
  float *s_nv = s_np + a.nmeas;                                 // nccf_pov   [nmeas]
  float *s_prev = s_nv + a.nmeas;                               // forward cost [nstates]
  float *s_new = s_prev + a.nstates;
  float *s_pen = s_new + a.nstates;
  float *s_lags = s_pen + a.nstates;
  float *s_upw = s_lags + a.nstates;                            // [nstates, up_nw_max]
  int32_t *s_upfirst = reinterpret_cast<int32_t *>(s_upw + a.nstates * a.up_nw_max);
  int32_t *s_upn = s_upfirst + a.nstates;
  __shared__ double s_red[kPitchThreads / 32];
  __shared__ int s_abp[kPitchThreads * kMaxStatesPerThread / 16 + 2];
  __shared__ float s_acost[kPitchThreads * kMaxStatesPerThread / 16 + 2];
  __shared__ float s_fred[kPitchThreads / 32];
  __shared__ int s_ired[kPitchThreads / 32];
  __shared__ float s_scalar[4];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ns = a.nstates, nm = a.nmeas;
  for (int i = tid; i < ns; i += kPitchThreads) {
    s_pen[i] = a.pen[i];
    s_lags[i] = a.lags[i];
    s_upfirst[i] = a.up_first[i];
    s_upn[i] = a.up_nw[i];
  }
  for (int i = tid; i < ns * a.up_nw_max; i += kPitchThreads) s_upw[i] = a.up_w[i];
  __syncthreads();

  int16_t *bp = a.bp + static_cast<int64_t>(blockIdx.x) * a.max_frames * ns;
  float *pov_raw = a.pov_raw + static_cast<int64_t>(blockIdx.x) * a.max_frames * nm;
  int32_t *states = a.states + static_cast<int64_t>(blockIdx.x) * a.max_frames;

  for (int64_t u = blockIdx.x; u < a.nutts; u += gridDim.x) {
    const int64_t doff = a.info[4 * u], m1 = a.info[4 * u + 1], m2 = a.info[4 * u + 2],
                  end1 = a.info[4 * u + 3];
    const int64_t row0 = a.frame_offsets[u], F = a.frame_offsets[u + 1] - row0;
    if (F <= 0) continue;
    const float *x = a.down + doff;
    // ---- global mean-square for the ballast (double sums, two phases) ----
    double p1 = 0.0, q1 = 0.0, p2 = 0.0, q2 = 0.0;
    for (int64_t i = tid; i < m2; i += kPitchThreads) {
      const double v = x[i];
      if (i < m1) { p1 += v; q1 += v * v; } else { p2 += v; q2 += v * v; }
    }
    const double sum1 = block_sum_f64(p1, s_red), sq1 = block_sum_f64(q1, s_red);
    const double sum2 = sum1 + block_sum_f64(p2, s_red), sq2 = sq1 + block_sum_f64(q2, s_red);
    const double ms1 = m1 > 0 ? sq1 / static_cast<double>(m1) - (sum1 / static_cast<double>(m1)) * (sum1 / static_cast<double>(m1)) : 0.0;
    const double ms2 = m2 > 0 ? sq2 / static_cast<double>(m2) - (sum2 / static_cast<double>(m2)) * (sum2 / static_cast<double>(m2)) : 0.0;
    const float ballast1 = static_cast<float>((ms1 * a.basic_len) * (ms1 * a.basic_len) * static_cast<double>(a.nccf_ballast));
    const float ballast2 = static_cast<float>((ms2 * a.basic_len) * (ms2 * a.basic_len) * static_cast<double>(a.nccf_ballast));
    for (int i = tid; i < ns; i += kPitchThreads) s_prev[i] = 0.0f;
    __syncthreads();

    for (int64_t f = 0; f < F; ++f) {
      const bool phase2 = f >= end1;
      const int64_t avail = phase2 ? m2 : m1;
      const float ballast = phase2 ? ballast2 : ballast1;
      int64_t start;
      if (a.snip_edges) start = f * a.shift;
      else start = static_cast<int64_t>((static_cast<double>(f) + 0.5) * a.shift) - a.full_len / 2;
      // ---- ExtractFrame (zeros outside the available signal) ----
      for (int i = tid; i < a.full_len; i += kPitchThreads) {
        const int64_t k = start + i;
        s_win[i] = (k >= 0 && k < avail) ? x[k] : 0.0f;
      }
      __syncthreads();
      if (a.preemph != 0.0f) {
        // window[i] -= c * window[i-1] (original values), window[0] *= (1 - c)
        float vals[16];
        int cnt = 0;
        for (int i = tid; i < a.full_len && cnt < 16; i += kPitchThreads, ++cnt)
          vals[cnt] = (i > 0) ? fmaf(-a.preemph, s_win[i - 1], s_win[i]) : s_win[0] * (1.0f - a.preemph);
        __syncthreads();
        cnt = 0;
        for (int i = tid; i < a.full_len && cnt < 16; i += kPitchThreads, ++cnt) s_win[i] = vals[cnt];
        __syncthreads();
      }
      // ---- ComputeCorrelation: subtract the mean of the first basic_len ----
      if (warp == 0) {
        float s = 0.0f;
        for (int i = lane; i < a.basic_len; i += 32) s += s_win[i];
        s = group_sum<32>(s);
        if (lane == 0) s_scalar[0] = __fdiv_rn(s, static_cast<float>(a.basic_len));
      }
      __syncthreads();
      const float mean = s_scalar[0];
      for (int i = tid; i < a.full_len; i += kPitchThreads) s_win[i] -= mean;
      __syncthreads();
      if (warp == 0) {
        float e = 0.0f;
        for (int i = lane; i < a.basic_len; i += 32) e = fmaf(s_win[i], s_win[i], e);
        e = group_sum<32>(e);
        if (lane == 0) s_scalar[1] = e;
      }
      __syncthreads();
      const float e1 = s_scalar[1];
      // ---- NCCF at the integer lags: 2 threads per lag ----
      for (int l0 = 0; l0 < nm; l0 += kPitchThreads / 2) {
        const int l = l0 + (tid >> 1), part = tid & 1;
        float e2 = 0.0f, inner = 0.0f;
        if (l < nm) {
          const int lag = a.first_lag + l;
          const int half = (a.basic_len + 1) / 2;
          const int i0 = part * half, i1 = min(a.basic_len, i0 + half);
          for (int i = i0; i < i1; ++i) {
            const float v = s_win[lag + i];
            e2 = fmaf(v, v, e2);
            inner = fmaf(s_win[i], v, inner);
          }
        }
        e2 += __shfl_xor_sync(SNB_FULL_MASK, e2, 1);
        inner += __shfl_xor_sync(SNB_FULL_MASK, inner, 1);
        if (l < nm && part == 0) {
          const float norm = __fmul_rn(e1, e2);
          const float den_p = sqrtf(__fadd_rn(norm, ballast));
          const float den_v = sqrtf(norm);
          s_np[l] = den_p != 0.0f ? __fdiv_rn(inner, den_p) : 0.0f;
          const float pv = den_v != 0.0f ? __fdiv_rn(inner, den_v) : 0.0f;
          s_nv[l] = pv;
          pov_raw[f * nm + l] = pv;
        }
      }
      __syncthreads();
      // ---- upsample nccf_pitch to the log-spaced lags; local cost ----
      for (int i = tid; i < ns; i += kPitchThreads) {
        const float *w = s_upw + i * a.up_nw_max;
        const int first = s_upfirst[i], n = s_upn[i];
        float acc = 0.0f;
        for (int j = 0; j < n; ++j) acc = fmaf(w[j], s_np[first + j], acc);
        // local_cost = 1 - nccf; += soft_min_f0 * lag * nccf
        float c = __fadd_rn(1.0f, -acc);
        c = __fadd_rn(__fmul_rn(__fmul_rn(a.soft_min_f0, s_lags[i]), acc), c);
        s_new[i] = c;                  // holds the local cost until the state is finalised
      }
      // ---- exact Viterbi step: min_j pen[|i-j|] + prev[j] (first minimal j) ----
      // The transition cost is convex in (i - j), so the minimising j is
      // non-decreasing in i (Monge property).  Level 1: every kAnchor-th state
      // (and the last one) scans all j, one warp per anchor.  Level 2: every
      // other state scans only [bp(left anchor), bp(right anchor)].  ~10x fewer
      // evaluations than the full 417 x 417 scan, same minimum.
      constexpr int kAnchor = 16;
      const int nanchor = (ns - 1 + kAnchor - 1) / kAnchor + 1;     // 0, 16, ..., ns-1
      for (int t = warp; t < nanchor; t += kPitchThreads / 32) {
        const int i = min(t * kAnchor, ns - 1);
        float best = FLT_MAX;
        int bj = 0x7fffffff;
        for (int j = lane; j < ns; j += 32) {
          const int d = j > i ? j - i : i - j;
          const float c = __fadd_rn(s_pen[d], s_prev[j]);
          if (c < best) { best = c; bj = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ob = __shfl_xor_sync(SNB_FULL_MASK, best, o);
          const int oj = __shfl_xor_sync(SNB_FULL_MASK, bj, o);
          if (ob < best || (ob == best && oj < bj)) { best = ob; bj = oj; }
        }
        if (lane == 0) { s_abp[t] = bj; s_acost[t] = best; }
      }
      __syncthreads();
      float lmin = FLT_MAX;
      for (int i = tid; i < ns; i += kPitchThreads) {
        const int ta = i / kAnchor;
        float best;
        int bj;
        if (i == ta * kAnchor || i == ns - 1) {
          const int t = (i == ns - 1) ? nanchor - 1 : ta;
          best = s_acost[t]; bj = s_abp[t];
        } else {
          const int ja = s_abp[ta], jb = s_abp[min(ta + 1, nanchor - 1)];
          const int jlo = min(ja, jb), jhi = max(ja, jb);
          best = FLT_MAX; bj = jlo;
          for (int j = jlo; j <= jhi; ++j) {
            const int d = j > i ? j - i : i - j;
            const float c = __fadd_rn(s_pen[d], s_prev[j]);
            if (c < best) { best = c; bj = j; }
          }
        }
        bp[f * ns + i] = static_cast<int16_t>(bj);
        const float v = __fadd_rn(best, s_new[i]);      // + local cost (own slot)
        s_new[i] = v;
        lmin = fminf(lmin, v);
      }
      // block min -> renormalise so the smallest forward cost is zero
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lmin = fminf(lmin, __shfl_xor_sync(SNB_FULL_MASK, lmin, o));
      if (lane == 0) s_fred[warp] = lmin;
      __syncthreads();
      float gmin = s_fred[0];
      for (int w = 1; w < kPitchThreads / 32; ++w) gmin = fminf(gmin, s_fred[w]);
      for (int i = tid; i < ns; i += kPitchThreads) s_prev[i] = __fadd_rn(s_new[i], -gmin);
      __syncthreads();
    }
    // ---- best final state (first minimum) and backtrace ----
    {
      float best = FLT_MAX;
      int bi = 0x7fffffff;
      for (int i = tid; i < ns; i += kPitchThreads)
        if (s_prev[i] < best) { best = s_prev[i]; bi = i; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(SNB_FULL_MASK, best, o);
        const int oi = __shfl_xor_sync(SNB_FULL_MASK, bi, o);
        if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (lane == 0) { s_fred[warp] = best; s_ired[warp] = bi; }
      __syncthreads();
      if (tid == 0) {
        float b = s_fred[0];
        int s = s_ired[0];
        for (int w = 1; w < kPitchThreads / 32; ++w)
          if (s_fred[w] < b || (s_fred[w] == b && s_ired[w] < s)) { b = s_fred[w]; s = s_ired[w]; }
        for (int64_t f = F - 1; f >= 0; --f) {
          states[f] = s;
          s = bp[f * ns + s];
        }
      }
      __syncthreads();
      __threadfence_block();
    }
    // ---- output rows: (NCCF_pov at the chosen lag, 1 / lag) ----
    for (int64_t f = tid; f < F; f += kPitchThreads) {
      const int s = states[f];
      const float *w = s_upw + s * a.up_nw_max;
      const float *pv = pov_raw + f * nm + s_upfirst[s];
      float acc = 0.0f;
      for (int j = 0; j < s_upn[s]; ++j) acc = fmaf(w[j], pv[j], acc);
      float *o = a.out + (row0 + f) * a.ld_out;
      o[0] = acc;
      o[1] = __fdiv_rn(1.0f, s_lags[s]);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// k2': warp-per-utterance tracker (default).  Same arithmetic as
// pitch_track_kernel, but each warp owns one utterance: no block barriers in
// the frame loop (only __syncwarp), up to 24 independent utterances per CTA
// (one CTA per SM, tables shared), utterances handed out through an atomic
// queue, window energies from a double prefix sum (one pass instead of one per
// lag), three NCCF lags per lane in flight, and a monotone divide-and-conquer
// Viterbi step.
//
// Viterbi step: bp(i) = first argmin_j pen[|i-j|] + prev[j] is non-decreasing
// in i (the penalty is convex), so bp(i) lies in [bp(i-s), bp(i+s)].
//   * anchors (multiples of kA1 and the last state): whole-warp scans, the
//     first and the last over every j, the others in bisection order over the
//     range left by their neighbours;
//   * levels s = kA1/2 ... 1: lane per state i = s(2k+1), serial scan of
//     [bp(i-s), bp(i+s)] -- 1-3 candidates on the plateaus of bp, 2s+1 where bp
//     follows i; the few states that straddle a jump of bp (range > 2s+2) are
//     handed to whole-warp scans instead of stalling their round.
// Every scan keeps the first minimum of its range, exactly like the
// brute-force reference step (oracle/kaldi_oracle.c, orc_compute_pitch).
// ---------------------------------------------------------------------------
constexpr int kTrackWarpsMax = 20;   // (24 fit in shared memory but run erratically slower)
#ifndef SNB_PITCH_A1LOG2
#define SNB_PITCH_A1LOG2 6
#endif
#ifndef SNB_PITCH_TMAX
#define SNB_PITCH_TMAX 32
#endif
constexpr int kA1Log2 = SNB_PITCH_A1LOG2, kA1 = 1 << kA1Log2;

struct WarpSmem {           // per-warp float offsets
  int win, pre, np, nv, prev, cost, abp, total;
};

// backpointers of the current frame are read with strides 2s: skew the index
__host__ __device__ inline int bp_slot(int i) { return i + (i >> 5); }

__host__ __device__ inline WarpSmem warp_smem_layout(int full_len, int nm, int ns, int nw_max) {
  WarpSmem w;
  int off = 0;
  w.win = off; off += (full_len + 4 + 3) / 4 * 4;             // + padding read by the NCCF loop
  w.pre = off; off += 2 * ((full_len + 1 + 1) / 2 * 2);      // doubles (as float pairs)
  w.np = off; off += (nm + nw_max + 3) / 4 * 4;               // + zero tail for the padded taps
  w.nv = off; off += (nm + 3) / 4 * 4;
  w.prev = off; off += (ns + 3) / 4 * 4;
  w.cost = off; off += (ns + 3) / 4 * 4;
  w.abp = off; off += (bp_slot(ns) + 4) / 4 * 4;              // int backpointers of this frame
  w.total = (off + 3) / 4 * 4;
  return w;
}

__host__ __device__ inline int warp_shared_floats(int ns, int nw_max) {
  return ((3 * ns + ns * (nw_max | 1) + ns) + 3) / 4 * 4;     // pen (mirrored), lags, up_w (odd row stride), up_first
}

// first argmin over j in [jlo, jhi] of pen[|i-j|] + prev[j], by the whole warp.
// Costs are non-negative floats (pen >= 0, prev >= 0): their bit patterns order
// like integers, so the warp minimum is one REDUX (then one more for the
// smallest j among the lanes that hold it: the first minimum).
// s_penc is the centre of the mirrored penalty table: s_penc[d] = pen[|d|].
__device__ __forceinline__ int coop_scan(const float *s_penc, const float *w_prev, int i, int jlo, int jhi,
                                         int lane) {
  int best = 0x7f7fffff;          // FLT_MAX
  int bj = 0x7fffffff;
  const float *pq = s_penc - i;
#pragma unroll 2
  for (int j = jlo + lane; j <= jhi; j += 32) {
    const int c = __float_as_int(__fadd_rn(pq[j], w_prev[j]));
    if (c < best) { best = c; bj = j; }
  }
  const int m = __reduce_min_sync(SNB_FULL_MASK, best);
  return __reduce_min_sync(SNB_FULL_MASK, best == m ? bj : 0x7fffffff);
}

__device__ __forceinline__ void warp_argmin(float &best, int &bj) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(SNB_FULL_MASK, best, o);
    const int oj = __shfl_xor_sync(SNB_FULL_MASK, bj, o);
    if (ob < best || (ob == best && oj < bj)) { best = ob; bj = oj; }
  }
}

template <int NWC>   // NWC > 0: number of upsampling taps known at compile time
__global__ void __launch_bounds__(kTrackWarpsMax * 32, 1) pitch_track_warp_kernel(const TrackArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarp_cta = blockDim.x >> 5;
  const int ns = a.nstates, nm = a.nmeas, bl = a.basic_len, fl = a.full_len;
  const int nw = NWC > 0 ? NWC : a.up_nw_max;
  const int nwp = nw | 1;                                     // odd row stride: conflict-free
  // ---- CTA-shared tables ----
  // penalty table mirrored around s_pen: s_pen[d] = pen[|d|], d in (-ns, ns):
  // the scans index it with the signed state distance, no absolute value
  float *s_pen = reinterpret_cast<float *>(smem_raw) + (ns - 1);
  float *s_lags = reinterpret_cast<float *>(smem_raw) + 2 * ns;
  float *s_upw = s_lags + ns;                                 // [ns][nwp]
  int32_t *s_upfirst = reinterpret_cast<int32_t *>(s_upw + ns * nwp);
  for (int i = tid; i < ns; i += blockDim.x) {
    s_pen[i] = a.pen[i]; s_pen[-i] = a.pen[i]; s_lags[i] = a.lags[i];
    s_upfirst[i] = min(max(a.up_first[i], 0), nm - 1);         // (a state without taps has zero weights)
  }
  for (int i = tid; i < ns * nw; i += blockDim.x) s_upw[(i / nw) * nwp + i % nw] = a.up_w[i];
  __syncthreads();
  // ---- warp-private buffers ----
  const WarpSmem L = warp_smem_layout(fl, nm, ns, nw);
  float *wbase = reinterpret_cast<float *>(smem_raw) + warp_shared_floats(ns, nw) + warp * L.total;
  float *w_win = wbase + L.win;
  double *w_pre = reinterpret_cast<double *>(wbase + L.pre);
  float *w_np = wbase + L.np, *w_nv = wbase + L.nv;
  float *w_prev = wbase + L.prev, *w_cost = wbase + L.cost;
  int *w_bp = reinterpret_cast<int *>(wbase + L.abp);

  const int64_t gw = static_cast<int64_t>(blockIdx.x) * nwarp_cta + warp;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * nwarp_cta;
  int16_t *bp = a.bp + gw * a.max_frames * ns;
  float *pov_raw = a.pov_raw + gw * a.max_frames * nm;
  int32_t *states = a.states + gw * a.max_frames;
  const int n1 = (ns - 1 + kA1 - 1) / kA1 + 1;                // anchors min(t kA1, ns-1), t < n1
  int top = 1;
  while (top < n1) top <<= 1;
  for (int i = nm + lane; i < nm + nw; i += 32) w_np[i] = 0.0f;
  if (lane < 4) w_win[fl + lane] = 0.0f;

  int64_t u = gw;
  while (u < a.nutts) {
    const int64_t doff = a.info[4 * u], m1 = a.info[4 * u + 1], m2 = a.info[4 * u + 2],
                  end1 = a.info[4 * u + 3];
    const int64_t row0 = a.frame_offsets[u], F = a.frame_offsets[u + 1] - row0;
    const float *x = a.down + doff;
    // ---- global mean-square for the ballast (double sums, two phases) ----
    double p1 = 0.0, q1 = 0.0, p2 = 0.0, q2 = 0.0;
    if (F > 0) {
      for (int64_t i = lane; i < m2; i += 32) {
        const double v = x[i];
        if (i < m1) { p1 += v; q1 += v * v; } else { p2 += v; q2 += v * v; }
      }
    }
    p1 = group_sum_f64<32>(p1); q1 = group_sum_f64<32>(q1);
    p2 = group_sum_f64<32>(p2); q2 = group_sum_f64<32>(q2);
    const double sum2 = p1 + p2, sq2 = q1 + q2;
    const double ms1 = m1 > 0 ? q1 / static_cast<double>(m1) - (p1 / static_cast<double>(m1)) * (p1 / static_cast<double>(m1)) : 0.0;
    const double ms2 = m2 > 0 ? sq2 / static_cast<double>(m2) - (sum2 / static_cast<double>(m2)) * (sum2 / static_cast<double>(m2)) : 0.0;
    const float ballast1 = static_cast<float>((ms1 * bl) * (ms1 * bl) * static_cast<double>(a.nccf_ballast));
    const float ballast2 = static_cast<float>((ms2 * bl) * (ms2 * bl) * static_cast<double>(a.nccf_ballast));
    for (int i = lane; i < ns; i += 32) w_prev[i] = 0.0f;
    __syncwarp();

    for (int64_t f = 0; f < F; ++f) {
      const bool phase2 = f >= end1;
      const int64_t avail = phase2 ? m2 : m1;
      const float ballast = phase2 ? ballast2 : ballast1;
      int64_t start;
      if (a.snip_edges) start = f * a.shift;
      else start = static_cast<int64_t>((static_cast<double>(f) + 0.5) * a.shift) - fl / 2;
      // ---- ExtractFrame ----
      for (int i = lane; i < fl; i += 32) {
        const int64_t k = start + i;
        w_win[i] = (k >= 0 && k < avail) ? x[k] : 0.0f;
      }
      __syncwarp();
      if (a.preemph != 0.0f) {
        for (int base = ((fl - 1) / 32) * 32; base >= 0; base -= 32) {   // downwards: original neighbours
          const int i = base + lane;
          float v = 0.0f;
          if (i < fl) v = (i > 0) ? fmaf(-a.preemph, w_win[i - 1], w_win[i]) : w_win[0] * (1.0f - a.preemph);
          __syncwarp();
          if (i < fl) w_win[i] = v;
          __syncwarp();
        }
      }
      // ---- mean of the first basic_len samples, subtracted from the whole window ----
      float sacc = 0.0f;
      for (int i = lane; i < bl; i += 32) sacc += w_win[i];
      const float mean = __fdiv_rn(group_sum<32>(sacc), static_cast<float>(bl));
      // ---- zero-mean window + double prefix sums of squares: pre[k] = sum_{i<k} z_i^2 ----
      {
        const int chunk = (fl + 31) / 32;
        const int i0 = lane * chunk, i1 = min(fl, i0 + chunk);
        double local = 0.0;
        for (int i = i0; i < i1; ++i) {
          const float z = w_win[i] - mean;
          w_win[i] = z;
          local += static_cast<double>(z) * z;
        }
        double incl = local;                       // inclusive scan over lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double up = __shfl_up_sync(SNB_FULL_MASK, incl, o);
          if (lane >= o) incl += up;
        }
        double run = incl - local;                 // exclusive
        for (int i = i0; i < i1; ++i) {
          w_pre[i] = run;
          run += static_cast<double>(w_win[i]) * w_win[i];
        }
        if (i1 == fl && i0 < fl) w_pre[fl] = run;
      }
      __syncwarp();
      const float e1 = static_cast<float>(w_pre[bl] - w_pre[0]);
      // ---- NCCF at the integer lags: lane = three consecutive lags; the window
      //      samples w[lag+i..] slide through six registers (4 new loads per 12 FMAs) ----
#pragma unroll 1
      for (int lb = 0; lb < nm; lb += 96) {
        const int l0 = lb + 3 * lane;
        const float *q = w_win + a.first_lag + min(l0, nm - 1);   // (reads up to 2 floats of padding)
        float in0 = 0.0f, in1 = 0.0f, in2 = 0.0f;
        float r0 = q[0], r1 = q[1];
        int i = 0;
#pragma unroll 1
        for (; i + 3 < bl; i += 4) {
          const float4 w4 = *reinterpret_cast<const float4 *>(w_win + i);
          const float r2 = q[i + 2], r3 = q[i + 3], r4 = q[i + 4], r5 = q[i + 5];
          in0 = fmaf(w4.x, r0, in0); in1 = fmaf(w4.x, r1, in1); in2 = fmaf(w4.x, r2, in2);
          in0 = fmaf(w4.y, r1, in0); in1 = fmaf(w4.y, r2, in1); in2 = fmaf(w4.y, r3, in2);
          in0 = fmaf(w4.z, r2, in0); in1 = fmaf(w4.z, r3, in1); in2 = fmaf(w4.z, r4, in2);
          in0 = fmaf(w4.w, r3, in0); in1 = fmaf(w4.w, r4, in1); in2 = fmaf(w4.w, r5, in2);
          r0 = r4; r1 = r5;
        }
        for (; i < bl; ++i) {
          const float wv = w_win[i];
          in0 = fmaf(wv, q[i], in0); in1 = fmaf(wv, q[i + 1], in1); in2 = fmaf(wv, q[i + 2], in2);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int l = l0 + k;
          const float inner = k == 0 ? in0 : (k == 1 ? in1 : in2);
          if (l < nm) {
            const int lag = a.first_lag + l;
            const float e2 = static_cast<float>(w_pre[lag + bl] - w_pre[lag]);
            const float norm = __fmul_rn(e1, e2);
            const float den_p = sqrtf(__fadd_rn(norm, ballast));
            const float den_v = sqrtf(norm);
            w_np[l] = den_p != 0.0f ? __fdiv_rn(inner, den_p) : 0.0f;
            const float pv = den_v != 0.0f ? __fdiv_rn(inner, den_v) : 0.0f;
            w_nv[l] = pv;
            pov_raw[f * nm + l] = pv;
          }
        }
      }
      __syncwarp();
      // ---- upsample to the log-spaced lags (taps padded with zeros to nw); local cost ----
      for (int i = lane; i < ns; i += 32) {
        const float *src = w_np + s_upfirst[i];
        const float *w = s_upw + i * nwp;
        float acc = 0.0f;
        if (NWC > 0) {
#pragma unroll
          for (int j = 0; j < NWC; ++j) acc = fmaf(w[j], src[j], acc);
        } else {
#pragma unroll 4
          for (int j = 0; j < nw; ++j) acc = fmaf(w[j], src[j], acc);
        }
        float c = __fadd_rn(1.0f, -acc);
        c = __fadd_rn(__fmul_rn(__fmul_rn(a.soft_min_f0, s_lags[i]), acc), c);
        w_cost[i] = c;
      }
      __syncwarp();
      // ---- Viterbi step: anchors by whole-warp scans, bisection order ----
      {
        const int last = ns - 1;
        int bj = coop_scan(s_pen, w_prev, 0, 0, last, lane);
        if (lane == 0) w_bp[bp_slot(0)] = bj;
        if (n1 > 1) {
          bj = coop_scan(s_pen, w_prev, last, bj, last, lane);
          if (lane == 0) w_bp[bp_slot(last)] = bj;
        }
        __syncwarp();
        for (int step = top >> 1; step >= 1; step >>= 1) {
          for (int t = step; t < n1 - 1; t += 2 * step) {
            const int i = t * kA1;
            const int il = (t - step) * kA1, ir = min((t + step) * kA1, last);
            const int b0 = w_bp[bp_slot(il)], b1 = w_bp[bp_slot(ir)];
            // (min/max: a float near-tie may break the monotonicity by one state)
            bj = coop_scan(s_pen, w_prev, i, min(b0, b1), max(b0, b1), lane);
            if (lane == 0) w_bp[bp_slot(i)] = bj;
          }
          __syncwarp();
        }
      }
      // ---- levels kA1/2 ... 1: lane per state, long ranges by the whole warp ----
#pragma unroll 1
      for (int ls = kA1Log2 - 1; ls >= 0; --ls) {
        const int s = 1 << ls;
        const int nodd = (ns - 2 >= s) ? ((ns - 2 - s) >> (ls + 1)) + 1 : 0;
        const int T = min(SNB_PITCH_TMAX, 2 * s + 2);
#pragma unroll 1
        for (int k0 = 0; k0 < nodd; k0 += 32) {
          const int k = k0 + lane;
          const bool act = k < nodd;
          const int i = s * (2 * k + 1);
          int jlo = 0, jhi = 0;
          if (act) {
            const int b0 = w_bp[bp_slot(i - s)], b1 = w_bp[bp_slot(min(i + s, ns - 1))];
            jlo = min(b0, b1); jhi = max(b0, b1);
          }
          unsigned longm = __ballot_sync(SNB_FULL_MASK, jhi - jlo >= T);
          int bj = jlo;
          if (jhi > jlo && jhi - jlo < T) {
            float best = FLT_MAX;
            const float *pp = w_prev + jlo, *pq = s_pen + (jlo - i);
            int bn = 0;
#pragma unroll 2
            for (int n = jhi - jlo; n >= 0; --n) {
              const float c = __fadd_rn(*pq, *pp);
              if (c < best) { best = c; bn = n; }
              ++pp; ++pq;
            }
            bj = jhi - bn;
          }
#pragma unroll 1
          while (longm) {
            const int src = __ffs(longm) - 1;
            longm &= longm - 1u;
            const int ci = __shfl_sync(SNB_FULL_MASK, i, src);
            const int clo = __shfl_sync(SNB_FULL_MASK, jlo, src);
            const int chi = __shfl_sync(SNB_FULL_MASK, jhi, src);
            const int cb = coop_scan(s_pen, w_prev, ci, clo, chi, lane);
            if (lane == src) bj = cb;
          }
          if (act) w_bp[bp_slot(i)] = bj;
        }
        __syncwarp();
      }
      // ---- forward costs, backpointers to global, renormalise ----
      float lmin = FLT_MAX;
      for (int i = lane; i < ns; i += 32) {
        const int bj = w_bp[bp_slot(i)];
        const float v = __fadd_rn(__fadd_rn(s_pen[bj - i], w_prev[bj]), w_cost[i]);
        bp[f * ns + i] = static_cast<int16_t>(bj);
        w_cost[i] = v;
        lmin = fminf(lmin, v);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lmin = fminf(lmin, __shfl_xor_sync(SNB_FULL_MASK, lmin, o));
      __syncwarp();
      for (int i = lane; i < ns; i += 32) w_prev[i] = __fadd_rn(w_cost[i], -lmin);
      __syncwarp();
    }
    if (F > 0) {
      // ---- best final state (first minimum), backtrace, output rows ----
      float best = FLT_MAX;
      int bi = 0x7fffffff;
      for (int i = lane; i < ns; i += 32)
        if (w_prev[i] < best) { best = w_prev[i]; bi = i; }
      warp_argmin(best, bi);
      if (lane == 0) {
        int sidx = bi;
        for (int64_t f = F - 1; f >= 0; --f) {
          states[f] = sidx;
          sidx = bp[f * ns + sidx];
        }
      }
      __syncwarp();
      __threadfence_block();
      for (int64_t f = lane; f < F; f += 32) {
        const int sidx = states[f];
        const float *w = s_upw + sidx * nwp;
        const float *pv = pov_raw + f * nm + s_upfirst[sidx];
        const int n = min(nw, nm - s_upfirst[sidx]);           // taps beyond are zero weights
        float acc = 0.0f;
#pragma unroll 1
        for (int j = 0; j < n; ++j) acc = fmaf(w[j], pv[j], acc);
        float *o = a.out + (row0 + f) * a.ld_out;
        o[0] = acc;
        o[1] = __fdiv_rn(1.0f, s_lags[sidx]);
      }
      __syncwarp();
    }
    // ---- next utterance from the queue ----
    unsigned long long nxt = 0;
    if (lane == 0) nxt = atomicAdd(a.queue, 1ULL);
    nxt = __shfl_sync(SNB_FULL_MASK, nxt, 0);
    u = nwarps + static_cast<int64_t>(nxt);
  }
}

// ---------------------------------------------------------------------------
// k3: ProcessPitch (OnlineProcessPitch in offline use), delay == 0
// grid = (nutts, chunks); each CTA owns kPostRows rows of one utterance plus
// the halo the normalisation window needs
// ---------------------------------------------------------------------------
constexpr int kPostRows = 128;

struct PostArgs {
  snb_pitch_post_opts o;
  const float *raw;
  int64_t ld_raw;
  const int64_t *frame_offsets;
  uint64_t seed;
  float *out;
  int64_t ld_out;
  int32_t halo;        // max(left, right, delta_window)
};

__device__ __forceinline__ float nccf_to_pov(float n) {
  float nd = fabsf(n);
  if (nd > 1.0f) nd = 1.0f;
  const double x = static_cast<double>(nd);
  const float r = static_cast<float>(-5.2 + 5.4 * exp(7.5 * (x - 1.0)) + 4.8 * x - 2.0 * exp(-10.0 * x) +
                                     4.2 * exp(20.0 * (x - 1.0)));
  return static_cast<float>(1.0 / (1.0 + exp(-1.0 * static_cast<double>(r))));
}

__global__ void __launch_bounds__(256) process_pitch_kernel(const PostArgs a) {
  extern __shared__ float s_post[];
  const int64_t u = blockIdx.x;
  const int64_t first = a.frame_offsets[u], F = a.frame_offsets[u + 1] - first;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * kPostRows;
  if (r0 >= F) return;
  const int span = kPostRows + 2 * a.halo;
  float *s_logp = s_post, *s_pov = s_post + span;
  const int64_t lo = r0 - a.halo;
  for (int i = threadIdx.x; i < span; i += blockDim.x) {
    const int64_t t = lo + i;
    float lp = 0.0f, pv = 0.0f;
    if (t >= 0 && t < F) {
      const float *row = a.raw + (first + t) * a.ld_raw;
      lp = logf(row[1]);
      pv = nccf_to_pov(row[0]);
    }
    s_logp[i] = lp;
    s_pov[i] = pv;
  }
  __syncthreads();
  const snb_pitch_post_opts &o = a.o;
  for (int r = threadIdx.x; r < kPostRows; r += blockDim.x) {
    const int64_t t = r0 + r;
    if (t >= F) break;
    float *out = a.out + (first + t) * a.ld_out;
    int col = 0;
    const float *row = a.raw + (first + t) * a.ld_raw;
    if (o.add_pov_feature) {
      float n = row[0];
      n = fminf(1.0f, fmaxf(-1.0f, n));
      const float f = static_cast<float>(pow(1.0001 - static_cast<double>(n), 0.15) - 1.0);
      out[col++] = __fadd_rn(__fmul_rn(o.pov_scale, f), o.pov_offset);
    }
    if (o.add_normalized_log_pitch) {
      int64_t b = t - o.normalization_left_context, e = t + o.normalization_right_context + 1;
      if (b < 0) b = 0;
      if (e > F) e = F;
      double sp = 0.0, slp = 0.0;
      for (int64_t f = b; f < e; ++f) {
        const float pv = s_pov[f - lo], lp = s_logp[f - lo];
        sp += pv;
        slp += static_cast<double>(__fmul_rn(pv, lp));
      }
      const float avg = static_cast<float>(slp / sp);
      out[col++] = __fmul_rn(__fadd_rn(s_logp[t - lo], -avg), o.pitch_scale);
    }
    if (o.add_delta_pitch) {
      const int w = o.delta_window;
      float norm = 0.0f;
      for (int j = -w; j <= w; ++j) norm += static_cast<float>(j * j);
      const float inv = static_cast<float>(1.0 / static_cast<double>(norm));
      float d = 0.0f;
      for (int j = -w; j <= w; ++j) {
        int64_t tt = t + j;
        tt = tt < 0 ? 0 : (tt >= F ? F - 1 : tt);
        const float s = __fmul_rn(static_cast<float>(j), inv);
        if (s != 0.0f) d = fmaf(s, s_logp[tt - lo], d);
      }
      float noise = 0.0f;
      if (o.delta_pitch_noise_stddev != 0.0f) {
        float g0, g1;
        gauss_pair(a.seed, static_cast<uint64_t>(first + t), 0u, &g0, &g1);
        noise = g0 * o.delta_pitch_noise_stddev;
      }
      out[col++] = __fmul_rn(__fadd_rn(d, noise), o.delta_pitch_scale);
    }
    if (o.add_raw_log_pitch) out[col++] = s_logp[t - lo];
  }
}

static size_t track_smem(const PitchTables *t);

static size_t warp_track_smem(const PitchTables *t, int warps) {
  const WarpSmem L = warp_smem_layout(t->full_len, t->nmeas, t->nstates, t->up_nw_max);
  return (static_cast<size_t>(warp_shared_floats(t->nstates, t->up_nw_max)) +
          static_cast<size_t>(warps) * L.total) * 4 + 16;
}

constexpr size_t kTrackSmemBudget = 224 * 1024;
constexpr int kTrackWarpsMin = 4;

// most warps (utterances in flight) one CTA can hold
static int warp_track_capacity(const PitchTables *t) {
  static const char *env = getenv("SNB_PITCH_WARPS");          // tuning knob: cap the warps per CTA
  int w = kTrackWarpsMax;
  if (env && atoi(env) >= kTrackWarpsMin) w = std::min(w, atoi(env) / 4 * 4);
  while (w >= kTrackWarpsMin && warp_track_smem(t, w) > kTrackSmemBudget) w -= 4;
  return w;
}

static bool use_warp_tracker(const PitchTables *t) {
  static const bool disabled = getenv("SNB_PITCH_CTA") != nullptr;
  return !disabled && warp_track_capacity(t) >= kTrackWarpsMin && t->nstates <= 32767;
}

// warp tracker launch shape: one CTA per SM, as many warps per CTA as the batch
// can feed (multiple of 4, up to the smem capacity); returns the number of
// concurrently tracked utterances ("slots" of per-utterance scratch)
static int64_t pitch_slots(const PitchTables *t, int64_t nutts, int *grid_out, int *warps_out) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
    cudaGetLastError();
    sms = 148;
  }
  const int cap = warp_track_capacity(t);
  int64_t w = (std::max<int64_t>(nutts, 1) + sms - 1) / sms;
  w = (w + 3) / 4 * 4;
  const int warps = static_cast<int>(std::max<int64_t>(kTrackWarpsMin, std::min<int64_t>(w, cap)));
  const int64_t want = (std::max<int64_t>(nutts, 1) + warps - 1) / warps;
  const int grid = static_cast<int>(std::min<int64_t>(want, sms));
  if (grid_out) *grid_out = grid;
  if (warps_out) *warps_out = warps;
  return static_cast<int64_t>(grid) * warps;
}

// persistent CTAs: exactly what is resident at once (a larger grid would run a
// second, underfilled wave)
static int pitch_grid(const PitchTables *t, int64_t nutts) {
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
    cudaGetLastError();
    sms = 148;
  }
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pitch_track_kernel, kPitchThreads,
                                                    track_smem(t)) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 1;
  }
  return static_cast<int>(std::min<int64_t>(nutts, static_cast<int64_t>(sms) * per_sm));
}

static size_t track_smem(const PitchTables *t) {
  return static_cast<size_t>(t->full_len + 2 * t->nmeas + 4 * t->nstates + t->nstates * t->up_nw_max) * 4 +
         static_cast<size_t>(2 * t->nstates) * 4 + 16;
}

}  // namespace snb

using namespace snb;

extern "C" int64_t snb_pitch_num_frames(int64_t nsamples, const snb_pitch_opts *po) {
  if (!pitch_opts_valid(*po)) return 0;
  int32_t first, last;
  lag_range(*po, &first, &last);
  return frames_available(resample_num_out(nsamples, *po, true), *po, last, true);
}

extern "C" int snb_pitch_plan_create(const snb_pitch_opts *po, snb_plan **out) {
  if (!po || !out) return set_error(SNB_ERR_VALUE, "null argument");
  *out = nullptr;
  snb_plan *plan = new snb_plan();
  plan->kind = 1;
  plan->po = *po;
  cudaGetDevice(&plan->device);
  cudaGetLastError();
  int rc = pitch_plan_init(plan);
  if (rc != SNB_OK) {
    pitch_plan_free(plan);
    delete plan;
    return rc;
  }
  *out = plan;
  return SNB_OK;
}

static int64_t max_frames_of(const snb_batch *b) {
  int64_t m = 0;
  for (int64_t u = 0; u < b->nutts; ++u) m = std::max(m, b->frame_offsets[u + 1] - b->frame_offsets[u]);
  return m;
}

static size_t align256(size_t v) { return (v + 255) / 256 * 256; }

extern "C" int64_t snb_pitch_workspace_bytes(const snb_plan *plan, const snb_batch *batch) {
  if (!plan || plan->kind != 1 || !batch) return -1;
  const PitchTables *t = plan->pitch;
  const int64_t grid = use_warp_tracker(t) ? pitch_slots(t, batch->nutts, nullptr, nullptr)
                                           : pitch_grid(t, batch->nutts);
  const int64_t mf = std::max<int64_t>(1, max_frames_of(batch));
  size_t bytes = align256(static_cast<size_t>(batch->total_down + 8) * 4) + 256;   // + utterance queue
  bytes += align256(static_cast<size_t>(grid) * mf * t->nstates * 2);
  bytes += align256(static_cast<size_t>(grid) * mf * t->nmeas * 4);
  bytes += align256(static_cast<size_t>(grid) * mf * 4);
  return static_cast<int64_t>(bytes);
}

extern "C" int snb_compute_pitch(const snb_plan *plan, const snb_batch *batch, const int16_t *d_pcm,
                                 void *d_workspace, int64_t workspace_bytes, float *d_out, int64_t ld_out,
                                 void *stream_) {
  if (!plan || plan->kind != 1 || !batch || batch->plan != plan)
    return set_error(SNB_ERR_VALUE, "plan/batch mismatch");
  if (batch->total_frames == 0) return SNB_OK;
  if (!d_pcm || !d_out || !d_workspace || ld_out < 2) return set_error(SNB_ERR_VALUE, "bad argument");
  if (workspace_bytes < snb_pitch_workspace_bytes(plan, batch))
    return set_error(SNB_ERR_VALUE, "pitch workspace too small");
  const PitchTables *t = plan->pitch;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  batch->note_stream(stream);
  const bool warp_path = use_warp_tracker(t);
  int warp_grid = 1, warp_count = kTrackWarpsMin;
  const int64_t grid = warp_path ? pitch_slots(t, batch->nutts, &warp_grid, &warp_count)
                                 : pitch_grid(t, batch->nutts);
  const int64_t mf = std::max<int64_t>(1, max_frames_of(batch));
  unsigned char *ws = static_cast<unsigned char *>(d_workspace);
  float *down = reinterpret_cast<float *>(ws);
  ws += align256(static_cast<size_t>(batch->total_down + 8) * 4);
  unsigned long long *queue = reinterpret_cast<unsigned long long *>(ws);
  ws += 256;
  int16_t *bp = reinterpret_cast<int16_t *>(ws);
  ws += align256(static_cast<size_t>(grid) * mf * t->nstates * 2);
  float *pov_raw = reinterpret_cast<float *>(ws);
  ws += align256(static_cast<size_t>(grid) * mf * t->nmeas * 4);
  int32_t *states = reinterpret_cast<int32_t *>(ws);

  ResampleArgs r;
  r.pcm = d_pcm;
  r.sample_begin = batch->d_sample_begin;
  r.sample_len = batch->d_sample_len;
  r.info = batch->d_down_offsets;
  r.nutts = batch->nutts;
  r.total_down = batch->total_down;
  r.in_unit = t->in_unit; r.out_unit = t->out_unit; r.nw_max = t->down_nw_max;
  r.first = t->d_down_first; r.nw = t->d_down_nw; r.w = t->d_down_w;
  r.down = down;
  if (batch->total_down > 0) {
    resample_kernel<<<static_cast<unsigned>((batch->total_down + 255) / 256), 256, 0, stream>>>(r);
    SNB_LAUNCH_CHECK();
  }
  TrackArgs a;
  a.down = down;
  a.info = batch->d_down_offsets;
  a.frame_offsets = batch->d_frame_offsets;
  a.nutts = batch->nutts;
  a.first_lag = t->first_lag; a.nmeas = t->nmeas; a.nstates = t->nstates;
  a.shift = t->shift; a.basic_len = t->basic_len; a.full_len = t->full_len; a.up_nw_max = t->up_nw_max;
  a.snip_edges = plan->po.snip_edges;
  a.preemph = plan->po.preemph_coeff;
  a.soft_min_f0 = plan->po.soft_min_f0;
  a.nccf_ballast = plan->po.nccf_ballast;
  a.lags = t->d_lags; a.pen = t->d_pen; a.up_w = t->d_up_w;
  a.up_first = t->d_up_first; a.up_nw = t->d_up_nw;
  a.bp = bp; a.pov_raw = pov_raw; a.states = states;
  a.max_frames = mf;
  a.out = d_out; a.ld_out = ld_out;
  a.queue = queue;
  if (warp_path) {
    const size_t wsmem = warp_track_smem(t, warp_count);
    static std::atomic<size_t> wcur{48 * 1024};
    size_t c = wcur.load();
    while (wsmem > c) {
      cudaError_t e = cudaFuncSetAttribute(pitch_track_warp_kernel<10>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(wsmem));
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(pitch_track_warp_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(wsmem));
      if (e != cudaSuccess) return set_error(SNB_ERR_CUDA, "pitch smem: %s", cudaGetErrorString(e));
      if (wcur.compare_exchange_weak(c, wsmem)) break;
    }
    cudaError_t e = cudaMemsetAsync(queue, 0, sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return set_error(SNB_ERR_CUDA, "pitch queue: %s", cudaGetErrorString(e));
    if (t->up_nw_max == 10)     // Kaldi's default resampling options
      pitch_track_warp_kernel<10><<<static_cast<unsigned>(warp_grid), warp_count * 32, wsmem, stream>>>(a);
    else
      pitch_track_warp_kernel<0><<<static_cast<unsigned>(warp_grid), warp_count * 32, wsmem, stream>>>(a);
    SNB_LAUNCH_CHECK();
    return SNB_OK;
  }
  const size_t smem = track_smem(t);
  {
    static std::atomic<size_t> cur{48 * 1024};
    size_t c = cur.load();
    while (smem > c) {
      cudaError_t e = cudaFuncSetAttribute(pitch_track_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem));
      if (e != cudaSuccess) return set_error(SNB_ERR_CUDA, "pitch smem: %s", cudaGetErrorString(e));
      if (cur.compare_exchange_weak(c, smem)) break;
    }
  }
  pitch_track_kernel<<<static_cast<unsigned>(grid), kPitchThreads, smem, stream>>>(a);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int32_t snb_process_pitch_dim(const snb_pitch_post_opts *o) {
  return (o->add_pov_feature != 0) + (o->add_normalized_log_pitch != 0) + (o->add_delta_pitch != 0) +
         (o->add_raw_log_pitch != 0);
}

extern "C" int snb_process_pitch(const snb_pitch_post_opts *o, const float *d_raw, int64_t ld_raw,
                                 const int64_t *d_frame_offsets, int64_t nutts, int64_t total_frames,
                                 int64_t max_frames, uint64_t seed, float *d_out, int64_t ld_out,
                                 void *stream) {
  if (!o) return set_error(SNB_ERR_VALUE, "null options");
  const int dim = snb_process_pitch_dim(o);
  if (dim == 0)
    return set_error(SNB_ERR_VALUE, "at least one of the pitch features must be enabled");
  if (o->delay != 0) return set_error(SNB_ERR_UNSUPPORTED, "delay != 0 is not supported on the GPU path");
  if (total_frames == 0 || nutts == 0) return SNB_OK;
  if (!d_raw || !d_out || ld_raw < 2 || ld_out < dim) return set_error(SNB_ERR_VALUE, "bad argument");
  if (o->normalization_left_context < 0 || o->normalization_right_context < 0 || o->delta_window <= 0)
    return set_error(SNB_ERR_VALUE, "invalid pitch post-processing contexts");
  PostArgs a;
  a.o = *o;
  a.raw = d_raw; a.ld_raw = ld_raw;
  a.frame_offsets = d_frame_offsets;
  a.seed = seed;
  a.out = d_out; a.ld_out = ld_out;
  a.halo = std::max(std::max(o->normalization_left_context, o->normalization_right_context), o->delta_window);
  const size_t smem = static_cast<size_t>(kPostRows + 2 * a.halo) * 2 * 4;
  if (smem > 200 * 1024) return set_error(SNB_ERR_UNSUPPORTED, "normalisation context too large");
  {
    static std::atomic<size_t> cur{48 * 1024};
    size_t c = cur.load();
    while (smem > c) {
      cudaError_t e = cudaFuncSetAttribute(process_pitch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem));
      if (e != cudaSuccess) return set_error(SNB_ERR_CUDA, "post smem: %s", cudaGetErrorString(e));
      if (cur.compare_exchange_weak(c, smem)) break;
    }
  }
  const unsigned chunks = static_cast<unsigned>((std::max<int64_t>(max_frames, 1) + kPostRows - 1) / kPostRows);
  if (chunks > 65535) return set_error(SNB_ERR_UNSUPPORTED, "utterance too long for process_pitch");
  process_pitch_kernel<<<dim3(static_cast<unsigned>(nutts), chunks), 256, smem, static_cast<cudaStream_t>(stream)>>>(a);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
