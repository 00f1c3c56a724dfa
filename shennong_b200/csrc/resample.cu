// resample.cu -- sample-rate conversion of a ragged batch on the device.
//
// SURVEY 8(f)3, the step before the path: the reference converts a signal with
// sox, or with scipy's FFT method when sox is missing
// (shennong/audio.py:358-423), one utterance at a time on the host.  Here the
// int16 payloads are already on the device for the feature kernels; the
// conversion is Kaldi's LinearResample (resample.cc, what
// kaldi::ResampleWaveform runs: Hann-windowed sinc low-pass, cutoff 0.99 x the
// lower Nyquist frequency, 6 zero crossings, flushed at the end of the signal)
// -- the resampler the pitch extractor of this path already uses
// (pitch.cu:resample_kernel), as a polyphase filter: rate_out / gcd phases,
// one weight row per phase.
//
//   snb_resampler_create / destroy    phase tables for a (rate_in, rate_out)
//   snb_resampler_num_out             samples an utterance of n gives
//   snb_resample_batch                every utterance of a packed batch
//
// Arithmetic: float weights x int16 samples accumulated in double in tap
// order, rounded to float once -- the order of the CPU oracle
// (oracle/kaldi_oracle.c:linear_resample), so the float result is
// bit-identical to it.  The int16 result truncates toward zero like the
// reference's `.astype(np.int16)` (audio.py:423), saturated.
#include <cmath>
#include <vector>

#include "snb_internal.h"

struct snb_resampler {
  int32_t rate_in = 0, rate_out = 0, in_unit = 0, out_unit = 0, nw_max = 0;
  int device = 0;
  void *d_blob = nullptr;
  const int32_t *d_first = nullptr, *d_nw = nullptr;
  const float *d_w = nullptr;
};

namespace snb {

static int32_t gcd32(int32_t a, int32_t b) {
  while (b) { const int32_t t = a % b; a = b; b = t; }
  return a;
}

// LinearResample::FilterFunc (BaseFloat argument, double math inside)
static float filter_func(float t, float cutoff, int32_t num_zeros) {
  const double two_pi = 6.283185307179586476925286766559005, pi = 3.1415926535897932384626433832795;
  float window = 0.0f, filter;
  if (std::fabs(static_cast<double>(t)) < num_zeros / (2.0 * cutoff))
    window = static_cast<float>(0.5 * (1 + std::cos(two_pi * cutoff / num_zeros * t)));
  if (t != 0.0f) filter = static_cast<float>(std::sin(two_pi * cutoff * t) / (pi * t));
  else filter = static_cast<float>(2.0 * cutoff);
  return filter * window;
}

struct ResampleBatchArgs {
  const int16_t *pcm;
  const int64_t *begin, *len, *out_begin;
  int32_t in_unit, out_unit, nw_max;
  const int32_t *first, *nw;
  const float *w;
  float *out_f32;
  int16_t *out_i16;
};

// grid = (pieces of 256 output samples, utterances); the weight rows of the
// out_unit phases are read through L1 (a few KB, shared by every thread)
__global__ void __launch_bounds__(256) resample_batch_kernel(const ResampleBatchArgs a, const int64_t *n_out) {
  const int64_t u = blockIdx.y;
  const int64_t count = n_out[u];
  const int64_t in0 = a.begin[u], n_in = a.len[u], out0 = a.out_begin[u];
  for (int64_t so = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; so < count;
       so += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t unit = so / a.out_unit;
    const int32_t wrapped = static_cast<int32_t>(so - unit * a.out_unit);
    const int64_t first_in = a.first[wrapped] + unit * a.in_unit;
    const float *w = a.w + static_cast<int64_t>(wrapped) * a.nw_max;
    const int32_t nw = a.nw[wrapped];
    double acc = 0.0;
    for (int32_t j = 0; j < nw; ++j) {
      const int64_t k = first_in + j;
      if (k >= 0 && k < n_in)
        acc = fma(static_cast<double>(__ldg(w + j)), static_cast<double>(a.pcm[in0 + k]), acc);
    }
    const float v = static_cast<float>(acc);
    if (a.out_f32) a.out_f32[out0 + so] = v;
    if (a.out_i16) {
      const float c = fminf(fmaxf(truncf(v), -32768.0f), 32767.0f);
      a.out_i16[out0 + so] = static_cast<int16_t>(c);
    }
  }
}

// number of outputs of a flushed LinearResample (GetNumOutputSamples)
static int64_t num_out(int64_t n_in, int32_t rate_in, int32_t rate_out) {
  const int32_t base = gcd32(rate_in, rate_out);
  const int64_t tick_freq = static_cast<int64_t>(rate_in) / base * rate_out;
  const int64_t interval = n_in * (tick_freq / rate_in);
  if (interval <= 0) return 0;
  const int64_t per_out = tick_freq / rate_out;
  int64_t last = interval / per_out;
  if (last * per_out == interval) --last;
  return last + 1;
}

__global__ void resample_counts_kernel(const int64_t *len, int64_t nutts, int32_t in_unit, int32_t out_unit,
                                       int64_t *n_out) {
  const int64_t u = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (u >= nutts) return;
  // interval in ticks = len * out_unit, one output every in_unit ticks
  const int64_t interval = len[u] * out_unit;
  int64_t count = 0;
  if (interval > 0) {
    int64_t last = interval / in_unit;
    if (last * in_unit == interval) --last;
    count = last + 1;
  }
  n_out[u] = count;
}

}  // namespace snb

using namespace snb;

extern "C" int snb_resampler_create(int32_t rate_in, int32_t rate_out, float lowpass_cutoff, int32_t num_zeros,
                                    snb_resampler **out) {
  if (!out) return set_error(SNB_ERR_VALUE, "bad argument");
  *out = nullptr;
  if (rate_in <= 0 || rate_out <= 0) return set_error(SNB_ERR_VALUE, "sample rates must be positive");
  if (lowpass_cutoff <= 0.0f) lowpass_cutoff = 0.99f * 0.5f * static_cast<float>(std::min(rate_in, rate_out));
  if (num_zeros <= 0) num_zeros = 6;
  if (lowpass_cutoff * 2 > static_cast<float>(std::min(rate_in, rate_out)))
    return set_error(SNB_ERR_OPTION, "lowpass cutoff above the Nyquist frequency");
  snb_resampler *r = new snb_resampler();
  r->rate_in = rate_in; r->rate_out = rate_out;
  const int32_t base = gcd32(rate_in, rate_out);
  r->in_unit = rate_in / base; r->out_unit = rate_out / base;
  // LinearResample::SetIndexesAndWeights
  const double window_width = num_zeros / (2.0 * lowpass_cutoff);
  std::vector<int32_t> first(r->out_unit), nw(r->out_unit);
  std::vector<std::vector<float>> rows(r->out_unit);
  for (int32_t i = 0; i < r->out_unit; ++i) {
    const double output_t = i / static_cast<double>(rate_out);
    const int32_t min_idx = static_cast<int32_t>(std::ceil((output_t - window_width) * rate_in));
    const int32_t max_idx = static_cast<int32_t>(std::floor((output_t + window_width) * rate_in));
    first[i] = min_idx; nw[i] = max_idx - min_idx + 1;
    for (int32_t j = 0; j < nw[i]; ++j) {
      const double delta_t = (min_idx + j) / static_cast<double>(rate_in) - output_t;
      rows[i].push_back(filter_func(static_cast<float>(delta_t), lowpass_cutoff, num_zeros) /
                        static_cast<float>(rate_in));
    }
    r->nw_max = std::max(r->nw_max, nw[i]);
  }
  const size_t words = 2 * static_cast<size_t>(r->out_unit) + static_cast<size_t>(r->out_unit) * r->nw_max;
  std::vector<int32_t> blob(words, 0);
  std::copy(first.begin(), first.end(), blob.begin());
  std::copy(nw.begin(), nw.end(), blob.begin() + r->out_unit);
  float *wdst = reinterpret_cast<float *>(blob.data() + 2 * r->out_unit);
  for (int32_t i = 0; i < r->out_unit; ++i)
    std::copy(rows[i].begin(), rows[i].end(), wdst + static_cast<size_t>(i) * r->nw_max);
  cudaError_t e = cudaGetDevice(&r->device);
  if (e == cudaSuccess) e = cudaMalloc(&r->d_blob, words * 4);
  if (e == cudaSuccess) e = upload(r->d_blob, blob.data(), words * 4);
  if (e != cudaSuccess) {
    if (r->d_blob) cudaFree(r->d_blob);
    delete r;
    return set_error(SNB_ERR_CUDA, "resampler tables: %s", cudaGetErrorString(e));
  }
  const int32_t *d = static_cast<const int32_t *>(r->d_blob);
  r->d_first = d; r->d_nw = d + r->out_unit;
  r->d_w = reinterpret_cast<const float *>(d + 2 * r->out_unit);
  *out = r;
  return SNB_OK;
}

extern "C" void snb_resampler_destroy(snb_resampler *r) {
  if (!r) return;
  if (r->d_blob) cudaFree(r->d_blob);
  delete r;
}

extern "C" int64_t snb_resampler_num_out(const snb_resampler *r, int64_t nsamples) {
  if (!r || nsamples <= 0) return 0;
  return num_out(nsamples, r->rate_in, r->rate_out);
}

extern "C" int snb_resample_batch(const snb_resampler *r, const int16_t *d_pcm, const int64_t *d_begin,
                                  const int64_t *d_len, const int64_t *d_out_begin, int64_t nutts,
                                  int64_t max_out, float *d_out_f32, int16_t *d_out_i16, int64_t *d_counts,
                                  void *stream) {
  if (!r || nutts < 0) return set_error(SNB_ERR_VALUE, "bad argument");
  if (nutts == 0 || max_out <= 0) return SNB_OK;
  if (!d_pcm || !d_begin || !d_len || !d_out_begin || !d_counts || (!d_out_f32 && !d_out_i16))
    return set_error(SNB_ERR_VALUE, "bad argument");
  if (nutts > 65535) return set_error(SNB_ERR_UNSUPPORTED, "at most 65535 utterances per call");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  resample_counts_kernel<<<static_cast<unsigned>((nutts + 255) / 256), 256, 0, s>>>(d_len, nutts, r->in_unit,
                                                                                      r->out_unit, d_counts);
  SNB_LAUNCH_CHECK();
  ResampleBatchArgs a;
  a.pcm = d_pcm; a.begin = d_begin; a.len = d_len; a.out_begin = d_out_begin;
  a.in_unit = r->in_unit; a.out_unit = r->out_unit; a.nw_max = r->nw_max;
  a.first = r->d_first; a.nw = r->d_nw; a.w = r->d_w;
  a.out_f32 = d_out_f32; a.out_i16 = d_out_i16;
  const unsigned gx = static_cast<unsigned>(std::min<int64_t>((max_out + 255) / 256, 1 << 20));
  resample_batch_kernel<<<dim3(gx, static_cast<unsigned>(nutts)), 256, 0, s>>>(a, d_counts);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
