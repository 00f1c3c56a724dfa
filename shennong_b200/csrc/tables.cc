// tables.cc -- host-side construction of the immutable tables a plan uploads
// once: framing arithmetic, window, mel banks (with VTLN warp), DCT, lifter,
// equal-loudness and IDFT bases.  Semantics: Kaldi src/feat as reached from
// shennong (frames.py:137, window.py:107-114, processor/base.py:308,
// processor/plp.py:468-506); float32 storage with double intermediates
// exactly where Kaldi has them.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "snb_internal.h"

namespace snb {

static thread_local char g_error[512] = "";
std::atomic<int64_t> g_launch_count{0};

int set_error(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}
const char *last_error() { return g_error; }

cudaError_t upload_stream(cudaStream_t *out) {
  static thread_local cudaStream_t stream = nullptr;
  static thread_local int stream_device = -1;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (stream == nullptr || stream_device != dev) {
    e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return e;
    stream_device = dev;
  }
  *out = stream;
  return cudaSuccess;
}

cudaError_t upload(void *dst, const void *src, size_t bytes) {
  if (bytes == 0) return cudaSuccess;
  cudaStream_t stream = nullptr;
  cudaError_t e = upload_stream(&stream);
  if (e != cudaSuccess) return e;
  e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(stream);
}

// ---- MemPool ----------------------------------------------------------------
bool MemPool::ready(PoolBlock *b) {
  for (cudaEvent_t ev : b->pending) {
    const cudaError_t e = cudaEventQuery(ev);
    if (e == cudaErrorNotReady) {
      cudaGetLastError();
      return false;
    }
    if (e != cudaSuccess) cudaGetLastError();     // broken event: nothing can still be queued on it
  }
  for (cudaEvent_t ev : b->pending) cudaEventDestroy(ev);
  b->pending.clear();
  return true;
}

void MemPool::free_block(PoolBlock *b) {
  for (cudaEvent_t ev : b->pending) {
    cudaEventSynchronize(ev);
    cudaEventDestroy(ev);
  }
  b->pending.clear();
  if (b->p) {
    if (pinned_) cudaFreeHost(b->p); else cudaFree(b->p);
  }
  cudaGetLastError();
  b->p = nullptr;
}

cudaError_t MemPool::acquire(size_t bytes, int device, PoolBlock *out) {
  const size_t want = std::max<size_t>(bytes, 1);
  {
    std::lock_guard<std::mutex> lock(mu_);
    int best = -1;
    for (size_t i = 0; i < free_.size(); ++i) {
      PoolBlock &b = free_[i];
      if (b.device != device || b.cap < want || b.cap > std::max<size_t>(4 * want, 1 << 20)) continue;
      if (best >= 0 && free_[best].cap <= b.cap) continue;
      if (!ready(&b)) continue;
      best = static_cast<int>(i);
    }
    if (best >= 0) {
      *out = free_[best];
      free_.erase(free_.begin() + best);
      return cudaSuccess;
    }
  }
  PoolBlock b;
  b.device = device;
  b.cap = (want + 65535) / 65536 * 65536;
  const cudaError_t e = pinned_ ? cudaHostAlloc(&b.p, b.cap, cudaHostAllocDefault) : cudaMalloc(&b.p, b.cap);
  if (e != cudaSuccess) return e;
  *out = b;
  return cudaSuccess;
}

void MemPool::release(PoolBlock block) {
  if (!block.p) return;
  constexpr size_t kMaxFree = 64;
  std::lock_guard<std::mutex> lock(mu_);
  free_.push_back(block);
  if (free_.size() > kMaxFree) {                  // trim: oldest block whose work has drained
    for (size_t i = 0; i < free_.size(); ++i) {
      if (ready(&free_[i])) {
        free_block(&free_[i]);
        free_.erase(free_.begin() + i);
        break;
      }
    }
  }
}

MemPool &device_pool() { static MemPool *p = new MemPool(false); return *p; }
MemPool &pinned_pool() { static MemPool *p = new MemPool(true); return *p; }

int32_t window_shift(const snb_frame_opts &o) {
  return static_cast<int32_t>(static_cast<double>(o.samp_freq) * 0.001 *
                              static_cast<double>(o.frame_shift_ms));
}
int32_t window_size(const snb_frame_opts &o) {
  return static_cast<int32_t>(static_cast<double>(o.samp_freq) * 0.001 *
                              static_cast<double>(o.frame_length_ms));
}
int32_t padded_window_size(const snb_frame_opts &o) {
  int32_t w = window_size(o);
  if (!o.round_to_power_of_two) return w;
  int32_t n = 1;
  while (n < w) n *= 2;
  return n;
}

int64_t num_frames(int64_t nsamples, const snb_frame_opts &o) {
  const int64_t shift = window_shift(o), len = window_size(o);
  if (shift <= 0 || len <= 0) return -1;
  if (o.snip_edges) return nsamples < len ? 0 : 1 + (nsamples - len) / shift;
  return (nsamples + shift / 2) / shift;
}

int64_t first_sample_of_frame(int64_t frame, const snb_frame_opts &o) {
  const int64_t shift = window_shift(o);
  if (o.snip_edges) return frame * shift;
  return shift * frame + shift / 2 - window_size(o) / 2;
}

void window_function(const snb_frame_opts &o, std::vector<float> *out) {
  const int32_t n = window_size(o);
  out->assign(n > 0 ? n : 0, 0.0f);
  const double two_pi = 6.283185307179586476925286766559005;
  const double a = two_pi / (n - 1);
  for (int32_t i = 0; i < n; ++i) {
    const double c = std::cos(a * i);
    double w = 1.0;
    switch (o.window_type) {
      case SNB_WIN_HAMMING: w = 0.54 - 0.46 * c; break;
      case SNB_WIN_HANNING: w = 0.5 - 0.5 * c; break;
      case SNB_WIN_POVEY: w = std::pow(0.5 - 0.5 * c, 0.85); break;
      case SNB_WIN_BLACKMAN:
        w = o.blackman_coeff - 0.5 * c +
            (0.5 - o.blackman_coeff) * std::cos(2 * a * i);
        break;
      default: w = 1.0;
    }
    (*out)[i] = static_cast<float>(w);
  }
}

// ---- mel scale (float32 like Kaldi's MelScale / InverseMelScale) ----------
static inline float hz_to_mel(float f) {
  return 1127.0f * logf(1.0f + f / 700.0f);
}
static inline float mel_to_hz(float m) {
  return 700.0f * (expf(m / 1127.0f) - 1.0f);
}

// piece-wise linear VTLN warp of a frequency (VtlnWarpFreq)
static float warp_hz(float vtln_lo, float vtln_hi, float lo, float hi,
                     float warp, float f) {
  if (f < lo || f > hi) return f;
  const float l = vtln_lo * fmaxf(1.0f, warp);
  const float h = vtln_hi * fminf(1.0f, warp);
  const float scale = 1.0f / warp;
  const float Fl = scale * l, Fh = scale * h;
  const float left_slope = (Fl - lo) / (l - lo);
  const float right_slope = (hi - Fh) / (hi - h);
  if (f < l) return lo + left_slope * (f - lo);
  if (f < h) return scale * f;
  return hi + right_slope * (f - hi);
}

int build_mel_banks(const snb_frame_opts &fo, const snb_mel_opts &mo,
                    float vtln_warp, MelBanksHost *out) {
  const int32_t B = mo.num_bins;
  if (B < 3) return set_error(SNB_ERR_OPTION, "Must have at least 3 mel bins");
  const float sr = fo.samp_freq;
  const int32_t padded = padded_window_size(fo);
  if (padded % 2 != 0)
    return set_error(SNB_ERR_OPTION, "padded window size must be even");
  const int32_t nfft = padded / 2;
  const float nyquist = 0.5f * sr;
  const float lo = mo.low_freq;
  const float hi = mo.high_freq > 0.0f ? mo.high_freq : nyquist + mo.high_freq;
  if (lo < 0.0f || lo >= nyquist || hi <= 0.0f || hi > nyquist || hi <= lo)
    return set_error(SNB_ERR_OPTION,
                     "Bad values in options: low-freq %g and high-freq %g vs. "
                     "nyquist %g", lo, hi, nyquist);
  const float bin_width = sr / padded;
  const float mel_lo = hz_to_mel(lo), mel_hi = hz_to_mel(hi);
  const float delta = (mel_hi - mel_lo) / (B + 1);
  float vlo = mo.vtln_low, vhi = mo.vtln_high;
  if (vhi < 0.0f) vhi += nyquist;
  if (vtln_warp != 1.0f &&
      (vlo < 0.0f || vlo <= lo || vlo >= hi || vhi <= 0.0f || vhi >= hi ||
       vhi <= vlo))
    return set_error(SNB_ERR_OPTION,
                     "Bad values in options: vtln-low %g and vtln-high %g, "
                     "versus low-freq %g and high-freq %g", vlo, vhi, lo, hi);
  out->num_bins = B;
  out->num_fft_bins = nfft;
  out->first.assign(B, 0);
  out->size.assign(B, 0);
  out->offset.assign(B, 0);
  out->center_freqs.assign(B, 0.0f);
  out->weights.clear();
  out->seg_first.assign(B + 1, 0);
  out->seg_size.assign(B + 1, 0);
  out->up.assign(nfft, 0.0f);
  out->down.assign(nfft, 0.0f);
  std::vector<int32_t> &seg_of = out->seg_of;
  seg_of.assign(nfft, -1);
  std::vector<float> bin_mel(nfft);
  for (int32_t i = 0; i < nfft; ++i) bin_mel[i] = hz_to_mel(bin_width * i);
  for (int32_t b = 0; b < B; ++b) {
    float left = mel_lo + b * delta, center = mel_lo + (b + 1) * delta,
          right = mel_lo + (b + 2) * delta;
    if (vtln_warp != 1.0f) {
      left = hz_to_mel(warp_hz(vlo, vhi, lo, hi, vtln_warp, mel_to_hz(left)));
      center = hz_to_mel(warp_hz(vlo, vhi, lo, hi, vtln_warp, mel_to_hz(center)));
      right = hz_to_mel(warp_hz(vlo, vhi, lo, hi, vtln_warp, mel_to_hz(right)));
    }
    out->center_freqs[b] = mel_to_hz(center);
    int32_t first = -1, last = -1;
    for (int32_t i = 0; i < nfft; ++i)
      if (bin_mel[i] > left && bin_mel[i] < right) {
        if (first < 0) first = i;
        last = i;
      }
    if (first < 0)
      return set_error(SNB_ERR_OPTION,
                       "You may have set num_bins too large (empty mel bin %d)",
                       b);
    out->first[b] = first;
    out->size[b] = last + 1 - first;
    out->offset[b] = static_cast<int32_t>(out->weights.size());
    for (int32_t i = first; i <= last; ++i) {
      const float m = bin_mel[i];
      float w = 0.0f;
      if (m > left && m < right) {
        if (m <= center) {
          w = (m - left) / (center - left);
          out->up[i] = w;
          seg_of[i] = b;
        } else {
          w = (right - m) / (right - center);
          out->down[i] = w;
          seg_of[i] = b + 1;
        }
      }
      out->weights.push_back(w);
    }
  }
  // segments are contiguous runs of FFT bins (mel is increasing with i)
  for (int32_t s = 0; s <= B; ++s) {
    int32_t first = -1, last = -1;
    for (int32_t i = 0; i < nfft; ++i)
      if (seg_of[i] == s) {
        if (first < 0) first = i;
        last = i;
      }
    out->seg_first[s] = first < 0 ? 0 : first;
    out->seg_size[s] = first < 0 ? 0 : last + 1 - first;
    for (int32_t i = out->seg_first[s]; i < out->seg_first[s] + out->seg_size[s]; ++i)
      if (seg_of[i] != s)
        return set_error(SNB_ERR_UNSUPPORTED, "non monotonic mel segments");
  }
  return SNB_OK;
}

void build_dct(int32_t num_ceps, int32_t num_bins, std::vector<float> *out) {
  const double pi = 3.1415926535897932384626433832795;
  out->assign(static_cast<size_t>(num_ceps) * num_bins, 0.0f);
  const float n0 = static_cast<float>(std::sqrt(1.0 / static_cast<float>(num_bins)));
  const float nk = static_cast<float>(std::sqrt(2.0 / static_cast<float>(num_bins)));
  for (int32_t n = 0; n < num_bins; ++n) (*out)[n] = n0;
  for (int32_t k = 1; k < num_ceps; ++k)
    for (int32_t n = 0; n < num_bins; ++n)
      (*out)[static_cast<size_t>(k) * num_bins + n] = static_cast<float>(
          nk * std::cos(pi / num_bins * (n + 0.5) * k));
}

void build_lifter(int32_t num_ceps, float q, std::vector<float> *out) {
  const double pi = 3.1415926535897932384626433832795;
  out->assign(num_ceps, 1.0f);
  if (q == 0.0f) return;
  for (int32_t i = 0; i < num_ceps; ++i)
    (*out)[i] = static_cast<float>(1.0 + 0.5 * q * std::sin(pi * i / q));
}

void build_equal_loudness(const std::vector<float> &cf, std::vector<float> *out) {
  out->assign(cf.size(), 0.0f);
  for (size_t i = 0; i < cf.size(); ++i) {
    const float fsq = cf[i] * cf[i];
    const float fsub = fsq / (fsq + 1.6e5f);
    (*out)[i] = fsub * fsub * ((fsq + 1.44e6f) / (fsq + 9.61e6f));
  }
}

void build_idft_bases(int32_t n_bases, int32_t dimension, std::vector<float> *out) {
  const double pi = 3.1415926535897932384626433832795;
  out->assign(static_cast<size_t>(n_bases) * dimension, 0.0f);
  const float angle = static_cast<float>(pi / static_cast<float>(dimension - 1));
  const float scale = static_cast<float>(1.0f / (2.0 * static_cast<float>(dimension - 1)));
  for (int32_t i = 0; i < n_bases; ++i) {
    float *row = out->data() + static_cast<size_t>(i) * dimension;
    row[0] = static_cast<float>(1.0 * scale);
    const float fi = static_cast<float>(i);
    for (int32_t j = 1; j < dimension - 1; ++j)
      row[j] = static_cast<float>(2.0 * scale *
                                  std::cos(static_cast<double>(angle * fi * static_cast<float>(j))));
    row[dimension - 1] = static_cast<float>(
        scale * std::cos(static_cast<double>(angle * fi * static_cast<float>(dimension - 1))));
  }
}

}  // namespace snb

// ---- C ABI: host-only helpers ---------------------------------------------
namespace snb { const char *last_error(); }

extern "C" {

int snb_version(void) { return SNB_VERSION; }
const char *snb_last_error(void) { return snb::last_error(); }
int64_t snb_launch_count(void) { return snb::g_launch_count.load(); }

int32_t snb_window_size(const snb_frame_opts *o) { return snb::window_size(*o); }
int32_t snb_window_shift(const snb_frame_opts *o) { return snb::window_shift(*o); }
int32_t snb_padded_window_size(const snb_frame_opts *o) {
  return snb::padded_window_size(*o);
}
int64_t snb_num_frames(int64_t nsamples, const snb_frame_opts *o) {
  int64_t n = snb::num_frames(nsamples, *o);
  if (n < 0) snb::set_error(SNB_ERR_VALUE, "sample rate too low for the frame shift/length");
  return n;
}
int64_t snb_first_sample_of_frame(int32_t frame, const snb_frame_opts *o) {
  return snb::first_sample_of_frame(frame, *o);
}
int snb_window_function(const snb_frame_opts *o, float *out, int32_t capacity) {
  std::vector<float> w;
  snb::window_function(*o, &w);
  if (static_cast<int32_t>(w.size()) > capacity)
    return snb::set_error(SNB_ERR_VALUE, "window buffer too small: %d < %zu",
                          capacity, w.size());
  std::memcpy(out, w.data(), w.size() * sizeof(float));
  return SNB_OK;
}
int snb_mel_banks_host(const snb_frame_opts *fo, const snb_mel_opts *mo,
                       float vtln_warp, float *weights, float *center_freqs) {
  snb::MelBanksHost mb;
  int rc = snb::build_mel_banks(*fo, *mo, vtln_warp, &mb);
  if (rc != SNB_OK) return rc;
  std::memset(weights, 0, sizeof(float) * static_cast<size_t>(mb.num_bins) * mb.num_fft_bins);
  for (int32_t b = 0; b < mb.num_bins; ++b) {
    for (int32_t i = 0; i < mb.size[b]; ++i)
      weights[static_cast<size_t>(b) * mb.num_fft_bins + mb.first[b] + i] =
          mb.weights[mb.offset[b] + i];
    if (center_freqs) center_freqs[b] = mb.center_freqs[b];
  }
  return SNB_OK;
}

}  // extern "C"
