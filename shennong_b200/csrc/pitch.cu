// pitch.cu -- Kaldi pitch extraction and post-processing for ragged batches.
//
// Replaces kaldi.feat.pitch.compute_kaldi_pitch / process_pitch
// (shennong/processor/pitch_kaldi.py:296-299, 535-537), i.e. Kaldi's
// pitch-functions.cc + resample.cc in OFFLINE use: one AcceptWaveform() with
// the resampler unflushed followed by InputFinished().
//
//   k1  resample_kernel       16k -> 4k windowed-sinc (LinearResample), one
//                             thread per output sample, all utterances at once
//   k2  pitch_ballast_kernel  per-utterance mean square -> NCCF ballast
//   k3  pitch_nccf_kernel     frame-parallel: NCCF at the integer lags, sinc-
//                             upsampling to the log-spaced lags (Arbitrary-
//                             Resample), local cost of every Viterbi state
//   k4  pitch_viterbi_kernel  warp per utterance, sequential over frames: exact
//                             monotone Viterbi step, backpointers, backtrace,
//                             the (NCCF, pitch) rows
//   k5  process_pitch_kernel  POV / normalised log-pitch / delta / raw log-pitch
//
// Every dot product of k1/k3/k4 accumulates in double in index order and is
// rounded to float once, like the oracle (oracle/kaldi_oracle.c): the Viterbi
// state sequence is bit-identical to the oracle's (tests/test_gpu_pitch.py).
#include <cfloat>
#include <cmath>
#include <algorithm>
#include <cstring>

#include "device_utils.cuh"
#include "snb_internal.h"

namespace snb {


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
  int32_t ns4, nwp;                             // nstates rounded up to 4 (cost rows); up_nw_max | 1
  std::vector<double> up_w_t;                   // [nstates, nwp] same taps as doubles (odd row stride:
                                                // conflict-free 64-bit loads by consecutive states)
  // device copies (one allocation)
  void *d_blob = nullptr;
  const int32_t *d_down_first, *d_down_nw, *d_up_first, *d_up_nw;
  const float *d_down_w, *d_lags, *d_pen, *d_up_w;
  const double *d_up_w_t;
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
  if (t->nstates > 4096 || t->full_len > 4096 || t->nmeas > 1024)
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
  t->ns4 = (t->nstates + 3) & ~3;
  t->nwp = t->up_nw_max | 1;
  t->up_w_t.assign(static_cast<size_t>(t->nstates) * t->nwp, 0.0);
  for (int32_t i = 0; i < t->nstates; ++i)
    for (int32_t j = 0; j < t->up_nw[i]; ++j)
      t->up_w_t[static_cast<size_t>(i) * t->nwp + j] = static_cast<double>(uw[i][j]);
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
  const size_t o9 = blob.size();
  blob.resize(o9 + t->up_w_t.size() * 2);
  std::memcpy(blob.data() + o9, t->up_w_t.data(), t->up_w_t.size() * 8);
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
  t->d_up_w_t = reinterpret_cast<const double *>(d + o9);
  return SNB_OK;
}

void pitch_plan_free(snb_plan *plan) {
  if (!plan->pitch) return;
  if (plan->pitch->d_blob) cudaFree(plan->pitch->d_blob);
  delete plan->pitch;
  plan->pitch = nullptr;
}

// per-utterance phase info packed as int64 x 4: down_offset, m1, m2, end1,
// followed by order[nutts] (utterances by decreasing number of frames, stable:
// the identity for equal lengths) and gfo[nutts+1] (frames cumulated in that
// order): the frame-parallel kernel and the tracker walk the batch in this
// order so that the utterances of a tracker wave have similar lengths
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
  std::vector<int64_t> order(static_cast<size_t>(b->nutts));
  for (int64_t u = 0; u < b->nutts; ++u) order[u] = u;
  std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) {
    return b->frame_offsets[x + 1] - b->frame_offsets[x] > b->frame_offsets[y + 1] - b->frame_offsets[y];
  });
  info.insert(info.end(), order.begin(), order.end());
  int64_t acc = 0;
  info.push_back(0);
  for (int64_t k = 0; k < b->nutts; ++k) {
    acc += b->frame_offsets[order[k] + 1] - b->frame_offsets[order[k]];
    info.push_back(acc);
  }
  b->down_offsets = info;       // uploaded by snb_batch_create with the other tables
  return SNB_OK;
}


// ---------------------------------------------------------------------------
// k1: LinearResample::Resample for the whole batch.  The sum runs in double
// like the oracle's (weight x int16 sample is exact in double, so one DFMA per
// tap is the oracle's "exact product, rounded add"), rounded to float once.
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
  double acc = 0.0;
  const int32_t nw = a.nw[wrapped];
  for (int32_t j = 0; j < nw; ++j) {
    const int64_t k = first_in + j;
    if (k >= 0 && k < n_in)
      acc = fma(static_cast<double>(w[j]), static_cast<double>(a.pcm[in0 + k]), acc);
  }
  a.down[idx] = static_cast<float>(acc);
}

// ---------------------------------------------------------------------------
// k2: NCCF ballast of every utterance.  Kaldi's online class normalises with
// the mean square of the downsampled signal "seen so far": offline that is
// two values, before and after the resampler is flushed.
// One CTA per utterance, double sums, fixed reduction order.
// ---------------------------------------------------------------------------
struct BallastArgs {
  const float *down;
  const int64_t *info;
  int32_t basic_len;
  float nccf_ballast;
  float *ballast;               // [nutts, 2]
};

__global__ void __launch_bounds__(256) pitch_ballast_kernel(const BallastArgs a) {
  __shared__ double s_red[4][8];
  const int64_t u = blockIdx.x;
  const int64_t doff = a.info[4 * u], m1 = a.info[4 * u + 1], m2 = a.info[4 * u + 2];
  const float *x = a.down + doff;
  double p1 = 0.0, q1 = 0.0, p2 = 0.0, q2 = 0.0;
  for (int64_t i = threadIdx.x; i < m2; i += blockDim.x) {
    const double v = x[i];
    if (i < m1) { p1 += v; q1 = fma(v, v, q1); } else { p2 += v; q2 = fma(v, v, q2); }
  }
  p1 = group_sum_f64<32>(p1); q1 = group_sum_f64<32>(q1);
  p2 = group_sum_f64<32>(p2); q2 = group_sum_f64<32>(q2);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_red[0][warp] = p1; s_red[1][warp] = q1; s_red[2][warp] = p2; s_red[3][warp] = q2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[4] = {0.0, 0.0, 0.0, 0.0};
    for (int k = 0; k < 4; ++k)
      for (int w = 0; w < 8; ++w) t[k] += s_red[k][w];
    const double sum1 = t[0], sq1 = t[1], sum2 = t[0] + t[2], sq2 = t[1] + t[3];
    const double n1 = static_cast<double>(m1), n2 = static_cast<double>(m2);
    const double ms1 = m1 > 0 ? sq1 / n1 - (sum1 / n1) * (sum1 / n1) : 0.0;
    const double ms2 = m2 > 0 ? sq2 / n2 - (sum2 / n2) * (sum2 / n2) : 0.0;
    a.ballast[2 * u] = static_cast<float>((ms1 * a.basic_len) * (ms1 * a.basic_len) *
                                           static_cast<double>(a.nccf_ballast));
    a.ballast[2 * u + 1] = static_cast<float>((ms2 * a.basic_len) * (ms2 * a.basic_len) *
                                               static_cast<double>(a.nccf_ballast));
  }
}

// ---------------------------------------------------------------------------
// k3: frame-parallel front end of the tracker.  Everything of a frame that does
// not depend on the previous frame -- ExtractFrame, mean removal, NCCF at the
// integer lags (pitch and POV flavours), sinc upsampling to the log-spaced lags
// and the local cost of every Viterbi state -- for ALL frames of a group of
// utterances at full occupancy (round 1 ran this inside the sequential
// warp-per-utterance loop: 26 % of its instructions).
//
// One warp per frame, tasks of kNccfTask consecutive frames.  Every dot
// product accumulates in double in index order, exactly like the oracle's
// dotf() (oracle/kaldi_oracle.c), and is rounded to float once: the local
// costs, hence the Viterbi state sequence, are bit-identical to the oracle's.
//
// Outputs: cost [group_frames, ns4] float32 (row stride ns4 = ns rounded up to
// 4: rows are 16-byte aligned for the tracker's cp.async), pov [group_frames, nm].
// ---------------------------------------------------------------------------
constexpr int kNccfTask = 8;
constexpr int kNccfWarps = 12;          // per CTA, two CTAs per SM (80 registers)

struct NccfArgs {
  const float *down;
  const int64_t *info;          // [nutts,4]
  const int64_t *order;         // [nutts] utterances by decreasing number of frames
  const int64_t *gfo;           // [nutts+1] cumulated frames in that order
  const float *ballast;         // [nutts,2]
  int64_t k0, k1, q0, q1;       // group: sorted utterances [k0,k1), frames [q0,q1)
  int32_t first_lag, nmeas, nstates, ns4, shift, basic_len, full_len, nw, nwp;
  int32_t snip_edges, table_in_smem;
  float preemph, soft_min_f0;
  const float *lags;
  const double *up_w_t;         // [ns, nwp] taps
  const int32_t *up_first;
  float *cost, *pov;
};

struct NccfSmem {               // per-warp offsets in doubles
  int z, pre, np, total;
};
__host__ __device__ inline NccfSmem nccf_smem_layout(int full_len, int nm, int nw) {
  NccfSmem s;
  int off = 0;
  s.z = off; off += (full_len + 6 + 1) & ~1;      // zero tail read by the sliding NCCF loop
  s.pre = off; off += (full_len + 2 + 1) & ~1;
  s.np = off; off += 2 * ((nm + nw + 1 + 1) & ~1);   // two frames; zero tails for the padded taps
  s.total = off;
  return s;
}
__host__ __device__ inline int nccf_shared_words(int ns, int nwp, bool table) {
  return (table ? 2 * (((ns * nwp) + 1) & ~1) : 0) + 2 * ((ns + 1) & ~1);   // taps (doubles), up_first, soft_min_f0*lag
}

// NWC > 0: number of upsampling taps known at compile time (10 for Kaldi's
// defaults) and the tap table in shared memory; 0: any tap count, table in
// shared memory when it fits, else read through L1
template <int NWC>
__global__ void __launch_bounds__(kNccfWarps * 32, 2) pitch_nccf_kernel(const NccfArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarp_cta = blockDim.x >> 5;
  const int ns = a.nstates, nm = a.nmeas, bl = a.basic_len, fl = a.full_len;
  const int nw = NWC > 0 ? NWC : a.nw, nwp = nw | 1;
  const bool table = NWC > 0 || a.table_in_smem;
  // ---- CTA-shared tables ----
  double *s_upw = reinterpret_cast<double *>(smem_raw);
  const int tab_words = table ? 2 * (((ns * nwp) + 1) & ~1) : 0;
  int32_t *s_upfirst = reinterpret_cast<int32_t *>(smem_raw) + tab_words;
  float *s_lagc = reinterpret_cast<float *>(s_upfirst + ((ns + 1) & ~1));
  if (table)
    for (int i = tid; i < ns * nwp; i += blockDim.x) s_upw[i] = a.up_w_t[i];
  for (int i = tid; i < ns; i += blockDim.x) {
    s_upfirst[i] = min(max(a.up_first[i], 0), nm - 1);       // (a state without taps has zero weights)
    s_lagc[i] = __fmul_rn(a.soft_min_f0, a.lags[i]);
  }
  __syncthreads();
  // ---- warp-private buffers ----
  const NccfSmem L = nccf_smem_layout(fl, nm, nw);
  double *wbase = reinterpret_cast<double *>(smem_raw + 4 * static_cast<size_t>(nccf_shared_words(ns, nwp, table))) +
                  static_cast<size_t>(warp) * L.total;
  double *w_z = wbase + L.z, *w_pre = wbase + L.pre, *w_np0 = wbase + L.np;
  const int np_stride = (nm + nw + 1 + 1) & ~1;
  for (int i = fl + lane; i < fl + 6; i += 32) w_z[i] = 0.0;
  for (int i = nm + lane; i < nm + nw + 1; i += 32) { w_np0[i] = 0.0; w_np0[np_stride + i] = 0.0; }

  const int64_t ntasks = (a.q1 - a.q0 + kNccfTask - 1) / kNccfTask;
  const int64_t gw = static_cast<int64_t>(blockIdx.x) * nwarp_cta + warp;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * nwarp_cta;
  for (int64_t task = gw; task < ntasks; task += nwarps) {
    const int64_t qa = a.q0 + task * kNccfTask, qb = min(qa + kNccfTask, a.q1);
    // utterance (in sorted order) of the task's first frame: equal-length guess, else bisection
    int64_t k;
    {
      k = a.k0 + static_cast<int64_t>(static_cast<double>(qa - a.q0) * static_cast<double>(a.k1 - a.k0) /
                                      static_cast<double>(a.q1 - a.q0));
      k = min(max(k, a.k0), a.k1 - 1);
      if (!(a.gfo[k] <= qa && qa < a.gfo[k + 1])) {
        int64_t lo = a.k0, hi = a.k1;
        while (hi - lo > 1) {
          const int64_t mid = (lo + hi) >> 1;
          if (a.gfo[mid] <= qa) lo = mid; else hi = mid;
        }
        k = lo;
      }
    }
    int64_t kcur = -1, doff = 0, m1 = 0, m2 = 0, end1 = 0, fbase = 0;
    float ballast1 = 0.0f, ballast2 = 0.0f;
    // two frames per pass: their NCCF rows are upsampled together (one load of
    // every tap weight for both)
#pragma unroll 1
    for (int64_t q2 = qa; q2 < qb; q2 += 2) {
     const int npass = q2 + 1 < qb ? 2 : 1;
#pragma unroll 1
     for (int sub = 0; sub < npass; ++sub) {
      const int64_t q = q2 + sub;
      double *w_np = w_np0 + sub * np_stride;
      while (a.gfo[k + 1] <= q) ++k;
      if (k != kcur) {
        kcur = k;
        const int64_t u = a.order[k];
        doff = a.info[4 * u]; m1 = a.info[4 * u + 1]; m2 = a.info[4 * u + 2]; end1 = a.info[4 * u + 3];
        ballast1 = a.ballast[2 * u]; ballast2 = a.ballast[2 * u + 1];
        fbase = a.gfo[k];
      }
      const int64_t f = q - fbase;
      const bool phase2 = f >= end1;
      const int64_t avail = phase2 ? m2 : m1;
      const float ballast = phase2 ? ballast2 : ballast1;
      const float *x = a.down + doff;
      int64_t start;
      if (a.snip_edges) start = f * a.shift;
      else start = static_cast<int64_t>((static_cast<double>(f) + 0.5) * a.shift) - fl / 2;
      // ---- ExtractFrame (zeros beyond the available signal), optional pre-emphasis ----
      __syncwarp();
      if (a.preemph == 0.0f) {
        double s = 0.0;
        for (int i = lane; i < fl; i += 32) {
          const int64_t kk = start + i;
          const float v = (kk >= 0 && kk < avail) ? x[kk] : 0.0f;
          w_z[i] = static_cast<double>(v);
          if (i < bl) s += static_cast<double>(v);
        }
        const float mean = static_cast<float>(group_sum_f64<32>(s) / static_cast<double>(bl));
        __syncwarp();
        for (int i = lane; i < fl; i += 32)
          w_z[i] = static_cast<double>(__fadd_rn(static_cast<float>(w_z[i]), -mean));
      } else {
        // window[i] -= c * window[i-1] for i = fl-1 .. 1 (original neighbours), window[0] *= 1 - c
        for (int i = lane; i < fl; i += 32) {
          const int64_t kk = start + i;
          w_pre[i] = static_cast<double>((kk >= 0 && kk < avail) ? x[kk] : 0.0f);
        }
        __syncwarp();
        double s = 0.0;
        for (int i = lane; i < fl; i += 32) {
          const float cur = static_cast<float>(w_pre[i]);
          float v;
          if (i > 0) v = __fadd_rn(cur, -__fmul_rn(a.preemph, static_cast<float>(w_pre[i - 1])));
          else v = __fmul_rn(cur, static_cast<float>(1.0 - static_cast<double>(a.preemph)));
          w_z[i] = static_cast<double>(v);
          if (i < bl) s += static_cast<double>(v);
        }
        const float mean = static_cast<float>(group_sum_f64<32>(s) / static_cast<double>(bl));
        __syncwarp();
        for (int i = lane; i < fl; i += 32)
          w_z[i] = static_cast<double>(__fadd_rn(static_cast<float>(w_z[i]), -mean));
      }
      __syncwarp();
      // ---- double prefix sums of squares: pre[k] = sum_{i<k} z_i^2 (every lag's window energy
      //      is a difference of two entries) ----
      {
        const int chunk = (fl + 31) / 32;
        const int i0 = min(lane * chunk, fl), i1 = min(fl, i0 + chunk);
        double local = 0.0;
        for (int i = i0; i < i1; ++i) local = fma(w_z[i], w_z[i], local);
        double incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double up = __shfl_up_sync(SNB_FULL_MASK, incl, o);
          if (lane >= o) incl += up;
        }
        double run = incl - local;
        for (int i = i0; i < i1; ++i) {
          w_pre[i] = run;
          run = fma(w_z[i], w_z[i], run);
        }
        if (i1 == fl && i0 < fl) w_pre[fl] = run;
      }
      __syncwarp();
      const float e1 = static_cast<float>(w_pre[bl]);
      const int64_t qrel = q - a.q0;
      // ---- NCCF at the integer lags: a lane owns three consecutive lags, the lagged samples
      //      slide through six registers (2 + 4 LDS.64 per 12 DFMA) ----
#pragma unroll 1
      for (int lb = 0; lb < nm; lb += 96) {
        const int l0 = lb + 3 * lane;
        const double *zq = w_z + a.first_lag + min(l0, nm - 1);    // (reads up to 5 doubles of zero tail)
        double in0 = 0.0, in1 = 0.0, in2 = 0.0;
        double r0 = zq[0], r1 = zq[1];
        int i = 0;
#pragma unroll 5
        for (; i + 3 < bl; i += 4) {
          const double2 wa = *reinterpret_cast<const double2 *>(w_z + i);
          const double2 wb = *reinterpret_cast<const double2 *>(w_z + i + 2);
          const double r2 = zq[i + 2], r3 = zq[i + 3], r4 = zq[i + 4], r5 = zq[i + 5];
          in0 = fma(wa.x, r0, in0); in1 = fma(wa.x, r1, in1); in2 = fma(wa.x, r2, in2);
          in0 = fma(wa.y, r1, in0); in1 = fma(wa.y, r2, in1); in2 = fma(wa.y, r3, in2);
          in0 = fma(wb.x, r2, in0); in1 = fma(wb.x, r3, in1); in2 = fma(wb.x, r4, in2);
          in0 = fma(wb.y, r3, in0); in1 = fma(wb.y, r4, in1); in2 = fma(wb.y, r5, in2);
          r0 = r4; r1 = r5;
        }
        for (; i < bl; ++i) {
          const double wv = w_z[i];
          in0 = fma(wv, zq[i], in0); in1 = fma(wv, zq[i + 1], in1); in2 = fma(wv, zq[i + 2], in2);
        }
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
          const int l = l0 + kk;
          const float inner = static_cast<float>(kk == 0 ? in0 : (kk == 1 ? in1 : in2));
          if (l < nm) {
            const int lag = a.first_lag + l;
            const float e2 = static_cast<float>(w_pre[lag + bl] - w_pre[lag]);
            const float norm = __fmul_rn(e1, e2);
            const float den_p = __fsqrt_rn(__fadd_rn(norm, ballast));
            const float den_v = __fsqrt_rn(norm);
            w_np[l] = static_cast<double>(den_p != 0.0f ? __fdiv_rn(inner, den_p) : 0.0f);
            a.pov[qrel * nm + l] = den_v != 0.0f ? __fdiv_rn(inner, den_v) : 0.0f;
          }
        }
      }
     }
      __syncwarp();
      // ---- upsample to the log-spaced lags (taps padded with zero weights to nw); local cost ----
      const int64_t qrel2 = q2 - a.q0;
      float *crow = a.cost + qrel2 * a.ns4 + lane;
      const bool two = npass == 2;
#pragma unroll 2
      for (int i = lane; i < ns; i += 32, crow += 32) {
        const double *src = w_np0 + s_upfirst[i];
        double acc0 = 0.0, acc1 = 0.0;
        if (NWC > 0) {
          const double *w = s_upw + i * nwp;
#pragma unroll
          for (int j = 0; j < NWC; ++j) {
            const double wj = w[j];
            acc0 = fma(wj, src[j], acc0);
            acc1 = fma(wj, src[np_stride + j], acc1);
          }
        } else if (table) {
          const double *w = s_upw + i * nwp;
#pragma unroll 2
          for (int j = 0; j < nw; ++j) {
            acc0 = fma(w[j], src[j], acc0);
            acc1 = fma(w[j], src[np_stride + j], acc1);
          }
        } else {
          const double *w = a.up_w_t + static_cast<int64_t>(i) * nwp;
#pragma unroll 2
          for (int j = 0; j < nw; ++j) {
            const double wj = __ldg(w + j);
            acc0 = fma(wj, src[j], acc0);
            acc1 = fma(wj, src[np_stride + j], acc1);
          }
        }
        // local_cost = 1 - nccf; local_cost += soft_min_f0 * lag * nccf
        const float lagc = s_lagc[i];
        const float n0 = static_cast<float>(acc0), n1 = static_cast<float>(acc1);
        crow[0] = __fadd_rn(__fmul_rn(lagc, n0), __fadd_rn(1.0f, -n0));
        if (two) crow[a.ns4] = __fadd_rn(__fmul_rn(lagc, n1), __fadd_rn(1.0f, -n1));
      }
    }
  }
}

// ---------------------------------------------------------------------------
// k4: Viterbi over the frames of an utterance, one WARP per utterance (the
// recursion is sequential in time: the parallelism is utterances).  Reads the
// local costs of k3 (next frame's row prefetched with cp.async while the
// current step runs), keeps the forward costs in the warp's shared memory,
// writes int16 backpointers to a per-warp global scratch, then backtraces and
// emits (NCCF_pov at the chosen lag, 1 / lag).
//
// Viterbi step: bp(i) = first argmin_j pen[|i-j|] + prev[j] is non-decreasing
// in i (the penalty is convex), so bp(i) lies in [bp(i-s), bp(i+s)].
//   * anchors (multiples of kA1 and the last state): whole-warp scans, the
//     first and the last over every j, the others in bisection order over the
//     range left by their neighbours;
//   * levels s = kA1/2 ... 1: lane per state i = s(2k+1), serial scan of
//     [bp(i-s), bp(i+s)] -- 1-3 candidates on the plateaus of bp, 2s+1 where bp
//     follows i; the few states that straddle a jump of bp (range > 2s+2) are
//     handed to whole-warp scans instead of stalling their round.  (Four such
//     states at a time, one per group of 8 lanes with a 64-bit (cost, j) key,
//     was tried: 9 ms slower per 10 000 utterances, most rounds have one or two.)
// Every scan keeps the first minimum of its range, exactly like the
// brute-force step (and like Kaldi's bound-tightening search, which the oracle
// restates: checked equal on 834 000 states in round 1).
// ---------------------------------------------------------------------------
// One CTA per SM, up to 32 warps (4 736 utterances in flight).  (Two CTAs of 17 warps with int16
// backpointers in shared memory -- 34 utterances in flight per SM, two rounds
// instead of three for 10 000 utterances -- measured slower per frame: 7.0
// against 5.9 ns, profiles/r02_pitch_variants.txt.)
constexpr int kTrackWarpsMax = 32;
constexpr int kTrackWarpsMin = 4;
constexpr int kTrackCtasPerSm = 1;
#ifndef SNB_PITCH_A1LOG2
#define SNB_PITCH_A1LOG2 6
#endif
#ifndef SNB_PITCH_TMAX
#define SNB_PITCH_TMAX 32
#endif
constexpr int kA1Log2 = SNB_PITCH_A1LOG2, kA1 = 1 << kA1Log2;

struct TrackArgs {
  const float *cost, *pov;
  const int64_t *order, *gfo, *frame_offsets;
  int64_t k0, k1, q0;
  int32_t nstates, ns4, nmeas, nw, nwp;
  const float *lags, *pen;
  const double *up_w_t;
  const int32_t *up_first;
  int16_t *bp;            // [slots, max_frames, nstates]
  int32_t *states;        // [slots, max_frames]
  int64_t max_frames;
  float *out;
  int64_t ld_out;
  unsigned long long *queue;   // next utterance to hand out (zeroed before the launch)
};

// backpointers of the current frame are read with strides 2s: skew the index
__host__ __device__ inline int bp_slot(int i) { return i + (i >> 5); }

__host__ __device__ inline int track_warp_words(int ns4, int ns) {
  return 3 * ns4 + ((bp_slot(ns) + 4) & ~3);     // prev | cur | next cost rows, backpointers
}
__host__ __device__ inline int track_shared_words(int ns) {
  return (2 * ns + 3) & ~3;                       // penalties mirrored around the centre
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

// one 16-byte cp.async per lane and round: row [ns4] floats, global -> shared
__device__ __forceinline__ void prefetch_row(float *dst, const float *src, int ns4, int lane) {
  for (int i = 4 * lane; i < ns4; i += 128)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + i)), "l"(src + i) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
}

__global__ void __launch_bounds__(kTrackWarpsMax * 32, kTrackCtasPerSm) pitch_viterbi_kernel(const TrackArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarp_cta = blockDim.x >> 5;
  const int ns = a.nstates, ns4 = a.ns4, nm = a.nmeas, nw = a.nw;
  // penalty table mirrored around s_pen: s_pen[d] = pen[|d|], d in (-ns, ns):
  // the scans index it with the signed state distance, no absolute value
  float *s_pen = reinterpret_cast<float *>(smem_raw) + (ns - 1);
  for (int i = tid; i < ns; i += blockDim.x) { s_pen[i] = a.pen[i]; s_pen[-i] = a.pen[i]; }
  __syncthreads();
  float *wbase = reinterpret_cast<float *>(smem_raw) + track_shared_words(ns) +
                 static_cast<size_t>(warp) * track_warp_words(ns4, ns);
  float *buf0 = wbase, *buf1 = wbase + ns4, *buf2 = wbase + 2 * ns4;
  int *w_bp = reinterpret_cast<int *>(wbase + 3 * ns4);     // this frame's backpointers

  const int64_t gw = static_cast<int64_t>(blockIdx.x) * nwarp_cta + warp;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * nwarp_cta;
  int16_t *bp = a.bp + gw * a.max_frames * ns;
  int32_t *states = a.states + gw * a.max_frames;
  const int n1 = (ns - 1 + kA1 - 1) / kA1 + 1;                // anchors min(t kA1, ns-1), t < n1
  int top = 1;
  while (top < n1) top <<= 1;

  int64_t k = a.k0 + gw;
  while (k < a.k1) {
    const int64_t u = a.order[k];
    const int64_t F = a.gfo[k + 1] - a.gfo[k];
    const int64_t row0 = a.frame_offsets[u];
    const float *crow = a.cost + (a.gfo[k] - a.q0) * ns4;
    const float *prow = a.pov + (a.gfo[k] - a.q0) * nm;
    float *w_prev = buf0, *w_cost = buf1, *w_next = buf2;
    if (F > 0) prefetch_row(w_cost, crow, ns4, lane);
    for (int i = lane; i < ns; i += 32) w_prev[i] = 0.0f;

    for (int64_t f = 0; f < F; ++f) {
      if (f + 1 < F) {
        prefetch_row(w_next, crow + (f + 1) * ns4, ns4, lane);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncwarp();
      // ---- anchors by whole-warp scans, bisection order ----
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
          const int kk = k0 + lane;
          const bool act = kk < nodd;
          const int i = s * (2 * kk + 1);
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
      for (int i = lane; i < ns; i += 32) w_cost[i] = __fadd_rn(w_cost[i], -lmin);
      __syncwarp();
      float *t = w_prev; w_prev = w_cost; w_cost = w_next; w_next = t;
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
        const int first = min(max(a.up_first[sidx], 0), nm - 1);
        const double *w = a.up_w_t + static_cast<int64_t>(sidx) * a.nwp;
        const float *pv = prow + f * nm + first;
        const int n = min(nw, nm - first);                     // taps beyond are zero weights
        double acc = 0.0;
#pragma unroll 1
        for (int j = 0; j < n; ++j) acc = fma(w[j], static_cast<double>(pv[j]), acc);
        float *o = a.out + (row0 + f) * a.ld_out;
        o[0] = static_cast<float>(acc);
        o[1] = __fdiv_rn(1.0f, a.lags[sidx]);
      }
      __syncwarp();
    }
    // ---- next utterance from the queue ----
    unsigned long long nxt = 0;
    if (lane == 0) nxt = atomicAdd(a.queue, 1ULL);
    nxt = __shfl_sync(SNB_FULL_MASK, nxt, 0);
    k = a.k0 + nwarps + static_cast<int64_t>(nxt);
  }
}

// ---------------------------------------------------------------------------
// k5: ProcessPitch (OnlineProcessPitch in offline use).  Output row t of an
// utterance holds the features of input frame max(0, t - delay); an utterance
// of F > 0 frames gives F + delay rows (Kaldi's NumFramesReady once the input
// is finished), written from row frame_offsets[u] + u * delay.
// grid = (nutts, chunks); each CTA owns kPostRows output rows of one utterance
// plus the halo the normalisation window needs
// ---------------------------------------------------------------------------
constexpr int kPostRows = 128;

struct PostArgs {
  snb_pitch_post_opts o;
  const float *raw;
  int64_t ld_raw;
  const int64_t *frame_offsets;
  const int64_t *out_offsets;   // optional output geometry (rows beyond it are dropped)
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
  const int64_t delay = a.o.delay;
  int64_t Fo = F > 0 ? F + delay : 0;
  int64_t first_out = first + u * delay;
  if (a.out_offsets) {
    first_out = a.out_offsets[u];
    Fo = min(Fo, a.out_offsets[u + 1] - first_out);
  }
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * kPostRows;
  if (r0 >= Fo) return;
  const int span = kPostRows + 2 * a.halo;
  float *s_logp = s_post, *s_pov = s_post + span;
  const int64_t tb = r0 - delay;                                 // input frame of the CTA's first row
  const int64_t lo = (tb > 0 ? tb : 0) - a.halo;
  for (int i = threadIdx.x; i < span; i += blockDim.x) {
    const int64_t t = lo + i;
    float lp = 0.0f, pv = 0.0f;
    if (t >= 0 && t < F) {
      const float *row = a.raw + (first + t) * a.ld_raw;
      lp = static_cast<float>(log(static_cast<double>(row[1])));
      pv = nccf_to_pov(row[0]);
    }
    s_logp[i] = lp;
    s_pov[i] = pv;
  }
  __syncthreads();
  const snb_pitch_post_opts &o = a.o;
  for (int r = threadIdx.x; r < kPostRows; r += blockDim.x) {
    const int64_t to = r0 + r;
    if (to >= Fo) break;
    const int64_t t = to < delay ? 0 : to - delay;
    float *out = a.out + (first_out + to) * a.ld_out;
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
        if (s != 0.0f) d = __fadd_rn(d, __fmul_rn(s, s_logp[tt - lo]));     // (axpy: product, then sum)
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

// ---------------------------------------------------------------------------
// launch geometry
// ---------------------------------------------------------------------------
constexpr size_t kSmemBudget = 224 * 1024;

static int sm_count() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
    cudaGetLastError();
    sms = 148;
  }
  return sms;
}

static size_t track_smem(const PitchTables *t, int warps) {
  return (static_cast<size_t>(track_shared_words(t->nstates)) +
          static_cast<size_t>(warps) * track_warp_words(t->ns4, t->nstates)) * 4 + 16;
}

// most warps (utterances in flight) one CTA of the tracker can hold with
// kTrackCtasPerSm CTAs per SM
static int track_capacity(const PitchTables *t) {
  static const char *env = getenv("SNB_PITCH_WARPS");          // tuning knob: cap the warps per CTA
  int w = kTrackWarpsMax;
  if (env && atoi(env) >= kTrackWarpsMin) w = std::min(w, atoi(env));
  while (w > 0 && track_smem(t, w) > kSmemBudget / kTrackCtasPerSm) --w;
  return w;
}

// Tracker launch shape for `nutts` utterances: kTrackCtasPerSm CTAs per SM;
// the number of warps per CTA balances the waves (10 000 utterances on 148
// SMs x 34 warps run two rounds of 5 000).  Returns the number of concurrently
// tracked utterances ("slots" of backpointer scratch).
static int64_t track_shape(const PitchTables *t, int64_t nutts, int *grid_out, int *warps_out) {
  const int cap = track_capacity(t);
  const int64_t ctas = static_cast<int64_t>(sm_count()) * kTrackCtasPerSm;
  const int64_t n = std::max<int64_t>(nutts, 1);
  if (cap < 1) {
    if (grid_out) *grid_out = 0;
    if (warps_out) *warps_out = 0;
    return 0;
  }
  const int64_t waves = (n + ctas * cap - 1) / (ctas * cap);
  const int64_t w = (n + waves * ctas - 1) / (waves * ctas);
  const int warps = static_cast<int>(std::max<int64_t>(std::min<int64_t>(kTrackWarpsMin, cap), std::min<int64_t>(w, cap)));
  const int grid = static_cast<int>(std::min<int64_t>((n + warps - 1) / warps, ctas));
  if (grid_out) *grid_out = grid;
  if (warps_out) *warps_out = warps;
  return static_cast<int64_t>(grid) * warps;
}

struct NccfShape { int warps; bool table; size_t smem; };
static NccfShape nccf_shape(const PitchTables *t) {
  NccfShape s;
  const NccfSmem L = nccf_smem_layout(t->full_len, t->nmeas, t->up_nw_max);
  s.table = t->up_nw_max == 10 || static_cast<size_t>(t->nstates) * t->nwp * 8 <= 96 * 1024;
  s.warps = kNccfWarps;
  for (;;) {
    s.smem = static_cast<size_t>(nccf_shared_words(t->nstates, t->nwp, s.table)) * 4 +
             static_cast<size_t>(s.warps) * L.total * 8 + 16;
    if (s.smem <= 110 * 1024 || s.warps == 1) break;        // two CTAs per SM when possible
    s.warps >>= 1;
  }
  return s;
}

// Groups of utterances (in sorted order) whose local costs are alive at once.
// The cost matrix is 4 (ns4 + nm) bytes per frame: a budget of kGroupBytes
// bounds the scratch whatever the batch.  The batch is cut in equal groups,
// each one tracker wave when it can be.
constexpr size_t kGroupBytes = static_cast<size_t>(9) << 30;

struct PitchGroups {
  int64_t per = 0;             // utterances per group
  int64_t group_frames = 0;    // most frames of a group
  int64_t max_frames = 0;      // longest utterance
  int64_t slots = 0;
  int grid = 0, warps = 0;
};

static PitchGroups plan_groups(const snb_batch *b, const PitchTables *t) {
  PitchGroups g;
  const int64_t *gfo = b->down_offsets.data() + 4 * (b->nutts + 1) + b->nutts;
  const size_t per_frame = static_cast<size_t>(t->ns4 + t->nmeas) * 4;
  const int64_t budget = std::max<int64_t>(1, static_cast<int64_t>(kGroupBytes / per_frame));
  const int64_t total = b->nutts > 0 ? gfo[b->nutts] : 0;
  int64_t ngroups = std::max<int64_t>(1, (total + budget - 1) / budget);
  {
    // ... and one group per tracker wave: 10 000 utterances on 148 x 32 slots
    // are three groups of 3 334 (23 warps per CTA), not 2.1 waves
    const int cap = track_capacity(t);
    const int64_t per_wave = static_cast<int64_t>(sm_count()) * kTrackCtasPerSm * std::max(cap, 1);
    ngroups = std::max(ngroups, (b->nutts + per_wave - 1) / per_wave);
  }
  for (;;) {                    // (sorted by decreasing length: the first group is the heaviest)
    g.per = std::max<int64_t>(1, (b->nutts + ngroups - 1) / ngroups);
    if (g.per == 1 || gfo[std::min(g.per, b->nutts)] <= budget) break;
    ++ngroups;
  }
  for (int64_t k0 = 0; k0 < b->nutts; k0 += g.per)
    g.group_frames = std::max(g.group_frames, gfo[std::min(k0 + g.per, b->nutts)] - gfo[k0]);
  g.max_frames = b->nutts > 0 ? gfo[1] - gfo[0] : 0;     // sorted: the first utterance is the longest
  g.slots = track_shape(t, std::min(g.per, std::max<int64_t>(b->nutts, 1)), &g.grid, &g.warps);
  return g;
}

}  // namespace snb

using namespace snb;

extern "C" int64_t snb_pitch_num_frames(int64_t nsamples, const snb_pitch_opts *po) {
  if (!pitch_opts_valid(*po)) return 0;
  int32_t first, last;
  lag_range(*po, &first, &last);
  return frames_available(resample_num_out(nsamples, *po, true), *po, last, true);
}

extern "C" void snb_pitch_num_frames_array(const int64_t *nsamples, int64_t n, const snb_pitch_opts *po,
                                           int64_t *out) {
  for (int64_t i = 0; i < n; ++i) out[i] = snb_pitch_num_frames(nsamples[i], po);
}

extern "C" int64_t snb_pitch_wave_utts(const snb_plan *plan) {
  if (!plan || plan->kind != 1) return 0;
  return static_cast<int64_t>(sm_count()) * kTrackCtasPerSm * std::max(track_capacity(plan->pitch), 1);
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

static size_t align256(size_t v) { return (v + 255) / 256 * 256; }

extern "C" int64_t snb_pitch_workspace_bytes(const snb_plan *plan, const snb_batch *batch) {
  if (!plan || plan->kind != 1 || !batch) return -1;
  const PitchTables *t = plan->pitch;
  const PitchGroups g = plan_groups(batch, t);
  const int64_t mf = std::max<int64_t>(1, g.max_frames);
  size_t bytes = align256(static_cast<size_t>(batch->total_down + 8) * 4) + 256;   // + utterance queue
  bytes += align256(static_cast<size_t>(batch->nutts + 1) * 2 * 4);                 // ballast
  bytes += align256(static_cast<size_t>(g.group_frames + 1) * t->ns4 * 4);          // local costs
  bytes += align256(static_cast<size_t>(g.group_frames + 1) * t->nmeas * 4);        // POV NCCF
  bytes += align256(static_cast<size_t>(g.slots) * mf * t->nstates * 2);            // backpointers
  bytes += align256(static_cast<size_t>(g.slots) * mf * 4);                         // state sequences
  return static_cast<int64_t>(bytes);
}

template <typename K>
static int raise_smem(K kernel, std::atomic<size_t> &cur, size_t want) {
  size_t c = cur.load();
  while (want > c) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(want));
    if (e != cudaSuccess) return set_error(SNB_ERR_CUDA, "pitch smem: %s", cudaGetErrorString(e));
    if (cur.compare_exchange_weak(c, want)) break;
  }
  return SNB_OK;
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
  const PitchGroups pg = plan_groups(batch, t);
  const int grid = pg.grid, warps = pg.warps;
  const int64_t slots = pg.slots;
  if (warps < 1) return set_error(SNB_ERR_UNSUPPORTED, "pitch options outside the GPU path limits");
  const int64_t mf = std::max<int64_t>(1, pg.max_frames);
  unsigned char *ws = static_cast<unsigned char *>(d_workspace);
  float *down = reinterpret_cast<float *>(ws);
  ws += align256(static_cast<size_t>(batch->total_down + 8) * 4);
  unsigned long long *queue = reinterpret_cast<unsigned long long *>(ws);
  ws += 256;
  float *ballast = reinterpret_cast<float *>(ws);
  ws += align256(static_cast<size_t>(batch->nutts + 1) * 2 * 4);
  float *cost = reinterpret_cast<float *>(ws);
  ws += align256(static_cast<size_t>(pg.group_frames + 1) * t->ns4 * 4);
  float *pov = reinterpret_cast<float *>(ws);
  ws += align256(static_cast<size_t>(pg.group_frames + 1) * t->nmeas * 4);
  int16_t *bp = reinterpret_cast<int16_t *>(ws);
  ws += align256(static_cast<size_t>(slots) * mf * t->nstates * 2);
  int32_t *states = reinterpret_cast<int32_t *>(ws);

  const int64_t *d_info = batch->d_down_offsets;
  const int64_t *d_order = d_info + 4 * (batch->nutts + 1);
  const int64_t *d_gfo = d_order + batch->nutts;
  const int64_t *h_gfo = batch->down_offsets.data() + 4 * (batch->nutts + 1) + batch->nutts;

  ResampleArgs r;
  r.pcm = d_pcm;
  r.sample_begin = batch->d_sample_begin;
  r.sample_len = batch->d_sample_len;
  r.info = d_info;
  r.nutts = batch->nutts;
  r.total_down = batch->total_down;
  r.in_unit = t->in_unit; r.out_unit = t->out_unit; r.nw_max = t->down_nw_max;
  r.first = t->d_down_first; r.nw = t->d_down_nw; r.w = t->d_down_w;
  r.down = down;
  if (batch->total_down > 0) {
    resample_kernel<<<static_cast<unsigned>((batch->total_down + 255) / 256), 256, 0, stream>>>(r);
    SNB_LAUNCH_CHECK();
  }
  BallastArgs ba;
  ba.down = down; ba.info = d_info; ba.basic_len = t->basic_len;
  ba.nccf_ballast = plan->po.nccf_ballast; ba.ballast = ballast;
  pitch_ballast_kernel<<<static_cast<unsigned>(batch->nutts), 256, 0, stream>>>(ba);
  SNB_LAUNCH_CHECK();

  const NccfShape nshape = nccf_shape(t);
  static std::atomic<size_t> nccf_cur{48 * 1024}, nccf_cur0{48 * 1024}, track_cur{48 * 1024};
  const bool nw10 = t->up_nw_max == 10 && nshape.table;     // Kaldi's default resampling options
  int rc = nw10 ? raise_smem(pitch_nccf_kernel<10>, nccf_cur, nshape.smem)
                : raise_smem(pitch_nccf_kernel<0>, nccf_cur0, nshape.smem);
  if (rc != SNB_OK) return rc;
  const size_t tsmem = track_smem(t, warps);
  rc = raise_smem(pitch_viterbi_kernel, track_cur, tsmem);
  if (rc != SNB_OK) return rc;
  const int sms = sm_count();

  for (int64_t k0 = 0; k0 < batch->nutts;) {
    const int64_t k1 = std::min(k0 + pg.per, batch->nutts);
    const int64_t q0 = h_gfo[k0], q1 = h_gfo[k1];
    if (q1 > q0) {
      NccfArgs n;
      n.down = down; n.info = d_info; n.order = d_order; n.gfo = d_gfo; n.ballast = ballast;
      n.k0 = k0; n.k1 = k1; n.q0 = q0; n.q1 = q1;
      n.first_lag = t->first_lag; n.nmeas = t->nmeas; n.nstates = t->nstates; n.ns4 = t->ns4;
      n.shift = t->shift; n.basic_len = t->basic_len; n.full_len = t->full_len;
      n.nw = t->up_nw_max; n.nwp = t->nwp;
      n.snip_edges = plan->po.snip_edges; n.table_in_smem = nshape.table ? 1 : 0;
      n.preemph = plan->po.preemph_coeff; n.soft_min_f0 = plan->po.soft_min_f0;
      n.lags = t->d_lags; n.up_w_t = t->d_up_w_t; n.up_first = t->d_up_first;
      n.cost = cost; n.pov = pov;
      const int64_t ntasks = (q1 - q0 + kNccfTask - 1) / kNccfTask;
      const int64_t want = (ntasks + nshape.warps - 1) / nshape.warps;
      const unsigned ngrid = static_cast<unsigned>(std::min<int64_t>(want, static_cast<int64_t>(sms) * 2));
      if (nw10) pitch_nccf_kernel<10><<<ngrid, nshape.warps * 32, nshape.smem, stream>>>(n);
      else pitch_nccf_kernel<0><<<ngrid, nshape.warps * 32, nshape.smem, stream>>>(n);
      SNB_LAUNCH_CHECK();

      TrackArgs a;
      a.cost = cost; a.pov = pov;
      a.order = d_order; a.gfo = d_gfo; a.frame_offsets = batch->d_frame_offsets;
      a.k0 = k0; a.k1 = k1; a.q0 = q0;
      a.nstates = t->nstates; a.ns4 = t->ns4; a.nmeas = t->nmeas; a.nw = t->up_nw_max; a.nwp = t->nwp;
      a.lags = t->d_lags; a.pen = t->d_pen; a.up_w_t = t->d_up_w_t; a.up_first = t->d_up_first;
      a.bp = bp; a.states = states; a.max_frames = mf;
      a.out = d_out; a.ld_out = ld_out;
      a.queue = queue;
      cudaError_t e = cudaMemsetAsync(queue, 0, sizeof(unsigned long long), stream);
      if (e != cudaSuccess) return set_error(SNB_ERR_CUDA, "pitch queue: %s", cudaGetErrorString(e));
      const int64_t gutts = k1 - k0;
      const int ggrid = static_cast<int>(std::min<int64_t>((gutts + warps - 1) / warps, grid));
      pitch_viterbi_kernel<<<static_cast<unsigned>(ggrid), warps * 32, tsmem, stream>>>(a);
      SNB_LAUNCH_CHECK();
    }
    k0 = k1;
  }
  return SNB_OK;
}

extern "C" int32_t snb_process_pitch_dim(const snb_pitch_post_opts *o) {
  return (o->add_pov_feature != 0) + (o->add_normalized_log_pitch != 0) + (o->add_delta_pitch != 0) +
         (o->add_raw_log_pitch != 0);
}

extern "C" int snb_process_pitch(const snb_pitch_post_opts *o, const float *d_raw, int64_t ld_raw,
                                 const int64_t *d_frame_offsets, const int64_t *d_out_frame_offsets,
                                 int64_t nutts, int64_t total_frames, int64_t max_frames, uint64_t seed,
                                 float *d_out, int64_t ld_out, void *stream) {
  if (!o) return set_error(SNB_ERR_VALUE, "null options");
  const int dim = snb_process_pitch_dim(o);
  if (dim == 0)
    return set_error(SNB_ERR_VALUE, "at least one of the pitch features must be enabled");
  // Kaldi asserts frame - delay < NumFramesReady() = F + delay: a negative delay aborts there
  if (o->delay < 0) return set_error(SNB_ERR_OPTION, "delay must be >= 0");
  if (total_frames == 0 || nutts == 0) return SNB_OK;
  if (!d_raw || !d_out || ld_raw < 2 || ld_out < dim) return set_error(SNB_ERR_VALUE, "bad argument");
  if (o->normalization_left_context < 0 || o->normalization_right_context < 0 || o->delta_window <= 0)
    return set_error(SNB_ERR_VALUE, "invalid pitch post-processing contexts");
  PostArgs a;
  a.o = *o;
  a.raw = d_raw; a.ld_raw = ld_raw;
  a.frame_offsets = d_frame_offsets;
  a.out_offsets = d_out_frame_offsets;
  a.seed = seed;
  a.out = d_out; a.ld_out = ld_out;
  a.halo = std::max(std::max(o->normalization_left_context, o->normalization_right_context), o->delta_window);
  const size_t smem = static_cast<size_t>(kPostRows + 2 * a.halo) * 2 * 4;
  if (smem > 200 * 1024) return set_error(SNB_ERR_UNSUPPORTED, "normalisation context too large");
  static std::atomic<size_t> cur{48 * 1024};
  int rc = raise_smem(process_pitch_kernel, cur, smem);
  if (rc != SNB_OK) return rc;
  const unsigned chunks =
      static_cast<unsigned>((std::max<int64_t>(max_frames, 1) + o->delay + kPostRows - 1) / kPostRows);
  if (chunks > 65535) return set_error(SNB_ERR_UNSUPPORTED, "utterance too long for process_pitch");
  process_pitch_kernel<<<dim3(static_cast<unsigned>(nutts), chunks), 256, smem, static_cast<cudaStream_t>(stream)>>>(a);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
