// post.cu -- streaming post-processors on packed [total_frames, ld] matrices:
// deltas (feature-functions.cc DeltaFeatures, postprocessor/delta.py:130),
// CMVN accumulate / apply (transform/cmvn.cc, postprocessor/cmvn.py:217-278),
// sliding-window CMN (cmvn.py:492), energy VAD (postprocessor/vad.py:183).
// These are HBM-bound: each input element is read once from DRAM (neighbour
// rows come from L1/L2), each output written once, rows are coalesced.
#include <atomic>
#include <cfloat>
#include <cmath>
#include <vector>

#include "device_utils.cuh"
#include "snb_internal.h"

namespace snb {

constexpr int kRowsPerCta = 32;
constexpr int kMaxTaps = 448;

struct DeltaArgs {
  const float *in;
  int64_t ld_in;
  int32_t dim;
  const int64_t *frame_offsets;
  int64_t nutts, total_frames;
  const float *norm;          // [ngroups, 2, dim] or NULL
  const int32_t *utt_group;   // or NULL
  int32_t order;
  int32_t tap_off[8];         // start of each order's taps in taps[]
  int32_t tap_half[8];        // max offset of each order
  float taps[kMaxTaps];
  float *out;
  int64_t ld_out;
};

__device__ __forceinline__ int64_t find_utt_row(const int64_t *offsets, int64_t nutts, int64_t row) {
  int64_t lo = 0, hi = nutts;
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (offsets[mid] <= row) lo = mid; else hi = mid;
  }
  return lo;
}

// out[t, i*dim + d] = sum_j taps_i[j] * norm(in[clamp(t + j), d])
__global__ void __launch_bounds__(256) delta_kernel(const DeltaArgs a) {
  __shared__ int64_t s_first[kRowsPerCta], s_last[kRowsPerCta];
  __shared__ int32_t s_group[kRowsPerCta];
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * kRowsPerCta;
  const int nrows = static_cast<int>(min(static_cast<int64_t>(kRowsPerCta), a.total_frames - row0));
  if (threadIdx.x < nrows) {
    const int64_t u = find_utt_row(a.frame_offsets, a.nutts, row0 + threadIdx.x);
    s_first[threadIdx.x] = a.frame_offsets[u];
    s_last[threadIdx.x] = a.frame_offsets[u + 1] - 1;
    s_group[threadIdx.x] = a.utt_group ? a.utt_group[u] : static_cast<int32_t>(u);
  }
  __syncthreads();
  const int dim = a.dim;
  for (int e = threadIdx.x; e < nrows * dim; e += blockDim.x) {
    const int r = e / dim, d = e - r * dim;
    const int64_t t = row0 + r, first = s_first[r], last = s_last[r];
    float scale = 1.0f, offset = 0.0f;
    const bool do_norm = a.norm != nullptr;
    if (do_norm) {
      const float *n = a.norm + static_cast<int64_t>(s_group[r]) * 2 * dim;
      offset = n[d];
      scale = n[dim + d];
    }
    float *o = a.out + t * a.ld_out;
    for (int i = 0; i <= a.order; ++i) {
      const int half = a.tap_half[i];
      const float *taps = a.taps + a.tap_off[i];
      float acc = 0.0f;
      for (int j = -half; j <= half; ++j) {
        const float s = taps[j + half];
        if (s == 0.0f) continue;
        int64_t tt = t + j;
        tt = tt < first ? first : (tt > last ? last : tt);
        float x = a.in[tt * a.ld_in + d];
        // ApplyCmvn: MulColsVec then AddVecToRows (two roundings)
        if (do_norm) x = __fadd_rn(__fmul_rn(x, scale), offset);
        acc = fmaf(s, x, acc);
      }
      o[i * dim + d] = acc;
    }
  }
}

// Tiled variant (ORDER <= 3, halo <= 16 rows): a CTA stages kTileRows + 2*halo
// NORMALISED rows in shared memory with coalesced loads (each input element is
// read once from global memory and normalised once); one thread then owns one
// (row, dimension) pair: it reads the 2*halo+1 neighbours once from shared
// memory and accumulates all ORDER+1 outputs in registers.
constexpr int kTileRows = 64;
constexpr int kMaxHalo = 16;

template <int ORDER>
__global__ void __launch_bounds__(256) delta_tiled_kernel(const DeltaArgs a, const int halo) {
  extern __shared__ float s_x[];                        // [nload, dim]
  __shared__ int s_lo[kTileRows + 2 * kMaxHalo], s_hi[kTileRows + 2 * kMaxHalo];   // clamp bounds (tile rows)
  __shared__ int32_t s_group[kTileRows + 2 * kMaxHalo];
  __shared__ float s_taps[kMaxTaps];
  const int tid = threadIdx.x;
  const int dim = a.dim;
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * kTileRows;
  const int64_t lo = row0 - halo;
  const int nload = kTileRows + 2 * halo;
  const int nrows = static_cast<int>(min(static_cast<int64_t>(kTileRows), a.total_frames - row0));
  // one binary search per tile (thread 0); rows of the same utterance reuse
  // it, the few rows beyond its end search on their own (in parallel)
  __shared__ int64_t s_u0[3];
  if (tid == 0) {
    const int64_t t = lo < 0 ? 0 : lo;
    const int64_t u = find_utt_row(a.frame_offsets, a.nutts, t);
    s_u0[0] = u; s_u0[1] = a.frame_offsets[u]; s_u0[2] = a.frame_offsets[u + 1];
  }
  __syncthreads();
  for (int i = tid; i < nload; i += 256) {
    const int64_t row = lo + i;
    int rlo = 0, rhi = -1, g = 0;
    if (row >= 0 && row < a.total_frames) {
      int64_t u = s_u0[0], first = s_u0[1], next = s_u0[2];
      if (row >= next) {
        u = find_utt_row(a.frame_offsets, a.nutts, row);
        first = a.frame_offsets[u]; next = a.frame_offsets[u + 1];
      }
      rlo = static_cast<int>(max(first - lo, static_cast<int64_t>(0)));
      rhi = static_cast<int>(min(next - 1 - lo, static_cast<int64_t>(nload - 1)));
      g = a.utt_group ? a.utt_group[u] : static_cast<int32_t>(u);
    }
    s_lo[i] = rlo; s_hi[i] = rhi; s_group[i] = g;
  }
  const int ntaps = a.tap_off[ORDER] + 2 * a.tap_half[ORDER] + 1;
  for (int i = tid; i < ntaps; i += 256) s_taps[i] = a.taps[i];
  __syncthreads();
  // ---- stage the normalised rows (flat, coalesced) ----
  const bool do_norm = a.norm != nullptr;
  {
    int i = tid / dim, d = tid - i * dim;
    const int step_i = 256 / dim, step_d = 256 - step_i * dim;
    for (int e = tid; e < nload * dim; e += 256) {
      float x = 0.0f;
      if (s_hi[i] >= s_lo[i]) {
        x = a.in[(lo + i) * a.ld_in + d];
        if (do_norm) {
          const float *n = a.norm + static_cast<int64_t>(s_group[i]) * 2 * dim;
          x = __fadd_rn(__fmul_rn(x, n[dim + d]), n[d]);
        }
      }
      s_x[e] = x;
      i += step_i; d += step_d;
      if (d >= dim) { d -= dim; ++i; }
    }
  }
  __syncthreads();
  // ---- one thread per (row, d): all orders from one pass over the halo ----
  {
    int r = tid / dim, d = tid - r * dim;
    const int step_r = 256 / dim, step_d = 256 - step_r * dim;
    for (int e = tid; e < nrows * dim; e += 256) {
      const int i0 = r + halo, clo = s_lo[i0], chi = s_hi[i0];
      float acc[ORDER + 1];
#pragma unroll
      for (int o = 0; o <= ORDER; ++o) acc[o] = 0.0f;
      for (int j = -halo; j <= halo; ++j) {
        const int ii = min(max(i0 + j, clo), chi);
        const float x = s_x[ii * dim + d];
#pragma unroll
        for (int o = 0; o <= ORDER; ++o) {
          const int half = a.tap_half[o];
          if (j >= -half && j <= half) acc[o] = fmaf(s_taps[a.tap_off[o] + half + j], x, acc[o]);
        }
      }
      float *o_row = a.out + (row0 + r) * a.ld_out + d;
#pragma unroll
      for (int o = 0; o <= ORDER; ++o) o_row[o * dim] = acc[o];
      r += step_r; d += step_d;
      if (d >= dim) { d -= dim; ++r; }
    }
  }
}

// Compile-time (ORDER, WINDOW) variant for the reference's defaults (delta
// window 2, order 1 or 2): 128-row tiles, tap weights read as constant-bank
// operands, the utterance of a tile found by a CTA-wide two-round search
// instead of one thread's dependent binary search, and -- for the ~94 % of the
// tiles that lie inside one utterance together with their halo -- no per-row
// bounds, no clamping and one normalisation row.
constexpr int kFixedTile = 128;

// first u with offsets[u] <= row < offsets[u + 1]; every thread of the CTA
// must call it (256 probes per round, __syncthreads_count as the vote)
__device__ __forceinline__ int64_t cta_find_utt_row(const int64_t *offsets, int64_t nutts, int64_t row) {
  int64_t lo = 0, hi = nutts;
  while (hi - lo > 1) {
    const int64_t step = (hi - lo + 255) / 256;
    const int64_t probe = lo + (static_cast<int64_t>(threadIdx.x) + 1) * step;
    const int ok = (probe < hi) && (offsets[probe] <= row);
    const int k = __syncthreads_count(ok);
    lo += k * step;
    hi = min(lo + step, hi);
  }
  return lo;
}

template <int ORDER, int WINDOW>
__global__ void __launch_bounds__(256) delta_fixed_kernel(const DeltaArgs a) {
  constexpr int HALO = ORDER * WINDOW;
  constexpr int NLOAD = kFixedTile + 2 * HALO;
  extern __shared__ float s_x[];                        // [NLOAD, dim]
  __shared__ int s_lo[NLOAD], s_hi[NLOAD];              // clamp bounds (tile rows), mixed tiles only
  __shared__ int32_t s_group[NLOAD];
  const int tid = threadIdx.x;
  const int dim = a.dim;
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * kFixedTile;
  const int64_t lo = row0 - HALO;
  const int nrows = static_cast<int>(min(static_cast<int64_t>(kFixedTile), a.total_frames - row0));
  // the utterance of the tile's first row: try the equal-length guess first (one
  // round trip to L2 instead of the search's three when the corpus is uniform)
  const int64_t row_s = lo < 0 ? 0 : lo;
  int64_t u0 = static_cast<int64_t>(static_cast<double>(row_s) * static_cast<double>(a.nutts) /
                                    static_cast<double>(a.total_frames));
  u0 = min(max(u0, static_cast<int64_t>(0)), a.nutts - 1);
  int64_t first0 = a.frame_offsets[u0], next0 = a.frame_offsets[u0 + 1];
  if (!(first0 <= row_s && row_s < next0)) {              // (uniform over the CTA)
    u0 = cta_find_utt_row(a.frame_offsets, a.nutts, row_s);
    first0 = a.frame_offsets[u0]; next0 = a.frame_offsets[u0 + 1];
  }
  const bool uniform = first0 <= lo && lo + NLOAD <= next0;     // tile + halo inside utterance u0
  const bool do_norm = a.norm != nullptr;
  if (!uniform) {
    for (int i = tid; i < NLOAD; i += 256) {
      const int64_t row = lo + i;
      int rlo = 0, rhi = -1, g = 0;
      if (row >= 0 && row < a.total_frames) {
        int64_t u = u0, first = first0, next = next0;
        if (row >= next) {
          u = find_utt_row(a.frame_offsets, a.nutts, row);
          first = a.frame_offsets[u]; next = a.frame_offsets[u + 1];
        }
        rlo = static_cast<int>(max(first - lo, static_cast<int64_t>(0)));
        rhi = static_cast<int>(min(next - 1 - lo, static_cast<int64_t>(NLOAD - 1)));
        g = a.utt_group ? a.utt_group[u] : static_cast<int32_t>(u);
      }
      s_lo[i] = rlo; s_hi[i] = rhi; s_group[i] = g;
    }
    __syncthreads();
  }
  // Every thread owns ONE column d for the whole tile: 256 / dim rows per pass,
  // consecutive threads on consecutive elements of a row (coalesced), the few
  // threads beyond rows_pp * dim idle.  Nothing per element but the load, the
  // normalisation and the store: no index arithmetic, and in a uniform tile the
  // column's (scale, offset) sit in two registers.
  const int rows_pp = 256 / dim;
  const int tr = tid / dim, d = tid - tr * dim;
  const bool active = tr < rows_pp;
  // ---- stage the normalised rows ----
  if (active) {
    const float *src = a.in + (lo + tr) * a.ld_in + d;
    const int64_t src_step = static_cast<int64_t>(rows_pp) * a.ld_in;
    float *dst = s_x + tr * dim + d;
    if (uniform) {
      float scale = 1.0f, offset = 0.0f;
      if (do_norm) {
        const int32_t g0 = a.utt_group ? a.utt_group[u0] : static_cast<int32_t>(u0);
        const float *n = a.norm + static_cast<int64_t>(g0) * 2 * dim;
        offset = n[d]; scale = n[dim + d];
      }
      // batches of four independent loads: with one load in flight per thread
      // the kernel was latency bound at a quarter of the DRAM bandwidth
      int i = tr;
      for (; i + 3 * rows_pp < NLOAD; i += 4 * rows_pp) {
        float x0 = src[0], x1 = src[src_step], x2 = src[2 * src_step], x3 = src[3 * src_step];
        if (do_norm) {
          // ApplyCmvn: MulColsVec then AddVecToRows (two roundings)
          x0 = __fadd_rn(__fmul_rn(x0, scale), offset); x1 = __fadd_rn(__fmul_rn(x1, scale), offset);
          x2 = __fadd_rn(__fmul_rn(x2, scale), offset); x3 = __fadd_rn(__fmul_rn(x3, scale), offset);
        }
        dst[0] = x0; dst[rows_pp * dim] = x1; dst[2 * rows_pp * dim] = x2; dst[3 * rows_pp * dim] = x3;
        src += 4 * src_step;
        dst += 4 * rows_pp * dim;
      }
      for (; i < NLOAD; i += rows_pp) {
        float x = *src;
        if (do_norm) x = __fadd_rn(__fmul_rn(x, scale), offset);
        *dst = x;
        src += src_step;
        dst += rows_pp * dim;
      }
    } else {
      for (int i = tr; i < NLOAD; i += rows_pp) {
        float x = 0.0f;
        if (s_hi[i] >= s_lo[i]) {
          x = *src;
          if (do_norm) {
            const float *n = a.norm + static_cast<int64_t>(s_group[i]) * 2 * dim;
            x = __fadd_rn(__fmul_rn(x, n[dim + d]), n[d]);
          }
        }
        *dst = x;
        src += src_step;
        dst += rows_pp * dim;
      }
    }
  }
  __syncthreads();
  // ---- one thread per (row, d): all orders from one pass over the halo ----
  if (active) {
    float *o_row = a.out + (row0 + tr) * a.ld_out + d;
    const int64_t out_step = static_cast<int64_t>(rows_pp) * a.ld_out;
    const float *col = s_x + tr * dim + d;                // row r of the tile = staged row r + HALO
    for (int r = tr; r < nrows; r += rows_pp) {
      float x[2 * HALO + 1];
      if (uniform) {
#pragma unroll
        for (int j = 0; j <= 2 * HALO; ++j) x[j] = col[j * dim];
      } else {
        const int i0 = r + HALO, clo = s_lo[i0], chi = s_hi[i0];
#pragma unroll
        for (int j = 0; j <= 2 * HALO; ++j) x[j] = s_x[min(max(r + j, clo), chi) * dim + d];
      }
      int off = 0;                                  // taps of order o start at sum_{i<o} (2 i WINDOW + 1)
#pragma unroll
      for (int o = 0; o <= ORDER; ++o) {
        float acc = 0.0f;
#pragma unroll
        for (int j = -o * WINDOW; j <= o * WINDOW; ++j)
          acc = fmaf(a.taps[off + o * WINDOW + j], x[HALO + j], acc);
        o_row[o * dim] = acc;
        off += 2 * o * WINDOW + 1;
      }
      o_row += out_step;
      col += rows_pp * dim;
    }
  }
}

static int build_delta_taps(int order, int window, DeltaArgs *a) {
  if (order < 0 || order > 7) return set_error(SNB_ERR_UNSUPPORTED, "delta order must be in [0, 7]");
  if (window <= 0 || window >= 1000) return set_error(SNB_ERR_VALUE, "window must be in [1, 999]");
  std::vector<std::vector<float>> scales(order + 1);
  scales[0] = {1.0f};
  for (int i = 1; i <= order; ++i) {
    const std::vector<float> &prev = scales[i - 1];
    const int prev_offset = (static_cast<int>(prev.size()) - 1) / 2, cur_offset = prev_offset + window;
    std::vector<float> cur(prev.size() + 2 * window, 0.0f);
    float normalizer = 0.0f;
    for (int j = -window; j <= window; ++j) {
      normalizer += static_cast<float>(j * j);
      for (int k = -prev_offset; k <= prev_offset; ++k)
        cur[j + k + cur_offset] += static_cast<float>(j) * prev[k + prev_offset];
    }
    const float inv = static_cast<float>(1.0 / normalizer);
    for (float &v : cur) v *= inv;
    scales[i] = cur;
  }
  int off = 0;
  for (int i = 0; i <= order; ++i) {
    if (off + static_cast<int>(scales[i].size()) > kMaxTaps)
      return set_error(SNB_ERR_UNSUPPORTED, "delta window*order too large for the GPU path (taps > %d)", kMaxTaps);
    a->tap_off[i] = off;
    a->tap_half[i] = (static_cast<int>(scales[i].size()) - 1) / 2;
    for (float v : scales[i]) a->taps[off++] = v;
  }
  a->order = order;
  return SNB_OK;
}

// ---- CMVN -----------------------------------------------------------------
// one CTA per utterance, blockDim = (32 columns, 8 row slices); double sums of
// float products, fixed reduction order => bit-reproducible
__global__ void __launch_bounds__(256) cmvn_accumulate_kernel(
    const float *feats, int64_t ld, int dim, const int64_t *frame_offsets, const float *weights,
    double *utt_stats) {
  __shared__ double s_sum[8][32], s_sq[8][32], s_cnt[8];
  const int64_t u = blockIdx.x;
  const int64_t first = frame_offsets[u], last = frame_offsets[u + 1];
  const int x = threadIdx.x, y = threadIdx.y;
  double *stats = utt_stats + u * 2 * (dim + 1);
  for (int dc = 0; dc < dim; dc += 32) {
    const int d = dc + x;
    double sum = 0.0, sq = 0.0, cnt = 0.0;
    for (int64_t t = first + y; t < last; t += 8) {
      const float w = weights ? weights[t] : 1.0f;
      if (weights && w == 0.0f) continue;
      if (dc == 0 && x == 0) cnt += w;
      if (d < dim) {
        const float v = feats[t * ld + d];
        sum += static_cast<double>(__fmul_rn(v, w));
        sq += static_cast<double>(__fmul_rn(__fmul_rn(v, v), w));
      }
    }
    s_sum[y][x] = sum;
    s_sq[y][x] = sq;
    if (dc == 0 && x == 0) s_cnt[y] = cnt;
    __syncthreads();
    if (y == 0) {
      double a = 0.0, b = 0.0;
      for (int k = 0; k < 8; ++k) { a += s_sum[k][x]; b += s_sq[k][x]; }
      if (d < dim) { stats[d] = a; stats[(dim + 1) + d] = b; }
      if (dc == 0 && x == 0) {
        double c = 0.0;
        for (int k = 0; k < 8; ++k) c += s_cnt[k];
        stats[dim] = c;
        stats[(dim + 1) + dim] = 0.0;
      }
    }
    __syncthreads();
  }
}

// dim <= 128: the 256 threads are (256 / dim row slices) x (dim columns), so
// that (almost) every thread works whatever the dimension, and the rows come
// in batches of four independent loads (one load in flight per thread left the
// kernel latency bound).  Same arithmetic per element; the slices are added in
// a fixed order => bit-reproducible.
__global__ void __launch_bounds__(256) cmvn_accumulate_cols_kernel(
    const float *feats, int64_t ld, int dim, const int64_t *frame_offsets, const float *weights,
    double *utt_stats) {
  __shared__ double s_sum[256], s_sq[256], s_cnt[256];
  const int64_t u = blockIdx.x;
  const int64_t first = frame_offsets[u], last = frame_offsets[u + 1];
  const int tid = threadIdx.x;
  const int rows_pp = 256 / dim;
  const int tr = tid / dim, d = tid - tr * dim;
  double sum = 0.0, sq = 0.0, cnt = 0.0;
  if (tr < rows_pp) {
    const int64_t step = rows_pp;
    int64_t t = first + tr;
    const float *src = feats + t * ld + d;
    const int64_t sstep = step * ld;
    auto add = [&](float v, float w) {
      if (weights && w == 0.0f) return;
      cnt += w;
      sum += static_cast<double>(__fmul_rn(v, w));
      sq += static_cast<double>(__fmul_rn(__fmul_rn(v, v), w));
    };
    for (; t + 3 * step < last; t += 4 * step) {
      const float v0 = src[0], v1 = src[sstep], v2 = src[2 * sstep], v3 = src[3 * sstep];
      float w0 = 1.0f, w1 = 1.0f, w2 = 1.0f, w3 = 1.0f;
      if (weights) { w0 = weights[t]; w1 = weights[t + step]; w2 = weights[t + 2 * step]; w3 = weights[t + 3 * step]; }
      add(v0, w0); add(v1, w1); add(v2, w2); add(v3, w3);
      src += 4 * sstep;
    }
    for (; t < last; t += step) {
      add(*src, weights ? weights[t] : 1.0f);
      src += sstep;
    }
  }
  s_sum[tid] = sum; s_sq[tid] = sq; s_cnt[tid] = cnt;
  __syncthreads();
  if (tid < dim) {
    double a = 0.0, b = 0.0;
    for (int k = 0; k < rows_pp; ++k) { a += s_sum[k * dim + tid]; b += s_sq[k * dim + tid]; }
    double *stats = utt_stats + u * 2 * (dim + 1);
    stats[tid] = a; stats[(dim + 1) + tid] = b;
    if (tid == 0) {
      double c = 0.0;
      for (int k = 0; k < rows_pp; ++k) c += s_cnt[k * dim];       // column 0 of every slice
      stats[dim] = c;
      stats[(dim + 1) + dim] = 0.0;
    }
  }
}

__global__ void cmvn_reduce_groups_kernel(const double *utt_stats, int dim, const int64_t *group_ptr,
                                          const int64_t *group_utts, int64_t ngroups,
                                          double *group_stats) {
  const int64_t width = 2 * (dim + 1);
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= ngroups * width) return;
  const int64_t g = idx / width, e = idx - g * width;
  double acc = group_stats[idx];
  for (int64_t k = group_ptr[g]; k < group_ptr[g + 1]; ++k) acc += utt_stats[group_utts[k] * width + e];
  group_stats[idx] = acc;
}

__global__ void cmvn_norm_kernel(const double *stats, int64_t ngroups, int dim, int norm_vars,
                                 int reverse, float *norm) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= ngroups * dim) return;
  const int64_t g = idx / dim;
  const int d = static_cast<int>(idx - g * dim);
  const double *s = stats + g * 2 * (dim + 1);
  const double count = s[dim];
  float offset, scale;
  if (!(count >= 1.0)) {
    offset = scale = nanf("");
  } else {
    const double mean = s[d] / count;
    if (!norm_vars) {
      scale = 1.0f;
      // ApplyCmvn: offset.AddVec(-1.0 / count, mean_stats)
      offset = reverse ? static_cast<float>(mean)
                       : static_cast<float>(static_cast<double>(static_cast<float>(-1.0 / count)) * s[d]);
    } else {
      double var = s[(dim + 1) + d] / count - mean * mean;
      if (var < 1.0e-20) var = 1.0e-20;
      if (!reverse) {
        const double sc = 1.0 / sqrt(var);
        scale = static_cast<float>(sc);
        offset = static_cast<float>(-(mean * sc));
      } else {
        scale = static_cast<float>(sqrt(var));
        offset = static_cast<float>(mean);
      }
    }
  }
  norm[g * 2 * dim + d] = offset;
  norm[g * 2 * dim + dim + d] = scale;
}

// ---- sliding-window CMN ------------------------------------------------------
struct SlideArgs {
  const float *in;
  int64_t ld_in;
  int dim;
  const int64_t *frame_offsets;
  int64_t nutts, total_frames;
  int center, cmn_window, min_window, normalize_variance;
  float *out;
  int64_t ld_out;
};

__global__ void __launch_bounds__(256) sliding_cmn_kernel(const SlideArgs a) {
  __shared__ int64_t s_first[kRowsPerCta], s_last[kRowsPerCta];
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * kRowsPerCta;
  const int nrows = static_cast<int>(min(static_cast<int64_t>(kRowsPerCta), a.total_frames - row0));
  if (threadIdx.x < nrows) {
    const int64_t u = find_utt_row(a.frame_offsets, a.nutts, row0 + threadIdx.x);
    s_first[threadIdx.x] = a.frame_offsets[u];
    s_last[threadIdx.x] = a.frame_offsets[u + 1];
  }
  __syncthreads();
  const int dim = a.dim;
  for (int e = threadIdx.x; e < nrows * dim; e += blockDim.x) {
    const int r = e / dim, d = e - r * dim;
    const int64_t base = s_first[r], nf = s_last[r] - base;
    const int64_t t = row0 + r - base;
    int64_t ws, we;
    if (a.center) { ws = t - a.cmn_window / 2; we = ws + a.cmn_window; }
    else { ws = t - a.cmn_window; we = t + 1; }
    if (ws < 0) { we -= ws; ws = 0; }
    if (!a.center && we > t) we = max(t + 1, static_cast<int64_t>(a.min_window));
    if (we > nf) { ws -= (we - nf); we = nf; if (ws < 0) ws = 0; }
    double sum = 0.0, sumsq = 0.0;
    for (int64_t f = ws; f < we; ++f) {
      const double x = a.in[(base + f) * a.ld_in + d];
      sum += x;
      sumsq += x * x;
    }
    const double wf = static_cast<double>(we - ws);
    double y = static_cast<double>(a.in[(base + t) * a.ld_in + d]) - sum / wf;
    if (a.normalize_variance) {
      if (we - ws == 1) y = 0.0;
      else {
        double v = sumsq / wf - sum * sum / (wf * wf);
        if (v < 1.0e-10) v = 1.0e-10;
        y *= 1.0 / sqrt(v);
      }
    }
    a.out[(base + t) * a.ld_out + d] = static_cast<float>(y);
  }
}

// ---- energy VAD: one CTA per utterance --------------------------------------
__global__ void __launch_bounds__(256) vad_kernel(const float *feats, int64_t ld,
                                                  const int64_t *frame_offsets, float thr0,
                                                  float mean_scale, int context, float prop,
                                                  float *out) {
  __shared__ double s_part[256];
  const int64_t u = blockIdx.x;
  const int64_t first = frame_offsets[u], T = frame_offsets[u + 1] - first;
  if (T <= 0) return;
  float thr = thr0;
  if (mean_scale != 0.0f) {
    double part = 0.0;
    for (int64_t t = threadIdx.x; t < T; t += blockDim.x) part += feats[(first + t) * ld];
    s_part[threadIdx.x] = part;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (threadIdx.x < s) s_part[threadIdx.x] += s_part[threadIdx.x + s];
      __syncthreads();
    }
    const float sum = static_cast<float>(s_part[0]);
    thr = __fadd_rn(thr, __fdiv_rn(__fmul_rn(mean_scale, sum), static_cast<float>(T)));
  }
  for (int64_t t = threadIdx.x; t < T; t += blockDim.x) {
    int num = 0, den = 0;
    for (int64_t t2 = t - context; t2 <= t + context; ++t2)
      if (t2 >= 0 && t2 < T) {
        ++den;
        if (feats[(first + t2) * ld] > thr) ++num;
      }
    out[first + t] = (static_cast<float>(num) >= __fmul_rn(static_cast<float>(den), prop)) ? 1.0f : 0.0f;
  }
}

__global__ void convert_f64_f32_kernel(const double *in, float *out, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = static_cast<float>(in[i]);
}

}  // namespace snb

using namespace snb;

static unsigned row_ctas(int64_t total_frames) {
  return static_cast<unsigned>((total_frames + kRowsPerCta - 1) / kRowsPerCta);
}

extern "C" int snb_cmvn_apply_deltas(const float *d_in, int64_t ld_in, int32_t dim,
                                     const int64_t *d_frame_offsets, int64_t nutts, int64_t total_frames,
                                     const float *d_norm, const int32_t *d_utt_group, int32_t order,
                                     int32_t window, float *d_out, int64_t ld_out, void *stream) {
  if (total_frames == 0) return SNB_OK;
  if (!d_in || !d_out || !d_frame_offsets || dim <= 0) return set_error(SNB_ERR_VALUE, "bad argument");
  if (ld_in < dim || ld_out < static_cast<int64_t>(dim) * (order + 1))
    return set_error(SNB_ERR_VALUE, "leading dimension too small");
  DeltaArgs a;
  int rc = build_delta_taps(order, order > 0 ? window : 1, &a);
  if (rc != SNB_OK) return rc;
  a.in = d_in; a.ld_in = ld_in; a.dim = dim;
  a.frame_offsets = d_frame_offsets; a.nutts = nutts; a.total_frames = total_frames;
  a.norm = d_norm; a.utt_group = d_utt_group;
  a.out = d_out; a.ld_out = ld_out;
  const int halo = a.tap_half[a.order];
  const size_t smem = static_cast<size_t>(kTileRows + 2 * halo) * dim * sizeof(float);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static const bool no_fixed = getenv("SNB_DELTA_FIXED") && atoi(getenv("SNB_DELTA_FIXED")) == 0;
  if (window == 2 && (order == 1 || order == 2) && dim <= 80 && !no_fixed) {
    // (kFixedTile + 8) * 80 floats < 48 KB: no opt-in needed
    const unsigned ctas = static_cast<unsigned>((total_frames + kFixedTile - 1) / kFixedTile);
    const size_t smem_fixed = static_cast<size_t>(kFixedTile + 2 * halo) * dim * sizeof(float);
    if (order == 1) delta_fixed_kernel<1, 2><<<ctas, 256, smem_fixed, st>>>(a);
    else delta_fixed_kernel<2, 2><<<ctas, 256, smem_fixed, st>>>(a);
  } else if (order <= 3 && halo <= kMaxHalo && smem <= 96 * 1024 && dim <= 256) {
    const unsigned ctas = static_cast<unsigned>((total_frames + kTileRows - 1) / kTileRows);
    static std::atomic<size_t> cur[4] = {{48 * 1024}, {48 * 1024}, {48 * 1024}, {48 * 1024}};
    auto launch = [&](auto kernel, std::atomic<size_t> *state) -> int {
      size_t c = state->load();
      while (smem > c) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem));
        if (e != cudaSuccess) return set_error(SNB_ERR_CUDA, "delta smem: %s", cudaGetErrorString(e));
        if (state->compare_exchange_weak(c, smem)) break;
      }
      kernel<<<ctas, 256, smem, st>>>(a, halo);
      return SNB_OK;
    };
    int rc2 = SNB_OK;
    switch (order) {
      case 0: rc2 = launch(delta_tiled_kernel<0>, &cur[0]); break;
      case 1: rc2 = launch(delta_tiled_kernel<1>, &cur[1]); break;
      case 2: rc2 = launch(delta_tiled_kernel<2>, &cur[2]); break;
      default: rc2 = launch(delta_tiled_kernel<3>, &cur[3]); break;
    }
    if (rc2 != SNB_OK) return rc2;
  } else {
    delta_kernel<<<row_ctas(total_frames), 256, 0, st>>>(a);
  }
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_compute_deltas(const float *d_in, int64_t ld_in, int32_t dim,
                                  const int64_t *d_frame_offsets, int64_t nutts, int64_t total_frames,
                                  int32_t order, int32_t window, float *d_out, int64_t ld_out,
                                  void *stream) {
  return snb_cmvn_apply_deltas(d_in, ld_in, dim, d_frame_offsets, nutts, total_frames, nullptr, nullptr,
                               order, window, d_out, ld_out, stream);
}

extern "C" int snb_cmvn_apply(const float *d_in, int64_t ld_in, int32_t dim,
                              const int64_t *d_frame_offsets, int64_t nutts, int64_t total_frames,
                              const float *d_norm, const int32_t *d_utt_group, float *d_out,
                              int64_t ld_out, void *stream) {
  if (!d_norm) return set_error(SNB_ERR_VALUE, "d_norm is required");
  return snb_cmvn_apply_deltas(d_in, ld_in, dim, d_frame_offsets, nutts, total_frames, d_norm,
                               d_utt_group, 0, 1, d_out, ld_out, stream);
}

extern "C" int snb_cmvn_accumulate(const float *d_feats, int64_t ld, int32_t dim,
                                   const int64_t *d_frame_offsets, int64_t nutts, const float *d_weights,
                                   double *d_utt_stats, void *stream) {
  if (nutts == 0) return SNB_OK;
  // d_feats may be NULL when the batch holds no frame at all (the kernel then
  // only writes zero statistics)
  if (!d_utt_stats || !d_frame_offsets || dim <= 0 || ld < dim)
    return set_error(SNB_ERR_VALUE, "bad argument");
  static const bool old_kernel = getenv("SNB_CMVN_COLS") && atoi(getenv("SNB_CMVN_COLS")) == 0;
  if (dim <= 128 && !old_kernel)
    cmvn_accumulate_cols_kernel<<<static_cast<unsigned>(nutts), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        d_feats, ld, dim, d_frame_offsets, d_weights, d_utt_stats);
  else
    cmvn_accumulate_kernel<<<static_cast<unsigned>(nutts), dim3(32, 8), 0, static_cast<cudaStream_t>(stream)>>>(
        d_feats, ld, dim, d_frame_offsets, d_weights, d_utt_stats);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_cmvn_reduce_groups(const double *d_utt_stats, int32_t dim, const int64_t *d_group_ptr,
                                      const int64_t *d_group_utts, int64_t ngroups, double *d_group_stats,
                                      void *stream) {
  if (ngroups == 0) return SNB_OK;
  const int64_t n = ngroups * 2 * (dim + 1);
  cmvn_reduce_groups_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      d_utt_stats, dim, d_group_ptr, d_group_utts, ngroups, d_group_stats);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_cmvn_norm_from_stats(const double *d_stats, int64_t ngroups, int32_t dim,
                                        int32_t norm_vars, int32_t reverse, float *d_norm, void *stream) {
  if (ngroups == 0) return SNB_OK;
  const int64_t n = ngroups * dim;
  cmvn_norm_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      d_stats, ngroups, dim, norm_vars, reverse, d_norm);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_sliding_window_cmn(const float *d_in, int64_t ld_in, int32_t dim,
                                      const int64_t *d_frame_offsets, int64_t nutts, int64_t total_frames,
                                      int32_t center, int32_t cmn_window, int32_t min_window,
                                      int32_t normalize_variance, float *d_out, int64_t ld_out,
                                      void *stream) {
  if (total_frames == 0) return SNB_OK;
  if (cmn_window <= 0) return set_error(SNB_ERR_VALUE, "cmn_window must be positive");
  SlideArgs a{d_in, ld_in, dim, d_frame_offsets, nutts, total_frames, center, cmn_window, min_window,
              normalize_variance, d_out, ld_out};
  sliding_cmn_kernel<<<row_ctas(total_frames), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_vad_energy(const float *d_feats, int64_t ld, const int64_t *d_frame_offsets,
                              int64_t nutts, int64_t total_frames, float energy_threshold,
                              float energy_mean_scale, int32_t frames_context,
                              float proportion_threshold, float *d_out, void *stream) {
  if (nutts == 0 || total_frames == 0) return SNB_OK;
  vad_kernel<<<static_cast<unsigned>(nutts), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_feats, ld, d_frame_offsets, energy_threshold, energy_mean_scale, frames_context,
      proportion_threshold, d_out);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_convert_f64_to_f32(const double *d_in, float *d_out, int64_t n, void *stream) {
  if (n == 0) return SNB_OK;
  convert_f64_f32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      d_in, d_out, n);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
