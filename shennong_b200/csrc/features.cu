// features.cu -- the fused frame-features kernels (spectrogram / filterbank /
// MFCC / PLP / energy) and their plan / batch objects.
//
// Replaces, for a whole ragged batch in ONE launch, what the reference runs
// per utterance and per frame on one CPU thread:
//   OfflineFeatureTpl<F>::Compute -> ExtractWindow/ProcessWindow -> SRFFT ->
//   ComputePowerSpectrum -> MelBanks::Compute -> log -> DCT/lifter | PLP tail
// (shennong/processor/base.py:427-431, spectrogram.py:137-140,
//  plp.py:510-626, energy.py:168-183).
//
// Fast path (padded FFT size 512, i.e. 16 kHz / 25 ms and every window of
// 257..512 samples): persistent CTAs, 256 threads = 16 half-warp "groups";
// each CTA stages the contiguous int16 PCM span of a tile of <= 32 frames into
// shared memory with one TMA bulk copy (cp.async.bulk + mbarrier), each group
// owns one frame: dither / DC removal / log-energy / pre-emphasis / window in
// registers, the 512-point real FFT as a 256-point complex FFT factored
// 16 x 16 (two in-register radix-16 butterflies, one padded shared-memory
// transpose), the real-FFT unpack with half-warp shuffles (each conjugate
// pair computed once), then warp-local mel / log / DCT / PLP tails.
// Generic path (any other FFT size, incl. non powers of two): one warp per
// frame, shared-memory radix-2 FFT or direct DFT, same tails.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

#include "device_utils.cuh"
#include "snb_internal.h"

namespace snb {

__device__ __forceinline__ int64_t first_sample_of_frame_dev(int64_t frame, const FeatParams &p) {
  if (p.fo.snip_edges) return frame * p.S;
  return frame * p.S + p.S / 2 - p.W / 2;
}

// ---------------------------------------------------------------------------
// shared-memory views
// ---------------------------------------------------------------------------
struct MelView {          // decoded mel blob (in shared or global memory)
  const float *loudness;
  // generic path (per-triangle ranges)
  const int32_t *first, *size, *offset;
  const float *weights;
  // fused path (see "mel energies" in feature_tail)
  const float *chunk_w;         // [32][20]: (up, down) of the lane's 8 bins, 4 floats of padding
  const uint32_t *chunk_meta;   // [32]: bits 0-7 "bin closes a run", bits 8+ first run of the lane
  const int32_t *run_first;     // [B+2]: first run of every segment
};

__device__ __forceinline__ MelView mel_view(const int32_t *blob, const FeatParams &p) {
  const int B = p.B;
  MelView v;
  v.loudness = reinterpret_cast<const float *>(blob);
  v.chunk_w = reinterpret_cast<const float *>(blob + p.mel_chunk_off);
  v.chunk_meta = reinterpret_cast<const uint32_t *>(blob + p.mel_meta_off);
  v.run_first = blob + p.mel_run_off;
  v.first = blob + p.mel_gen_off;
  v.size = v.first + B;
  v.offset = v.first + 2 * B;
  v.weights = reinterpret_cast<const float *>(v.first + 3 * B);
  return v;
}

struct TailTables {       // tables the tails read (shared memory in both paths)
  const float *dct, *lifter, *idft;
  int dct_stride;         // floats between DCT rows (fast path: padded, see fast_layout)
  // fused path only: the warp's run partials, and the distance between the
  // buffers of the two lane groups (= frames) of a warp
  float4 *part;
  int grp_floats;
  int mel_max_runs;       // longest run list of a segment (over the VTLN warps seen so far: plan-wide bound)
  MelView mel;
};

// Power-spectrum layout.  Generic path: natural order.  Fused path: bin k at
// (k / 8) * 12 + k % 8, so that a lane reading its 8 consecutive bins with two
// 128-bit loads (lane stride 12 words) never shares a bank with the other
// seven lanes of its quarter warp.
constexpr int kPStride = 12;
constexpr int kPFloats = 32 * kPStride + 4;      // bins 0..256 (+ padding): scratch starts here
template <int G>
__device__ __forceinline__ int pidx(int k) {
  return (G == 16) ? (k >> 3) * kPStride + (k & 7) : k;
}

// PLP after the (loudness-weighted, compressed) mel energies are in
// mel[1..B] = scratch[1..B] (plp.py:590-626): duplicate the ends, IDFT to the
// autocorrelation, lane-parallel Levinson-Durbin and LPC -> cepstrum, lifter,
// scale, energy, HTK reorder.
template <int G>
__device__ __forceinline__ void plp_finish(const FeatParams &p, const TailTables &t, float *scratch,
                                           float log_energy, float *out_row, bool valid, int gl) {
  const snb_feat_opts &xo = p.xo;
  const int B = p.B, nc = xo.num_ceps;
  float *mel = scratch;
  // ---- PLP tail (plp.py:590-626) ----
  const int L = xo.lpc_order;       // <= G - 1 (checked at plan creation)
  if (gl == 0) { mel[0] = mel[1]; mel[B + 1] = mel[B]; }
  __syncwarp();
  float *ac = scratch + (B + 2);    // [L+1] (useg/dseg follow at B+2 + L+2)
  for (int i = gl; i <= L; i += G) {
    const float *row = t.idft + i * (B + 2);
    float acc = 0.0f;
    for (int j = 0; j < B + 2; ++j) acc = fmaf(row[j], mel[j], acc);
    ac[i] = acc;
  }
  __syncwarp();
  // Levinson-Durbin, lanes parallel over the coefficient index j
  const int gbase = (threadIdx.x & 31) & ~(G - 1);  // first lane of my group
  float E = ac[0];
  float lpc = 0.0f;                 // lane j holds lpc[j]
  for (int i = 0; i < L; ++i) {
    // ki = (ac[i+1] + sum_{j<i} lpc[j] * ac[i-j]) / E
    float part = (gl < i) ? lpc * ac[i - gl] : 0.0f;
    part = group_sum<G>(part);
    float ki = (ac[i + 1] + part) / E;
    float c = 1.0f - ki * ki;
    if (c < 1.0e-5f) c = 1.0e-5f;
    E *= c;
    // lpc'[j] = lpc[j] - ki * lpc[i-j-1] (j<i); lpc'[i] = -ki
    const int src = (i - gl - 1) & (G - 1);
    const float other = __shfl_sync(SNB_FULL_MASK, lpc, gbase + src);
    if (gl < i) lpc = lpc - ki * other;
    else if (gl == i) lpc = -ki;
  }
  // ComputeLpc returns -log(1/E); plp.py:603 floors with float64 eps
  float residual = -logf(1.0f / E);
  residual = fmaxf(residual, 2.220446049250313e-16f);
  // LPC -> cepstrum (plp.py:164-168), lane i holds cep[i]
  float cep = 0.0f;
  for (int i = 0; i < L; ++i) {
    // sum_{j<i} (i-j) * lpc[j] * cep[i-j-1]
    const int src = (i - gl - 1) & (G - 1);
    const float cj = __shfl_sync(SNB_FULL_MASK, cep, gbase + src);
    float part = (gl < i) ? static_cast<float>(i - gl) * lpc * cj : 0.0f;
    part = group_sum<G>(part);
    const float mine = -lpc - part / static_cast<float>(i + 1);
    if (gl == i) cep = mine;
  }
  // out[0] = residual, out[c] = cep[c-1]; lifter, scale, energy, htk reorder
  // (nc <= L + 1 <= G: one column per lane; the shuffle is warp-uniform)
  const float prev = __shfl_sync(SNB_FULL_MASK, cep, gbase + ((gl - 1) & (G - 1)));
  if (gl < nc) {
    const int c = gl;
    float v = (c == 0) ? residual : prev;
    if (xo.cepstral_lifter != 0.0f) v *= t.lifter[c];
    if (xo.cepstral_scale != 1.0f) v *= xo.cepstral_scale;
    if (c == 0 && xo.use_energy) v = log_energy;
    int col = c;
    if (xo.htk_compat) col = (c == 0) ? nc - 1 : c - 1;
    if (valid) out_row[col] = v;
  }
}

// ---------------------------------------------------------------------------
// tails: from the power spectrum P[0..N/2] (shared memory, private to the
// lane group) to one output row.  G = lanes per frame (16 or 32); all lanes of
// the warp execute this in lockstep, `valid` only gates the global stores.
// scratch: >= B + 2 + lpc_order + 1 floats private to the group.
// ---------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ void feature_tail(const FeatParams &p, const TailTables &t,
                                             float *P, float *scratch, float log_energy,
                                             float *out_row, bool valid, int gl) {
  const snb_feat_opts &xo = p.xo;
  const int half = p.N / 2;
  if (xo.energy_floor > 0.0f && log_energy < p.log_energy_floor)
    log_energy = p.log_energy_floor;

  if (xo.kind == SNB_FEAT_SPECTROGRAM) {
    for (int k = gl; k <= half; k += G) {
      float v = logf(fmaxf(P[pidx<G>(k)], FLT_EPSILON));
      if (k == 0) v = log_energy;
      if (valid) out_row[k] = v;
    }
    return;
  }
  const int B = p.B;
  if (xo.kind == SNB_FEAT_FBANK && !xo.use_power) {
    for (int k = gl; k <= half; k += G) P[pidx<G>(k)] = sqrtf(P[pidx<G>(k)]);
    __syncwarp();
  }
  // mel energies
  float *mel = scratch;  // [B+2]; PLP uses mel[1..B] with duplicated ends
  const int moff = (xo.kind == SNB_FEAT_PLP) ? 1 : 0;
  if (G == 16) {
    // Both frames of the warp at once: lane l (0..31) owns the FFT bins
    // [8 l, 8 l + 8) of the two power spectra.  Every bin belongs to one
    // segment s (the bins between two consecutive triangle centres) and feeds
    // the rising side of triangle s (weight up) and the falling side of
    // triangle s - 1 (weight down).  The lane accumulates (up, down) sums of
    // both frames along a RUN of bins of one segment, and drops the four sums
    // into the warp's partial table when the run ends (host-built flags;
    // runs end at segment and at chunk boundaries).  Weights and spectra come
    // in with conflict-free 128-bit loads, nothing is read twice.
    const int lane32 = threadIdx.x & 31;
    float *P0 = P - (lane32 >> 4) * t.grp_floats;           // spectrum of the warp's first frame
    const float4 *wv = reinterpret_cast<const float4 *>(t.mel.chunk_w + 20 * lane32);
    const float4 *pa = reinterpret_cast<const float4 *>(P0 + kPStride * lane32);
    const float4 *pb = reinterpret_cast<const float4 *>(P0 + t.grp_floats + kPStride * lane32);
    const uint32_t meta = t.mel.chunk_meta[lane32];
    float4 *part = t.part + (meta >> 8);
    float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#define SNB_MEL_BIN(bit, up, dn, xa, xb)                                          \
    acc.x = fmaf(up, xa, acc.x); acc.y = fmaf(dn, xa, acc.y);                       \
    acc.z = fmaf(up, xb, acc.z); acc.w = fmaf(dn, xb, acc.w);                       \
    if (meta & (1u << (bit))) { *part++ = acc; acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f); }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 xa = pa[h], xb = pb[h];
      const float4 w01 = wv[2 * h], w23 = wv[2 * h + 1];
      SNB_MEL_BIN(4 * h + 0, w01.x, w01.y, xa.x, xb.x)
      SNB_MEL_BIN(4 * h + 1, w01.z, w01.w, xa.y, xb.y)
      SNB_MEL_BIN(4 * h + 2, w23.x, w23.y, xa.z, xb.z)
      SNB_MEL_BIN(4 * h + 3, w23.z, w23.w, xa.w, xb.w)
    }
#undef SNB_MEL_BIN
    __syncwarp();
  }
  for (int b = gl; b < B; b += G) {
    float acc = 0.0f;
    if (G == 16) {
      // triangle b = rising sums of segment b + falling sums of segment b + 1,
      // runs added in bin order (deterministic)
      const float2 *mine = reinterpret_cast<const float2 *>(t.part) + ((threadIdx.x & 31) >> 4);
      const int r0 = t.mel.run_first[b], r1 = t.mel.run_first[b + 1], r2 = t.mel.run_first[b + 2];
      // (blocks of four predicated steps up to the longest run list of the
      // plan: uniform trip count, no divergent loop per lane)
      float u = 0.0f, d = 0.0f;
      for (int j0 = 0; j0 < t.mel_max_runs; j0 += 4) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ju = r0 + j0 + i, jd = r1 + j0 + i;
          if (ju < r1) u += mine[2 * ju].x;
          if (jd < r2) d += mine[2 * jd].y;
        }
      }
      acc = u + d;
    } else {
      const int first = t.mel.first[b], size = t.mel.size[b];
      const float *w = t.mel.weights + t.mel.offset[b];
      for (int i = 0; i < size; ++i) acc = fmaf(w[i], P[first + i], acc);
    }
    if (xo.kind == SNB_FEAT_FBANK) {
      if (xo.use_log_fbank) acc = logf(fmaxf(acc, FLT_EPSILON));
      const int off = (xo.use_energy && !xo.htk_compat) ? 1 : 0;
      if (valid) out_row[off + b] = acc;
    } else if (xo.kind == SNB_FEAT_MFCC) {
      mel[b] = logf(fmaxf(acc, FLT_EPSILON));
    } else {  // PLP: loudness, compression (plp.py:587-588)
      mel[moff + b] = powf(acc * t.mel.loudness[b], xo.compress_factor);
    }
  }
  if (xo.kind == SNB_FEAT_FBANK) {
    if (xo.use_energy && gl == 0 && valid)
      out_row[xo.htk_compat ? B : 0] = log_energy;
    return;
  }
  const int nc = xo.num_ceps;
  if (G == 16 && xo.kind == SNB_FEAT_MFCC && B + gl < ((B + 3) & ~3)) mel[B + gl] = 0.0f;
  __syncwarp();
  if (xo.kind == SNB_FEAT_MFCC) {
    for (int c = gl; c < nc; c += G) {
      const float *row = t.dct + c * t.dct_stride;
      float acc = 0.0f;
      if (G == 16) {
        // rows and mel[] are zero padded to a multiple of four: 128-bit loads,
        // same summation order as the scalar loop (the padding adds +0 terms)
        const float4 *row4 = reinterpret_cast<const float4 *>(row);
        const float4 *mel4 = reinterpret_cast<const float4 *>(mel);
        for (int n4 = 0; n4 < (B + 3) / 4; ++n4) {
          const float4 r = row4[n4], m = mel4[n4];
          acc = fmaf(r.x, m.x, acc); acc = fmaf(r.y, m.y, acc);
          acc = fmaf(r.z, m.z, acc); acc = fmaf(r.w, m.w, acc);
        }
      } else {
        for (int n = 0; n < B; ++n) acc = fmaf(row[n], mel[n], acc);
      }
      if (xo.cepstral_lifter != 0.0f) acc *= t.lifter[c];
      if (c == 0 && xo.use_energy) acc = log_energy;
      int col = c;
      if (xo.htk_compat) {
        if (c == 0) {
          col = nc - 1;
          if (!xo.use_energy) acc *= 1.41421356237309504880f;
        } else {
          col = c - 1;
        }
      }
      if (valid) out_row[col] = acc;
    }
    return;
  }
  plp_finish<G>(p, t, scratch, log_energy, out_row, valid, gl);
}

// energy kind (energy.py:171-183): float64 sum of squares of the processed
// window, compressed; one float64 per frame
__device__ __forceinline__ double compress_energy(double e, int mode) {
  if (e < DBL_MIN) e = DBL_MIN;
  if (mode == 1) return log(e);
  if (mode == 2) return sqrt(e);
  return e;
}

// ---------------------------------------------------------------------------
// fast path: N = 512
// ---------------------------------------------------------------------------
// packed fp32 (FADD2/FMUL2/FFMA2) per stage of the fused kernel; compile-time
// switches so that each stage's gain can be measured separately
#ifndef SNB_PACK_EW
#define SNB_PACK_EW 1       // load / DC / energies / window
#endif
#ifndef SNB_PACK_FFT
#define SNB_PACK_FFT 1      // radix-16 butterflies
#endif
#ifndef SNB_PACK_UNPACK
#define SNB_PACK_UNPACK 1   // real-FFT unpack
#endif
constexpr bool kPackEw = SNB_PACK_EW != 0, kPackFft = SNB_PACK_FFT != 0, kPackUn = SNB_PACK_UNPACK != 0;
constexpr int kFastThreads = 256;
constexpr int kFastGroups = 16;         // half-warps per CTA
constexpr int kXStride = 17;            // float2 row stride of the transpose buffer

struct FastSmemLayout {                 // byte offsets into dynamic smem
  int window, tw1, tw2, dct, lifter, idft, mel, pcm, grp, part, desc, bar, total;
  int grp_floats, span_cap, dct_stride, part_floats;
};

// two int16 samples packed in one 32-bit word -> two floats without the
// quarter-rate I2F: splice each half (biased by 0x8000) into the mantissa of
// 2^23 and subtract 2^23 + 2^15; exact for every int16 value
__device__ __forceinline__ f32x2 s16x2_to_f32x2(uint32_t pr) {
  const uint32_t u = pr ^ 0x80008000u;
  return sub2t<kPackEw>(pk(__uint_as_float(__byte_perm(u, 0x4B000000u, 0x7410)),
                 __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7432))),
              pk(8421376.0f, 8421376.0f));
}

struct FastArgs {
  FeatParams p;
  FastSmemLayout sm;
  const TileDesc *tiles;
  int64_t ntiles;
  const int64_t *sample_begin;
  const int64_t *sample_len;
  const int64_t *frame_offsets;
  const int32_t *mel_blobs;
  const int16_t *pcm;
  int64_t total_samples;
  void *out;
  int64_t ld_out;
  uint64_t seed;
  int use_tma;
  int mel_max_runs;
};

// the 16-byte aligned piece of the packed PCM buffer that covers a tile;
// false when the tile has to be gathered element by element
struct TileSpan {
  int64_t gstart;      // first sample of the bulk copy (multiple of 8)
  uint32_t bytes;      // multiple of 16
  int mis;             // samples between gstart and the tile's first sample
};
__device__ __forceinline__ bool tile_span(const TileDesc &td, const FastArgs &a, int S, int W, TileSpan *out) {
  if (!a.use_tma || td.g0 < 0) return false;
  const int span = (td.nf - 1) * S + W;
  out->mis = static_cast<int>(td.g0 & 7);
  out->gstart = td.g0 - out->mis;
  const int64_t nsamp = (static_cast<int64_t>(span) + out->mis + 7) & ~7ll;
  out->bytes = static_cast<uint32_t>(nsamp * 2);
  return out->gstart + nsamp <= a.total_samples;
}

// kMinBlocks = 3: 80 registers; kMinBlocks = 4: 64 registers (no spills), 32
// resident warps per SM -- selected at plan time when the shared memory of four
// CTAs fits (SNB_FUSED_OCC=3|4 overrides).
// kW > 0: the window length is a compile-time constant (400 = 25 ms at 16 kHz,
// every BASELINE configuration): the per-element "is this sample inside the
// window" tests of the unrolled stages fold away and the stages stop at the
// last register that can hold a sample.  kW = 0: any window of 257..512.
template <int kMinBlocks, int kW>
__global__ void __launch_bounds__(kFastThreads, kMinBlocks)
fused_features_512_kernel(const FastArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const FeatParams &p = a.p;
  float *s_window = reinterpret_cast<float *>(smem + a.sm.window);
  float2 *s_tw1 = reinterpret_cast<float2 *>(smem + a.sm.tw1);   // [k1*16+hl] W256^(hl*k1)
  float2 *s_tw2 = reinterpret_cast<float2 *>(smem + a.sm.tw2);   // [k2*16+hl] W512^(hl+16k2)
  float *s_dct = reinterpret_cast<float *>(smem + a.sm.dct);
  float *s_lifter = reinterpret_cast<float *>(smem + a.sm.lifter);
  float *s_idft = reinterpret_cast<float *>(smem + a.sm.idft);
  int32_t *s_mel = reinterpret_cast<int32_t *>(smem + a.sm.mel);
  int16_t *s_pcm = reinterpret_cast<int16_t *>(smem + a.sm.pcm);
  float *s_grp_all = reinterpret_cast<float *>(smem + a.sm.grp);
  uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem + a.sm.bar);          // [2]
  TileDesc *s_desc = reinterpret_cast<TileDesc *>(smem + a.sm.desc);        // ring of 3

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int hl = lane & 15;                       // lane within the group
  const int grp = (tid >> 5) * 2 + (lane >> 4);   // 0..15
  const int gbase = lane & 16;                    // first lane of my group
  float *s_grp = s_grp_all + grp * a.sm.grp_floats;

  const int W = (kW > 0) ? kW : p.W;
  const int S = p.S, B = p.B;
  constexpr int kN1 = (kW > 0) ? (kW + 31) / 32 : 16;   // registers that can hold samples
  const snb_feat_opts &xo = p.xo;
  const int kind = xo.kind;

  // ---- one-time table load ----
  for (int i = tid; i < 512; i += kFastThreads) s_window[i] = (i < W) ? 0.5f * p.t.window[i] : 0.0f;   // see "window" below
  for (int i = tid; i < 256; i += kFastThreads) {
    const int k1 = i >> 4, l = i & 15;
    s_tw1[i] = p.t.tw_half[(l * k1) & 255];
    if (i < 128) s_tw2[i] = p.t.tw_full[l + 16 * k1];   // k2 = k1 < 8
  }
  if (kind == SNB_FEAT_MFCC)
    for (int i = tid; i < xo.num_ceps * a.sm.dct_stride; i += kFastThreads) {
      const int c = i / a.sm.dct_stride, n = i - c * a.sm.dct_stride;
      s_dct[i] = (n < B) ? p.t.dct[c * B + n] : 0.0f;
    }
  if (kind == SNB_FEAT_MFCC || kind == SNB_FEAT_PLP)
    for (int i = tid; i < xo.num_ceps; i += kFastThreads) s_lifter[i] = p.t.lifter[i];
  if (kind == SNB_FEAT_PLP)
    for (int i = tid; i < (xo.lpc_order + 1) * (B + 2); i += kFastThreads) s_idft[i] = p.t.idft[i];
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    fence_barrier_init();
  }
  int cur_mel = -1;
  uint32_t parity = 0;               // bit b: phase of the mbarrier of PCM buffer b
  TailTables tt;
  tt.dct = s_dct; tt.lifter = s_lifter; tt.idft = s_idft;
  tt.dct_stride = a.sm.dct_stride;
  tt.mel = mel_view(s_mel, p);
  tt.part = reinterpret_cast<float4 *>(smem + a.sm.part) + (tid >> 5) * (a.sm.part_floats / 4);
  tt.grp_floats = a.sm.grp_floats;
  tt.mel_max_runs = a.mel_max_runs;

  const bool pair_ok_static = (S % 2) == 0;
  const float dither = p.fo.dither;
  const int nfull = W / 32;           // n1 iterations with 32 valid samples

  // Software pipeline over the CTA's tiles (tile = up to 16 frames of one
  // utterance, one frame per lane group): while tile i is computed, the PCM
  // span of tile i+1 is already in flight (TMA bulk copy into the other
  // buffer, completion on that buffer's mbarrier) and the descriptor of tile
  // i+2 is being fetched by cp.async into a ring of three.  One CTA barrier
  // per tile.
  const int64_t stride = gridDim.x;
  if (tid == 0 && blockIdx.x < a.ntiles) {
    s_desc[0] = a.tiles[blockIdx.x];
    if (blockIdx.x + stride < a.ntiles) s_desc[1] = a.tiles[blockIdx.x + stride];
    TileSpan sp;
    if (tile_span(s_desc[0], a, S, W, &sp)) {
      fence_proxy_async();
      mbar_arrive_expect_tx(&s_bar[0], sp.bytes);
      bulk_copy_g2s(s_pcm, a.pcm + sp.gstart, sp.bytes, &s_bar[0]);
    }
  }
  __syncthreads();                   // tables, barriers and the first descriptors are visible
  int slot = 0, buf = 0;
  for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += stride) {
    const TileDesc td = s_desc[slot];
    const int slot1 = (slot == 2) ? 0 : slot + 1, slot2 = (slot1 == 2) ? 0 : slot1 + 1;
    if (tid == 0) {
      // buffer buf^1 and ring slot slot2 were last read before the barrier
      // that ended the previous iteration
      if (tile + stride < a.ntiles) {
        TileSpan sp;
        if (tile_span(s_desc[slot1], a, S, W, &sp)) {
          fence_proxy_async();
          mbar_arrive_expect_tx(&s_bar[buf ^ 1], sp.bytes);
          bulk_copy_g2s(s_pcm + (buf ^ 1) * a.sm.span_cap, a.pcm + sp.gstart, sp.bytes, &s_bar[buf ^ 1]);
        }
      }
      if (tile + 2 * stride < a.ntiles) {
        const TileDesc *src = a.tiles + (tile + 2 * stride);
        cp_async_16(&s_desc[slot2], src);
        cp_async_16(reinterpret_cast<char *>(&s_desc[slot2]) + 16, reinterpret_cast<const char *>(src) + 16);
        cp_async_commit();
      }
    }
    if (B > 0 && td.mel_idx != cur_mel) {
      // (the barrier that ended the previous iteration freed s_mel)
      const int32_t *src = a.mel_blobs + static_cast<int64_t>(td.mel_idx) * p.mel_blob_stride;
      for (int i = tid; i < p.mel_fast_words; i += kFastThreads) s_mel[i] = src[i];
      cur_mel = td.mel_idx;
      __syncthreads();
    }
    // ---- the PCM span of this tile ----
    int16_t *pcm_buf = s_pcm + buf * a.sm.span_cap;
    TileSpan sp;
    const bool tma = tile_span(td, a, S, W, &sp);
    const int mis = tma ? sp.mis : 0;              // samples of misalignment
    if (tma) {
      mbar_wait(&s_bar[buf], (parity >> buf) & 1u);
      parity ^= 1u << buf;
    } else {
      // element loads with reflection at the utterance edges (snip_edges=False,
      // ExtractWindow restated at plp.py:239-254) or spans that cannot be
      // bulk-copied (unaligned buffer, end of the allocation)
      const int64_t utt_len = a.sample_len[td.utt];
      const int16_t *src = a.pcm + a.sample_begin[td.utt];
      const int64_t a0 = first_sample_of_frame_dev(td.f0, p);      // may be < 0
      const int span = (td.nf - 1) * S + W;
      for (int i = tid; i < span; i += kFastThreads) {
        int64_t k = a0 + i;
        while (k < 0 || k >= utt_len) k = (k < 0) ? (-k - 1) : (2 * utt_len - 1 - k);
        pcm_buf[i] = src[k];
      }
      __syncthreads();
    }

    const int64_t row0 = td.row0;
    do {                             // one frame per lane group (`break` = frame done)
      const int fl = grp;
      const bool valid = fl < td.nf;
      const int fidx = valid ? fl : td.nf - 1;     // idle groups redo the last frame
      const int16_t *fr = pcm_buf + mis + fidx * S;
      const bool pair_ok = pair_ok_static && ((mis & 1) == 0);
      // Lane hl holds the samples (2p, 2p+1), p = 16 n1 + hl, as ONE packed pair
      // x[n1] = (re, im) of z[p] = s[2p] + i s[2p+1]: every stage that treats the
      // two halves alike is a single FADD2 / FMUL2 / FFMA2 per register pair.
      f32x2 x[16];
      // ---- load + int16 -> float (+ dither) ----
      // n1 < nfull: every lane holds two valid samples; n1 == nfull: the
      // ragged tail (W % 32 samples); beyond: zero padding up to 512
      f32x2 lsum2 = pk(0.0f, 0.0f);
      __syncwarp();                 // the previous pass is done with this group's buffer
      float2 *s_noise = reinterpret_cast<float2 *>(s_grp);   // [256] (dither * N(0,1)) pairs
      if (dither != 0.0f) {
        // rolled on purpose: ONE copy of the hash + Box-Muller code (the fully
        // unrolled form put ~11 KB of straight-line SASS in the hot path and
        // made the pass overflow the 32 KB instruction cache); every lane reads
        // back only what it wrote, so no synchronisation is needed
        const uint32_t key = frame_noise_key(a.seed, static_cast<uint64_t>(row0 + fidx));
        const int npair = (W + 1) / 2;
        const float c = -1.3862943611198906f * dither * dither;     // -2 ln 2 dither^2
#pragma unroll 2
        for (int pidx = hl; pidx < npair; pidx += 16) s_noise[pidx] = dither_pair(key, pidx, c);
      }
      // The usual case -- 32-bit aligned frame, even window length: a lane's two
      // samples are inside or outside the window together -- gets straight-line
      // code per dither setting (the tests on pair_ok / dither are hoisted out of
      // the unrolled loop); anything else takes the element-wise loop.
#define SNB_LOAD_PAIRS(WITH_NOISE)                                                        \
      _Pragma("unroll")                                                                   \
      for (int n1 = 0; n1 < 16; ++n1) {                                                   \
        const int i0 = 2 * (16 * n1 + hl);                                                \
        f32x2 v = pk(0.0f, 0.0f);                                                         \
        if (n1 < kN1 && (n1 < nfull || (n1 == nfull && i0 < W))) {                        \
          v = s16x2_to_f32x2(*reinterpret_cast<const uint32_t *>(fr + i0));               \
          if (WITH_NOISE) v = add2t<kPackEw>(v, *reinterpret_cast<const f32x2 *>(&s_noise[16 * n1 + hl])); \
        }                                                                                 \
        x[n1] = v;                                                                        \
        lsum2 = add2t<kPackEw>(lsum2, v);                                                 \
      }
      if (pair_ok && (W & 1) == 0) {
        if (dither != 0.0f) { SNB_LOAD_PAIRS(true) } else { SNB_LOAD_PAIRS(false) }
      } else {
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
          const int i0 = 2 * (16 * n1 + hl);
          float v0 = 0.0f, v1 = 0.0f;
          if (n1 < kN1 && n1 <= nfull) {
            if (i0 < W) v0 = static_cast<float>(fr[i0]);
            if (i0 + 1 < W) v1 = static_cast<float>(fr[i0 + 1]);
            if (dither != 0.0f && i0 < W) {
              const float2 nz = s_noise[16 * n1 + hl];
              v0 += nz.x;
              if (i0 + 1 < W) v1 += nz.y;
            }
          }
          x[n1] = pk(v0, v1);
          lsum2 = add2t<kPackEw>(lsum2, x[n1]);
        }
      }
#undef SNB_LOAD_PAIRS
      // ---- DC removal (ProcessWindow) ----
      if (p.fo.remove_dc_offset) {
        float l0, l1;
        upk(lsum2, l0, l1);
        const float mean = __fdiv_rn(group_sum<16>(l0 + l1), static_cast<float>(W));
        const f32x2 mean2 = pk(mean, mean);
#pragma unroll
        for (int n1 = 0; n1 < kN1; ++n1) {
          if (n1 < nfull) {
            x[n1] = sub2t<kPackEw>(x[n1], mean2);
          } else if (n1 == nfull) {
            const int i0 = 2 * (16 * n1 + hl);
            float r, i;
            upk(x[n1], r, i);
            if (i0 < W) r -= mean;
            if (i0 + 1 < W) i -= mean;
            x[n1] = pk(r, i);
          }
        }
      }
      // ---- raw log-energy / float64 energy ----
      float log_energy = 0.0f;
      if (p.need_raw_energy) {
        f32x2 e2 = pk(0.0f, 0.0f);
#pragma unroll
        for (int n1 = 0; n1 < kN1; ++n1) e2 = fma2t<kPackEw>(x[n1], x[n1], e2);
        float e0, e1;
        upk(e2, e0, e1);
        log_energy = logf(fmaxf(group_sum<16>(e0 + e1), p.eps_energy));
      }
      // ---- pre-emphasis (needs the ORIGINAL previous sample) ----
      if (p.fo.preemph_coeff != 0.0f) {
        const float c = p.fo.preemph_coeff;
#pragma unroll
        for (int n1 = kN1 - 1; n1 >= 0; --n1) {
          float r, i, rb, ib;
          upk(x[n1], r, i);
          upk(x[(n1 + 15) & 15], rb, ib);
          const float send = (hl == 15) ? ib : i;
          float prev = __shfl_sync(SNB_FULL_MASK, send, gbase | ((hl + 15) & 15));
          if (n1 == 0 && hl == 0) prev = r;
          x[n1] = pk(fmaf(-c, prev, r), fmaf(-c, r, i));
        }
      }
      // ---- window (zero beyond W: also clears pre-emphasis spill).  The table
      //      holds HALF the window: the real-FFT unpack then needs no 1/4 ----
#pragma unroll
      for (int n1 = 0; n1 < kN1; ++n1)
        x[n1] = mul2t<kPackEw>(x[n1], *reinterpret_cast<const f32x2 *>(s_window + 2 * (16 * n1 + hl)));
      if (kind == SNB_FEAT_ENERGY) {
        double e = 0.0;
#pragma unroll
        for (int n1 = 0; n1 < kN1; ++n1) {
          float r, i;
          upk(x[n1], r, i);
          e += static_cast<double>(r) * r + static_cast<double>(i) * i;
        }
        e = 4.0 * group_sum_f64<16>(e);
        if (valid && hl == 0)
          reinterpret_cast<double *>(a.out)[(row0 + fl) * a.ld_out] =
              compress_energy(e, xo.energy_compression);
        break;
      }
      if (p.need_post_energy) {
        f32x2 e2 = pk(0.0f, 0.0f);
#pragma unroll
        for (int n1 = 0; n1 < kN1; ++n1) e2 = fma2t<kPackEw>(x[n1], x[n1], e2);
        float e0, e1;
        upk(e2, e0, e1);
        log_energy = logf(fmaxf(4.0f * group_sum<16>(e0 + e1), p.eps_energy));
      }

      // ---- 256-point complex FFT = 16 x 16 ----
      // (a 2-trip loop that is NOT unrolled: one copy of the radix-16 code)
      f32x2 *s_x = reinterpret_cast<f32x2 *>(s_grp);
#pragma unroll 1
      for (int fft_pass = 0; fft_pass < 2; ++fft_pass) {
        fft16p<kPackFft>(x);                  // pass 0: over n1 (lane = n2); pass 1: over n2 -> Z[hl + 16 k2]
        if (fft_pass == 0) {
#pragma unroll
          for (int k1 = 1; k1 < 16; ++k1) {
            const float2 w = s_tw1[k1 * 16 + hl];
            float r, i;
            upk(x[k1], r, i);
            x[k1] = pk(r * w.x - i * w.y, r * w.y + i * w.x);
          }
          __syncwarp();                 // every lane is done reading its noise
#pragma unroll
          for (int k1 = 0; k1 < 16; ++k1) s_x[k1 * kXStride + hl] = x[k1];
          __syncwarp();
#pragma unroll
          for (int n2 = 0; n2 < 16; ++n2) x[n2] = s_x[hl * kXStride + n2];
        }
      }
      __syncwarp();                                   // transpose buffer free -> reuse as P
      float *P = s_grp;
      // ---- real-FFT unpack: pairs (k, 256-k), k = hl + 16 k2, k2 < 8 ----
      // With Z = FFT of the half-scaled frame: X[k] = E + W^k O, X[256-k] =
      // conj(E - W^k O), E = Z[k] + conj Z[256-k], O = -i (Z[k] - conj Z[256-k]).
      // P in the padded layout of pidx<16>: k -> 24 k2 + pk, 256-k -> 24 (15-k2) + pq
      const int partner = gbase | ((16 - hl) & 15);
      const int pko = kPStride * (hl >> 3) + (hl & 7);
      const int pqo = (hl == 0) ? 2 * kPStride : kPStride * ((16 - hl) >> 3) + ((16 - hl) & 7);
      const f32x2 pm = pk(1.0f, -1.0f), mp = pk(-1.0f, 1.0f);
#pragma unroll
      for (int k2 = 0; k2 < 8; ++k2) {
        float sr, si, qr, qi;
        upk(x[15 - k2], sr, si);
        upk(x[(16 - k2) & 15], qr, qi);
        const float cr = __shfl_sync(SNB_FULL_MASK, (hl == 0) ? qr : sr, partner);   // Z[256-k]
        const float ci = __shfl_sync(SNB_FULL_MASK, (hl == 0) ? qi : si, partner);
        const f32x2 c = pk(cr, ci), z = x[k2];
        const int k = hl + 16 * k2;
        if (k == 0) {
          float ar, ai;
          upk(z, ar, ai);
          const float s0 = ar + ai, d0 = ar - ai;
          P[0] = (s0 + s0) * (s0 + s0);
          P[pidx<16>(256)] = (d0 + d0) * (d0 + d0);
        } else {
          const float2 w = s_tw2[k2 * 16 + hl];
          const f32x2 e = fma2t<kPackUn>(c, pm, z);    // (ar + cr, ai - ci)
          float u, v;
          upk(fma2t<kPackUn>(z, mp, c), v, u);         // (cr - ar, ci + ai)
          const f32x2 o = pk(w.x * u - w.y * v, w.x * v + w.y * u);   // W O
          float x1r, x1i, x2r, x2i;
          upk(add2t<kPackUn>(e, o), x1r, x1i);
          upk(sub2t<kPackUn>(e, o), x2r, x2i);
          P[2 * kPStride * k2 + pko] = fmaf(x1r, x1r, x1i * x1i);
          P[2 * kPStride * (15 - k2) + pqo] = fmaf(x2r, x2r, x2i * x2i);
        }
      }
      if (hl == 0) {
        float r8, i8;
        upk(x[8], r8, i8);
        P[pidx<16>(128)] = 4.0f * (r8 * r8 + i8 * i8);
      }
      __syncwarp();

      float *out_row = reinterpret_cast<float *>(a.out) + (row0 + fl) * a.ld_out;
      feature_tail<16>(p, tt, P, s_grp + kPFloats, log_energy, out_row, valid, hl);
    } while (false);
    if (tid == 0) cp_async_wait_all();   // descriptor of tile i+2 landed (visible after the barrier)
    __syncthreads();                     // this tile's PCM buffer, s_mel and ring slot are free
    slot = slot1;
    buf ^= 1;
  }
}

// ---------------------------------------------------------------------------
// generic path: any FFT size; one warp per frame
// ---------------------------------------------------------------------------
constexpr int kGenWarps = 4;

struct GenArgs {
  FeatParams p;
  int log2n;                 // -1 when N is not a power of two
  int warp_floats;           // floats of private smem per warp
  int tables_floats;         // floats of shared tables (dct|lifter|idft)
  const int64_t *sample_begin;
  const int64_t *sample_len;
  const int64_t *frame_offsets;
  const int32_t *utt_mel_idx;
  const int32_t *mel_blobs;
  const int16_t *pcm;
  const float *pcm_f32;      // float input (energy kind) when non-null
  int64_t nutts, total_frames;
  void *out;
  int64_t ld_out;
  uint64_t seed;
};

// (total = offsets[nutts]: the equal-length guess is tried first -- two loads
// instead of a chain of ~log2(nutts) dependent ones per frame)
__device__ __forceinline__ int64_t find_utt(const int64_t *offsets, int64_t nutts, int64_t row, int64_t total) {
#ifndef SNB_UTT_GUESS
#define SNB_UTT_GUESS 1
#endif
  if (SNB_UTT_GUESS && total > 0) {
    int64_t g = static_cast<int64_t>(static_cast<double>(row) * static_cast<double>(nutts) /
                                     static_cast<double>(total));
    g = min(max(g, static_cast<int64_t>(0)), nutts - 1);
    if (offsets[g] <= row && row < offsets[g + 1]) return g;
  }
  int64_t lo = 0, hi = nutts;      // offsets[lo] <= row < offsets[hi]
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (offsets[mid] <= row) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kGenWarps * 32)
generic_features_kernel(const GenArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const FeatParams &p = a.p;
  float *s_tables = reinterpret_cast<float *>(smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const snb_feat_opts &xo = p.xo;
  const int W = p.W, N = p.N, B = p.B, half = N / 2;
  float *s_dct = s_tables;
  float *s_lifter = s_dct + ((xo.kind == SNB_FEAT_MFCC) ? xo.num_ceps * B : 0);
  float *s_idft = s_lifter + xo.num_ceps;
  if (xo.kind == SNB_FEAT_MFCC)
    for (int i = tid; i < xo.num_ceps * B; i += blockDim.x) s_dct[i] = p.t.dct[i];
  if (xo.kind == SNB_FEAT_MFCC || xo.kind == SNB_FEAT_PLP)
    for (int i = tid; i < xo.num_ceps; i += blockDim.x) s_lifter[i] = p.t.lifter[i];
  if (xo.kind == SNB_FEAT_PLP)
    for (int i = tid; i < (xo.lpc_order + 1) * (B + 2); i += blockDim.x) s_idft[i] = p.t.idft[i];
  __syncthreads();
  float *buf_a = s_tables + a.tables_floats + static_cast<int64_t>(warp) * a.warp_floats;  // [N]
  float *buf_b = buf_a + N;                                                                // [N]
  float *scratch = buf_b + N;

  TailTables tt;
  tt.dct = s_dct; tt.lifter = s_lifter; tt.idft = s_idft;
  tt.dct_stride = B;

  for (int64_t row = static_cast<int64_t>(blockIdx.x) * kGenWarps + warp; row < a.total_frames;
       row += static_cast<int64_t>(gridDim.x) * kGenWarps) {
    const int64_t utt = find_utt(a.frame_offsets, a.nutts, row, a.total_frames);
    const int64_t f = row - a.frame_offsets[utt];
    const int64_t utt_off = a.sample_begin[utt];
    const int64_t n = a.sample_len[utt];
    const int64_t start = first_sample_of_frame_dev(f, p);
    if (B > 0)
      tt.mel = mel_view(a.mel_blobs + static_cast<int64_t>(a.utt_mel_idx[utt]) * p.mel_blob_stride, p);
    __syncwarp();
    // load (reflect at edges), dither
    float lsum = 0.0f;
    // dither: the generator of the fused path (one MUFU-only Box-Muller per
    // sample pair, keyed by the frame)
    const uint32_t nkey = frame_noise_key(a.seed, static_cast<uint64_t>(row));
    const float nc = -1.3862943611198906f * p.fo.dither * p.fo.dither;     // -2 ln 2 dither^2
    for (int i0 = 2 * lane; i0 < W; i0 += 64) {          // a lane takes sample PAIRS
      float g[2] = {0.0f, 0.0f};
      if (p.fo.dither != 0.0f) {
        const float2 nz = dither_pair(nkey, static_cast<uint32_t>(i0 >> 1), nc);
        g[0] = nz.x; g[1] = nz.y;
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = i0 + e;
        if (i >= W) break;
        int64_t k = start + i;
        while (k < 0 || k >= n) k = (k < 0) ? (-k - 1) : (2 * n - 1 - k);
        float v = a.pcm_f32 ? a.pcm_f32[utt_off + k] : static_cast<float>(a.pcm[utt_off + k]);
        v += g[e];
        buf_a[i] = v;
        lsum += v;
      }
    }
    float mean = 0.0f;
    if (p.fo.remove_dc_offset) mean = __fdiv_rn(group_sum<32>(lsum), static_cast<float>(W));
    __syncwarp();
    float log_energy = 0.0f;
    {
      float e = 0.0f;
      for (int i = lane; i < W; i += 32) {
        const float v = buf_a[i] - mean;
        buf_a[i] = v;
        e = fmaf(v, v, e);
      }
      if (p.need_raw_energy) log_energy = logf(fmaxf(group_sum<32>(e), p.eps_energy));
    }
    __syncwarp();
    // pre-emphasis + window -> buf_b (bit-reversed when a pow2 FFT follows)
    const float c = p.fo.preemph_coeff;
    float e_post = 0.0f;
    double e64 = 0.0;
    for (int i = lane; i < N; i += 32) {
      float v = 0.0f;
      if (i < W) {
        const float prev = buf_a[i > 0 ? i - 1 : 0];
        v = buf_a[i];
        if (c != 0.0f) v = fmaf(-c, prev, v);
        v *= p.t.window[i];
      }
      e_post = fmaf(v, v, e_post);
      e64 += static_cast<double>(v) * v;
      // (bit-reversed for the small radix-2 transform; natural order for the
      //  Stockham transform and the direct DFT)
      const int dst = (a.log2n >= 0 && a.log2n < 4)
                          ? static_cast<int>(__brev(static_cast<unsigned>(i)) >> (32 - a.log2n)) : i;
      buf_b[dst] = v;
    }
    if (xo.kind == SNB_FEAT_ENERGY) {
      e64 = group_sum_f64<32>(e64);
      if (lane == 0)
        reinterpret_cast<double *>(a.out)[row * a.ld_out] = compress_energy(e64, xo.energy_compression);
      continue;
    }
    if (p.need_post_energy) log_energy = logf(fmaxf(group_sum<32>(e_post), p.eps_energy));
    __syncwarp();
    float *P;
    if (a.log2n >= 4) {
      // Real FFT of N points as ONE complex FFT of M = N/2 points on
      // z[n] = x[2n] + i x[2n+1] (the float2 view of buf_b), Stockham autosort
      // (no bit reversal) with radix-4 stages -- one radix-2 stage first when
      // log2 M is odd -- ping-ponging between the two buffers, then the
      // split X[k] = E[k] + W_N^k O[k].  N = 256: 3 radix-4 stages + 1 radix-2
      // of 32 / 64 butterflies, one per lane (the radix-2 transform of the full
      // complex N-point problem this replaces ran 8 passes of 128).
      const int M = half, log2m = a.log2n - 1;
      float2 *zin = reinterpret_cast<float2 *>(buf_b), *zout = reinterpret_cast<float2 *>(buf_a);
      const float2 *tw = p.t.tw_dft;                       // exp(-2 pi i k / N), k < N
      int Ns = 1;
      if (log2m & 1) {
        for (int j = lane; j < M / 2; j += 32) {
          const float2 u = zin[j], v = zin[j + M / 2];
          zout[2 * j] = make_float2(u.x + v.x, u.y + v.y);
          zout[2 * j + 1] = make_float2(u.x - v.x, u.y - v.y);
        }
        __syncwarp();
        float2 *t = zin; zin = zout; zout = t;
        Ns = 2;
      }
      for (; Ns < M; Ns <<= 2) {
        const int tstep = 2 * (M / (Ns * 4));              // W_M^(k M / 4 Ns) = tw[2 k M / 4 Ns]
        const int q = M / 4;
        for (int j = lane; j < q; j += 32) {
          const int k = j & (Ns - 1);
          float2 v0 = zin[j], v1 = zin[j + q], v2 = zin[j + 2 * q], v3 = zin[j + 3 * q];
          if (Ns > 1) {
            const float2 w1 = __ldg(tw + k * tstep), w2 = __ldg(tw + 2 * k * tstep),
                         w3 = __ldg(tw + 3 * k * tstep);
            v1 = make_float2(v1.x * w1.x - v1.y * w1.y, v1.x * w1.y + v1.y * w1.x);
            v2 = make_float2(v2.x * w2.x - v2.y * w2.y, v2.x * w2.y + v2.y * w2.x);
            v3 = make_float2(v3.x * w3.x - v3.y * w3.y, v3.x * w3.y + v3.y * w3.x);
          }
          const float2 a0 = make_float2(v0.x + v2.x, v0.y + v2.y), a1 = make_float2(v0.x - v2.x, v0.y - v2.y);
          const float2 a2 = make_float2(v1.x + v3.x, v1.y + v3.y);
          const float2 a3 = make_float2(v1.y - v3.y, v3.x - v1.x);          // -i (v1 - v3)
          const int j0 = ((j - k) << 2) + k;
          zout[j0] = make_float2(a0.x + a2.x, a0.y + a2.y);
          zout[j0 + Ns] = make_float2(a1.x + a3.x, a1.y + a3.y);
          zout[j0 + 2 * Ns] = make_float2(a0.x - a2.x, a0.y - a2.y);
          zout[j0 + 3 * Ns] = make_float2(a1.x - a3.x, a1.y - a3.y);
        }
        __syncwarp();
        float2 *t = zin; zin = zout; zout = t;
      }
      // split + power spectrum into the other buffer
      float *pw = reinterpret_cast<float *>(zout);
      for (int k = lane; k <= M; k += 32) {
        const float2 zk = zin[k & (M - 1)], zm = zin[(M - k) & (M - 1)];
        const float er = 0.5f * (zk.x + zm.x), ei = 0.5f * (zk.y - zm.y);     // E = (Z[k] + conj Z[M-k]) / 2
        const float dr = zk.x - zm.x, di = zk.y + zm.y;                       // D = Z[k] - conj Z[M-k]
        const float orr = 0.5f * di, oi = -0.5f * dr;                         // O = -i D / 2
        const float2 w = __ldg(tw + k);
        const float xr = er + (orr * w.x - oi * w.y), xi = ei + (orr * w.y + oi * w.x);
        pw[k] = (k == 0 || k == M) ? xr * xr : xr * xr + xi * xi;
      }
      P = pw;
    } else if (a.log2n >= 0) {
      // in-place radix-2 DIT: real part buf_b, imaginary part buf_a
      for (int i = lane; i < N; i += 32) buf_a[i] = 0.0f;
      __syncwarp();
      for (int h = 1; h < N; h <<= 1) {
        const int tstride = N / (2 * h);
        for (int j = lane; j < half; j += 32) {
          const int pos = j & (h - 1);
          const int i0 = ((j - pos) << 1) + pos, i1 = i0 + h;
          const float2 w = p.t.tw_dft[pos * tstride];
          const float ur = buf_b[i0], ui = buf_a[i0];
          const float xr = buf_b[i1], xi = buf_a[i1];
          const float vr = xr * w.x - xi * w.y, vi = xr * w.y + xi * w.x;
          buf_b[i0] = ur + vr; buf_a[i0] = ui + vi;
          buf_b[i1] = ur - vr; buf_a[i1] = ui - vi;
        }
        __syncwarp();
      }
      for (int k = lane; k <= half; k += 32) {
        const float r = buf_b[k], i = buf_a[k];
        buf_b[k] = (k == 0 || k == half) ? r * r : r * r + i * i;
      }
      P = buf_b;
    } else {
      // direct DFT (round_to_power_of_two=False, Kaldi's generic RealFft)
      for (int k = lane; k <= half; k += 32) {
        float sr = 0.0f, si = 0.0f;
        int idx = 0;
        for (int t = 0; t < N; ++t) {
          const float2 w = p.t.tw_dft[idx];
          const float v = buf_b[t];
          sr = fmaf(v, w.x, sr);
          si = fmaf(v, w.y, si);
          idx += k;
          if (idx >= N) idx -= N;
        }
        buf_a[k] = (k == 0 || 2 * k == N) ? sr * sr : sr * sr + si * si;
      }
      P = buf_a;
    }
    __syncwarp();
    float *out_row = reinterpret_cast<float *>(a.out) + row * a.ld_out;
    feature_tail<32>(p, tt, P, scratch, log_energy, out_row, true, lane);
  }
}

// ---------------------------------------------------------------------------
// RASTA-PLP (plp.py:64-146): per (utterance, mel bin), frames in order.
// x = log(mel + FLT_EPSILON) in float32; frames 0..3 output exp(0) and prime
// the FIR state (lfilter(num, 1, first4, zi = lfilter_zi(num, 1) * x[0]));
// from frame 4 on y = sum b_k x_{t-k} + 0.94 y_{t-1} in float64 (direct form
// II transposed, like scipy.signal.lfilter), output exp((float)y).  In place.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) rasta_kernel(float *mel, int64_t ld, int col0, int B,
                                                    const int64_t *frame_offsets, int64_t nutts) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= nutts * B) return;
  const int64_t u = idx / B;
  const int b = static_cast<int>(idx - u * B);
  const int64_t first = frame_offsets[u], F = frame_offsets[u + 1] - first;
  const double b0 = 0.2, b1 = 0.1, b2 = -0.0, b3 = -0.1, b4 = -0.2;
  float *col = mel + first * ld + col0 + b;
  double z0 = 0.0, z1 = 0.0, z2 = 0.0, z3 = 0.0;
  for (int64_t t = 0; t < F; ++t) {
    const float xf = logf(col[t * ld] + FLT_EPSILON);
    const double x = static_cast<double>(xf);
    if (t < 4) {
      if (t == 0) {
        // lfilter_zi(num, 1) = [b1+b2+b3+b4, b2+b3+b4, b3+b4, b4], scaled by x[0]
        z3 = b4 * x; z2 = (b3 + b4) * x; z1 = (b2 + (b3 + b4)) * x; z0 = (b1 + (b2 + (b3 + b4))) * x;
      }
      z0 = b1 * x + z1; z1 = b2 * x + z2; z2 = b3 * x + z3; z3 = b4 * x;
      col[t * ld] = expf(0.0f);
    } else {
      const double y = b0 * x + z0;
      z0 = b1 * x + z1 + 0.94 * y; z1 = b2 * x + z2; z2 = b3 * x + z3; z3 = b4 * x;
      col[t * ld] = expf(static_cast<float>(y));
    }
  }
}

// PLP tail from precomputed (filtered) mel energies: one warp per frame
struct PlpMelArgs {
  FeatParams p;
  const float *mel;          // [F, ld_mel]: column 0 log-energy, columns 1..B mel
  int64_t ld_mel;
  const int64_t *frame_offsets;
  const int32_t *utt_mel_idx;   // may be NULL (blob 0)
  const int32_t *mel_blobs;
  int64_t nutts, total_frames;
  float *out;
  int64_t ld_out;
  int tables_floats, warp_floats;
};

__global__ void __launch_bounds__(kGenWarps * 32) plp_from_mel_kernel(const PlpMelArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const FeatParams &p = a.p;
  const snb_feat_opts &xo = p.xo;
  float *s_tables = reinterpret_cast<float *>(smem);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int B = p.B;
  float *s_lifter = s_tables;
  float *s_idft = s_lifter + xo.num_ceps;
  for (int i = tid; i < xo.num_ceps; i += blockDim.x) s_lifter[i] = p.t.lifter[i];
  for (int i = tid; i < (xo.lpc_order + 1) * (B + 2); i += blockDim.x) s_idft[i] = p.t.idft[i];
  __syncthreads();
  float *scratch = s_tables + a.tables_floats + static_cast<int64_t>(warp) * a.warp_floats;
  TailTables tt;
  tt.dct = nullptr; tt.lifter = s_lifter; tt.idft = s_idft;
  tt.dct_stride = 0;
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * kGenWarps + warp; row < a.total_frames;
       row += static_cast<int64_t>(gridDim.x) * kGenWarps) {
    const int64_t utt = find_utt(a.frame_offsets, a.nutts, row, a.total_frames);
    // all the utterances of a RASTA batch share the blob unless VTLN warps differ
    int mel_idx = 0;
    if (a.utt_mel_idx) mel_idx = a.utt_mel_idx[utt];
    tt.mel = mel_view(a.mel_blobs + static_cast<int64_t>(mel_idx) * p.mel_blob_stride, p);
    const float *mrow = a.mel + row * a.ld_mel;
    float log_energy = mrow[0];
    if (xo.energy_floor > 0.0f && log_energy < p.log_energy_floor) log_energy = p.log_energy_floor;
    __syncwarp();
    for (int b = lane; b < B; b += 32)
      scratch[1 + b] = powf(mrow[1 + b] * tt.mel.loudness[b], xo.compress_factor);
    __syncwarp();
    plp_finish<32>(p, tt, scratch, log_energy, a.out + row * a.ld_out, true, lane);
  }
}

// ---------------------------------------------------------------------------
// tile table of a fused-path batch, expanded on the device from O(utterances)
// host data: utterance u owns tiles [tile_first[u], tile_first[u+1]) of T frames
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(64) expand_tiles_kernel(
    TileDesc *out, const int64_t *tile_first, const int64_t *sample_begin, const int64_t *sample_len,
    const int64_t *frame_offsets, const int32_t *utt_mel, int64_t nutts, int T, int S, int W, int snip_edges) {
  for (int64_t u = blockIdx.x; u < nutts; u += gridDim.x) {
    const int64_t t0 = tile_first[u], nt = tile_first[u + 1] - t0;
    const int64_t row0 = frame_offsets[u], nf_utt = frame_offsets[u + 1] - row0;
    const int64_t begin = sample_begin[u], len = sample_len[u];
    const int32_t mel_idx = utt_mel ? utt_mel[u] : 0;
    for (int64_t t = threadIdx.x; t < nt; t += blockDim.x) {
      TileDesc td;
      td.utt = static_cast<int32_t>(u);
      td.f0 = static_cast<int32_t>(t * T);
      td.nf = static_cast<int32_t>(nf_utt - t * T < T ? nf_utt - t * T : T);
      td.mel_idx = mel_idx;
      td.row0 = row0 + td.f0;
      const int64_t a0 = snip_edges ? static_cast<int64_t>(td.f0) * S
                                    : static_cast<int64_t>(td.f0) * S + S / 2 - W / 2;
      const int64_t span = static_cast<int64_t>(td.nf - 1) * S + W;
      td.g0 = (a0 >= 0 && a0 + span <= len) ? begin + a0 : -1;
      out[t0 + t] = td;
    }
  }
}

// ---------------------------------------------------------------------------
// host: plan
// ---------------------------------------------------------------------------
static int align_up(int v, int a) { return (v + a - 1) / a * a; }

static void fast_layout(const snb_plan *plan, FastSmemLayout *sm) {
  const FeatParams &p = plan->params;
  const snb_feat_opts &xo = p.xo;
  int off = 0;
  sm->window = off; off += 512 * 4;
  sm->tw1 = off; off += 256 * 8;
  sm->tw2 = off; off += 128 * 8;
  // DCT rows padded to a stride = 4 (mod 8) floats: a quarter warp's 128-bit
  // row loads then fall in distinct banks
  sm->dct_stride = align_up(p.B, 4);
  if (sm->dct_stride % 8 == 0) sm->dct_stride += 4;
  sm->dct = off; off += align_up((xo.kind == SNB_FEAT_MFCC ? xo.num_ceps * sm->dct_stride : 0) * 4, 16);
  sm->lifter = off; off += align_up(xo.num_ceps * 4, 16);
  sm->idft = off; off += align_up((xo.kind == SNB_FEAT_PLP ? (xo.lpc_order + 1) * (p.B + 2) : 0) * 4, 16);
  sm->mel = off; off += align_up(p.mel_fast_words * 4, 16);
  sm->span_cap = align_up((plan->tile_frames - 1) * p.S + p.W + 16, 8);
  sm->pcm = off; off += 2 * sm->span_cap * 2;                   // double buffered
  // per-group buffers 16 banks apart (size = 16 mod 32 floats): the two
  // groups of a warp use the same offsets inside their buffers, and with a
  // multiple of 32 every 32-bit access of the warp was a 2-way bank conflict.
  // Contents: the transpose buffer (16 x 17 float2), later the power spectrum
  // [kPFloats] followed by the tail scratch (mel[B+4] | autocorrelation)
  sm->grp_floats = align_up(std::max(16 * kXStride * 2, kPFloats + (p.B + 4) + (xo.lpc_order + 2)), 32) + 16;
  sm->grp = off; off += kFastGroups * sm->grp_floats * 4;
  // run partials of the mel stage: one float4 per run, per warp
  sm->part_floats = 4 * align_up(32 + p.B + 2, 2);
  sm->part = off; off += (kFastThreads / 32) * sm->part_floats * 4;
  sm->desc = off; off += 3 * static_cast<int>(sizeof(TileDesc));
  sm->bar = off; off += 16;
  sm->total = off;
}

int feature_plan_finalize(snb_plan *plan) {
  FeatParams &p = plan->params;
  const snb_feat_opts &xo = p.xo;
  plan->fast_path = false;
  const int group_need = (xo.kind == SNB_FEAT_PLP) ? xo.lpc_order + 1 : 1;
  if (p.N == 512 && p.W > 256 && p.B <= 200 && group_need <= 16) {
    // one frame per lane group and tile; fewer when a huge frame shift would
    // make the double-buffered span larger than 2 x 24 KB
    static const int forced_occ = getenv("SNB_FUSED_OCC") ? atoi(getenv("SNB_FUSED_OCC")) : 0;
    int t = kFastGroups;
    while (t > 1 && ((t - 1) * p.S + p.W + 16) * 2 > 24 * 1024) t /= 2;
    plan->tile_frames = t;
    FastSmemLayout sm;
    fast_layout(plan, &sm);
    const int budget4 = (227 * 1024) / 4 - 2048;     // per CTA, minus static + reserved
    plan->fused_occ = (sm.total <= budget4 && forced_occ != 3) ? 4 : 3;
    if (sm.total <= 200 * 1024) {
      plan->fast_path = true;
      plan->smem_bytes = sm.total;
      return SNB_OK;
    }
  }
  if (group_need > 32)
    return set_error(SNB_ERR_UNSUPPORTED, "lpc_order > 31 is not supported");
  if (p.N > 8192)
    return set_error(SNB_ERR_UNSUPPORTED, "frame length above 8192 samples is not supported");
  plan->tile_frames = 1;
  const int tables = (xo.kind == SNB_FEAT_MFCC ? xo.num_ceps * p.B : 0) + xo.num_ceps +
                     (xo.kind == SNB_FEAT_PLP ? (xo.lpc_order + 1) * (p.B + 2) : 0);
  const int warp_floats = 2 * p.N + 3 * (p.B + 2) + xo.lpc_order + 2 + 8;
  plan->smem_bytes = static_cast<size_t>(align_up(tables, 4) + kGenWarps * align_up(warp_floats, 4)) * 4;
  if (plan->smem_bytes > 220 * 1024)
    return set_error(SNB_ERR_UNSUPPORTED, "options need too much shared memory");
  return SNB_OK;
}

// builds the int32 mel blob for one VTLN warp (cached in the plan)
static int get_mel_blob(const snb_plan *plan, float warp, const std::vector<int32_t> **out) {
  uint32_t key;
  std::memcpy(&key, &warp, 4);
  std::lock_guard<std::mutex> lock(plan->mu);
  auto it = plan->mel_blobs.find(key);
  if (it == plan->mel_blobs.end()) {
    MelBanksHost mb;
    int rc = build_mel_banks(plan->fo, plan->mo, warp, &mb);
    if (rc != SNB_OK) return rc;
    const FeatParams &p = plan->params;
    if (static_cast<int>(mb.weights.size()) > p.mel_wcap)
      return set_error(SNB_ERR_UNSUPPORTED, "mel weights overflow (%zu > %d)", mb.weights.size(), p.mel_wcap);
    std::vector<int32_t> blob(p.mel_blob_stride, 0);
    const int B = p.B;
    std::vector<float> loud;
    build_equal_loudness(mb.center_freqs, &loud);
    int32_t *gen = blob.data() + p.mel_gen_off;
    for (int b = 0; b < B; ++b) {
      std::memcpy(&blob[b], &loud[b], 4);
      gen[b] = mb.first[b];
      gen[B + b] = mb.size[b];
      gen[2 * B + b] = mb.offset[b];
    }
    std::memcpy(gen + 3 * B, mb.weights.data(), mb.weights.size() * 4);
    if (mb.num_fft_bins == 256) {
      // fused path: bins dealt to the 32 lanes in chunks of 8; a run is a
      // maximal stretch of bins of one segment inside one chunk.  Bins below
      // the first / above the last triangle get the pseudo segments -1 / B+1
      // (zero weights, their runs are never read back).
      std::vector<int32_t> sg(256);
      bool seen = false;
      for (int k = 0; k < 256; ++k) {
        sg[k] = mb.seg_of[k];
        if (sg[k] >= 0) seen = true;
        else if (seen) sg[k] = B + 1;
      }
      float *cw = reinterpret_cast<float *>(blob.data() + p.mel_chunk_off);
      int32_t *meta = blob.data() + p.mel_meta_off;
      int32_t *run_first = blob.data() + p.mel_run_off;
      std::vector<int32_t> run_seg;
      for (int l = 0; l < 32; ++l) {
        uint32_t flags = 0;
        const uint32_t first_run = static_cast<uint32_t>(run_seg.size());
        for (int i = 0; i < 8; ++i) {
          const int k = 8 * l + i;
          cw[20 * l + 2 * i] = mb.up[k];
          cw[20 * l + 2 * i + 1] = mb.down[k];
          if (i == 7 || sg[k + 1] != sg[k]) {
            flags |= 1u << i;
            run_seg.push_back(sg[k]);
          }
        }
        meta[l] = static_cast<int32_t>(flags | (first_run << 8));
      }
      for (size_t j = 1; j < run_seg.size(); ++j)
        if (run_seg[j] < run_seg[j - 1])
          return set_error(SNB_ERR_UNSUPPORTED, "non monotonic mel segments");
      if (static_cast<int>(run_seg.size()) > 32 + B + 2)
        return set_error(SNB_ERR_UNSUPPORTED, "too many mel runs (%zu)", run_seg.size());
      for (int sgm = 0; sgm <= B + 1; ++sgm)
        run_first[sgm] = static_cast<int32_t>(
            std::lower_bound(run_seg.begin(), run_seg.end(), sgm) - run_seg.begin());
      // longest run list of a segment, kept in the spare word after run_first
      int32_t longest = 1;
      for (int sgm = 0; sgm <= B; ++sgm) longest = std::max(longest, run_first[sgm + 1] - run_first[sgm]);
      run_first[B + 2] = longest;
    }
    it = plan->mel_blobs.emplace(key, std::move(blob)).first;
  }
  *out = &it->second;
  return SNB_OK;
}

}  // namespace snb

using namespace snb;

extern "C" int snb_feature_plan_create(const snb_frame_opts *fo, const snb_mel_opts *mo,
                                       const snb_feat_opts *xo, snb_plan **out) {
  if (!fo || !xo || !out) return set_error(SNB_ERR_VALUE, "null argument");
  *out = nullptr;
  const int kind = xo->kind;
  if (kind < SNB_FEAT_SPECTROGRAM || kind > SNB_FEAT_ENERGY)
    return set_error(SNB_ERR_VALUE, "unknown feature kind %d", kind);
  const bool needs_mel = kind == SNB_FEAT_FBANK || kind == SNB_FEAT_MFCC || kind == SNB_FEAT_PLP;
  if (needs_mel && !mo) return set_error(SNB_ERR_VALUE, "mel options required");
  snb_plan *plan = new snb_plan();
  plan->kind = 0;
  plan->fo = *fo;
  plan->xo = *xo;
  if (needs_mel) plan->mo = *mo;
  plan->has_mel = needs_mel;
  if (kind == SNB_FEAT_ENERGY && xo->raw_energy) {
    // energy.py:148-151: raw energy = no pre-emphasis, rectangular window
    plan->fo.preemph_coeff = 0.0f;
    plan->fo.window_type = SNB_WIN_RECTANGULAR;
  }
  cudaGetDevice(&plan->device);
  cudaGetLastError();
  FeatParams &p = plan->params;
  std::memset(&p, 0, sizeof(p));
  p.fo = plan->fo;
  p.xo = *xo;
  p.W = window_size(plan->fo);
  p.S = window_shift(plan->fo);
  p.N = padded_window_size(plan->fo);
  int rc = SNB_OK;
  auto fail = [&](int code) { delete plan; return code; };
  if (p.W <= 0 || p.S <= 0)
    return fail(set_error(SNB_ERR_OPTION, "frame length/shift too small for the sample rate"));
  if (p.W < 2) return fail(set_error(SNB_ERR_OPTION, "window of less than 2 samples"));
  p.B = needs_mel ? mo->num_bins : 0;
  switch (kind) {
    case SNB_FEAT_SPECTROGRAM:
      p.dim = p.N / 2 + 1;
      p.need_raw_energy = xo->raw_energy != 0;
      p.need_post_energy = !xo->raw_energy;
      break;
    case SNB_FEAT_FBANK:
      p.dim = p.B + (xo->use_energy ? 1 : 0);
      break;
    case SNB_FEAT_MFCC:
      if (xo->num_ceps <= 0 || xo->num_ceps > p.B)
        return fail(set_error(SNB_ERR_OPTION, "num-ceps cannot be larger than num-mel-bins: %d vs %d",
                              xo->num_ceps, p.B));
      p.dim = xo->num_ceps;
      break;
    case SNB_FEAT_PLP:
      if (xo->num_ceps <= 0 || xo->num_ceps > xo->lpc_order + 1)
        return fail(set_error(SNB_ERR_OPTION, "num_ceps must be in [1, lpc_order+1]"));
      p.dim = xo->num_ceps;
      break;
    default:
      p.dim = 1;
  }
  if (kind != SNB_FEAT_SPECTROGRAM && kind != SNB_FEAT_ENERGY) {
    p.need_raw_energy = xo->use_energy && xo->raw_energy;
    p.need_post_energy = xo->use_energy && !xo->raw_energy;
  }
  p.log_energy_floor = xo->energy_floor > 0.0f ? logf(xo->energy_floor) : 0.0f;
  p.eps_energy = (kind == SNB_FEAT_PLP) ? 2.220446049250313e-16f : FLT_EPSILON;
  if (p.N % 2 != 0 && kind != SNB_FEAT_ENERGY)
    return fail(set_error(SNB_ERR_OPTION, "padded window size must be even, it is %d", p.N));
  // validate the mel options now (KALDI_ERR at construction of the computer)
  if (needs_mel) {
    p.mel_wcap = 2 * (p.N / 2) + 2 * p.B + 8;
    p.mel_chunk_off = align_up(p.B, 4);
    p.mel_meta_off = p.mel_chunk_off + 32 * 20;
    p.mel_run_off = p.mel_meta_off + 32;
    p.mel_fast_words = align_up(p.mel_run_off + p.B + 3, 4);     // run_first[B+2] + longest run list
    p.mel_gen_off = p.mel_fast_words;
    p.mel_blob_stride = align_up(p.mel_gen_off + 3 * p.B + p.mel_wcap, 4);
    const std::vector<int32_t> *blob;
    rc = get_mel_blob(plan, 1.0f, &blob);
    if (rc != SNB_OK) return fail(rc);
  }
  // ---- fixed tables: window | tw_half | tw_full | dct | lifter | idft | tw_dft
  std::vector<float> win;
  window_function(plan->fo, &win);
  const int half = p.N / 2;
  std::vector<float> host;
  auto append = [&](const std::vector<float> &v) {
    size_t off = host.size();
    host.insert(host.end(), v.begin(), v.end());
    while (host.size() % 4) host.push_back(0.0f);
    return off;
  };
  const double two_pi = 6.283185307179586476925286766559005;
  std::vector<float> tw_half(2 * std::max(half, 1)), tw_full(2 * std::max(half, 1)), tw_dft(2 * p.N);
  for (int m = 0; m < half; ++m) {
    tw_half[2 * m] = static_cast<float>(std::cos(two_pi * m / half));
    tw_half[2 * m + 1] = static_cast<float>(-std::sin(two_pi * m / half));
    tw_full[2 * m] = static_cast<float>(std::cos(two_pi * m / p.N));
    tw_full[2 * m + 1] = static_cast<float>(-std::sin(two_pi * m / p.N));
  }
  for (int m = 0; m < p.N; ++m) {
    tw_dft[2 * m] = static_cast<float>(std::cos(two_pi * m / p.N));
    tw_dft[2 * m + 1] = static_cast<float>(-std::sin(two_pi * m / p.N));
  }
  std::vector<float> dct, lifter, idft;
  if (kind == SNB_FEAT_MFCC) build_dct(xo->num_ceps, p.B, &dct);
  if (kind == SNB_FEAT_MFCC || kind == SNB_FEAT_PLP) build_lifter(xo->num_ceps, xo->cepstral_lifter, &lifter);
  if (kind == SNB_FEAT_PLP) build_idft_bases(xo->lpc_order + 1, p.B + 2, &idft);
  const size_t o_win = append(win), o_th = append(tw_half), o_tf = append(tw_full),
               o_dct = append(dct), o_lift = append(lifter), o_idft = append(idft),
               o_dft = append(tw_dft);
  float *d = nullptr;
  cudaError_t e = cudaMalloc(&d, host.size() * sizeof(float));
  if (e == cudaSuccess)
    e = upload(d, host.data(), host.size() * sizeof(float));
  if (e != cudaSuccess) {
    // Without a usable GPU the product cannot run: fail loudly.
    cudaGetLastError();
    if (d) cudaFree(d);
    return fail(set_error(SNB_ERR_CUDA, "cannot upload plan tables: %s", cudaGetErrorString(e)));
  }
  plan->d_tables = d;
  p.t.window = d + o_win;
  p.t.tw_half = reinterpret_cast<const float2 *>(d + o_th);
  p.t.tw_full = reinterpret_cast<const float2 *>(d + o_tf);
  p.t.dct = d + o_dct;
  p.t.lifter = d + o_lift;
  p.t.idft = d + o_idft;
  p.t.tw_dft = reinterpret_cast<const float2 *>(d + o_dft);
  rc = feature_plan_finalize(plan);
  if (rc != SNB_OK) {
    cudaFree(d);
    plan->d_tables = nullptr;
    return fail(rc);
  }
  if (kind == SNB_FEAT_PLP && xo->rasta) {
    // the frame-recursive RASTA filter sits between the mel energies and the
    // PLP tail: an internal filterbank-kind plan emits [log-energy | linear mel]
    snb_feat_opts mx;
    std::memset(&mx, 0, sizeof(mx));
    mx.kind = SNB_FEAT_FBANK;
    mx.use_energy = 1;
    mx.raw_energy = xo->raw_energy;
    mx.use_log_fbank = 0;
    mx.use_power = 1;
    snb_plan *mel_plan = nullptr;
    rc = snb_feature_plan_create(fo, mo, &mx, &mel_plan);
    if (rc != SNB_OK) {
      snb_plan_destroy(plan);
      return rc;
    }
    mel_plan->params.eps_energy = p.eps_energy;   // plp.py floors with float64 eps
    if (!xo->use_energy) {
      mel_plan->params.need_raw_energy = 0;
      mel_plan->params.need_post_energy = 0;
    }
    plan->rasta_mel_plan = mel_plan;
  }
  *out = plan;
  return SNB_OK;
}

extern "C" void snb_plan_destroy(snb_plan *plan) {
  if (!plan) return;
  if (plan->d_tables) cudaFree(plan->d_tables);
  if (plan->rasta_mel_plan) snb_plan_destroy(plan->rasta_mel_plan);
  if (plan->kind == 1) pitch_plan_free(plan);
  delete plan;
}

extern "C" int32_t snb_plan_dim(const snb_plan *plan) {
  return plan->kind == 0 ? plan->params.dim : 2;
}
extern "C" int32_t snb_plan_uses_fast_path(const snb_plan *plan) { return plan->fast_path ? 1 : 0; }

// ---------------------------------------------------------------------------
// host: batch
// ---------------------------------------------------------------------------
static int batch_create_impl(const snb_plan *plan, const int64_t *sample_begin, const int64_t *sample_len,
                             int64_t nutts, const float *vtln_warps, bool on_stream, cudaStream_t stream,
                             snb_batch **out);

extern "C" int snb_batch_create(const snb_plan *plan, const int64_t *sample_begin,
                                const int64_t *sample_len, int64_t nutts, const float *vtln_warps,
                                snb_batch **out) {
  return batch_create_impl(plan, sample_begin, sample_len, nutts, vtln_warps, false, nullptr, out);
}

extern "C" int snb_batch_create_on_stream(const snb_plan *plan, const int64_t *sample_begin,
                                          const int64_t *sample_len, int64_t nutts, const float *vtln_warps,
                                          void *stream, snb_batch **out) {
  return batch_create_impl(plan, sample_begin, sample_len, nutts, vtln_warps, true,
                           static_cast<cudaStream_t>(stream), out);
}

static int batch_create_impl(const snb_plan *plan, const int64_t *sample_begin, const int64_t *sample_len,
                             int64_t nutts, const float *vtln_warps, bool on_stream, cudaStream_t stream,
                             snb_batch **out) {
  if (!plan || !out || nutts < 0 || (nutts > 0 && (!sample_begin || !sample_len)))
    return set_error(SNB_ERR_VALUE, "bad argument");
  *out = nullptr;
  snb_batch *b = new snb_batch();
  b->plan = plan;
  b->nutts = nutts;
  b->sample_begin.assign(sample_begin, sample_begin + nutts);
  b->sample_len.assign(sample_len, sample_len + nutts);
  b->total_samples = 0;
  b->frame_offsets.assign(nutts + 1, 0);
  auto fail = [&](int code) { snb_batch_destroy(b); return code; };
  for (int64_t u = 0; u < nutts; ++u) {
    const int64_t n = sample_len[u];
    if (n < 0 || sample_begin[u] < 0) return fail(set_error(SNB_ERR_VALUE, "negative sample begin/length"));
    b->total_samples = std::max(b->total_samples, sample_begin[u] + n);
    int64_t nf;
    if (plan->kind == 0) nf = num_frames(n, plan->fo);
    else nf = snb_pitch_num_frames(n, &plan->po);
    if (nf < 0) nf = 0;
    if (nf > 0x7fffffff) return fail(set_error(SNB_ERR_VALUE, "utterance too long"));
    b->frame_offsets[u + 1] = b->frame_offsets[u] + nf;
  }
  b->total_frames = b->frame_offsets[nutts];
  // every device-side table of the batch lives in ONE allocation filled by
  // ONE upload (16-byte aligned sections): batch creation costs a single
  // cudaMalloc + one host->device copy
  std::vector<unsigned char> stage;
  auto add_section = [&](const void *src, size_t bytes) {
    const size_t off = (stage.size() + 15) / 16 * 16;
    stage.resize(off + bytes);
    if (bytes) std::memcpy(stage.data() + off, src, bytes);
    return off;
  };
  const size_t o_begin = add_section(b->sample_begin.data(), nutts * sizeof(int64_t));
  const size_t o_len = add_section(b->sample_len.data(), nutts * sizeof(int64_t));
  const size_t o_foff = add_section(b->frame_offsets.data(), (nutts + 1) * sizeof(int64_t));
  size_t o_down = 0, o_mel = 0, o_uttmel = 0, o_tfirst = 0;
  bool has_down = false, has_mel = false, has_tiles = false, has_uttmel = false;
  if (plan->kind == 1) {
    int rc = pitch_batch_init(plan, b);
    if (rc != SNB_OK) return fail(rc);
    o_down = add_section(b->down_offsets.data(), b->down_offsets.size() * sizeof(int64_t));
    has_down = true;
  } else {
    // ---- mel blobs for the distinct VTLN warps of this batch ----
    std::vector<int32_t> utt_mel(nutts, 0);
    if (plan->has_mel) {
      std::vector<int32_t> blobs;
      std::map<uint32_t, int32_t> index;
      for (int64_t u = 0; u < nutts; ++u) {
        const float w = vtln_warps ? vtln_warps[u] : 1.0f;
        uint32_t key;
        std::memcpy(&key, &w, 4);
        auto it = index.find(key);
        if (it == index.end()) {
          const std::vector<int32_t> *blob;
          int rc = get_mel_blob(plan, w, &blob);
          if (rc != SNB_OK) return fail(rc);
          it = index.emplace(key, static_cast<int32_t>(index.size())).first;
          blobs.insert(blobs.end(), blob->begin(), blob->end());
          b->mel_max_runs = std::max(b->mel_max_runs, (*blob)[plan->params.mel_run_off + plan->params.B + 2]);
        }
        utt_mel[u] = it->second;
      }
      b->nblobs = static_cast<int32_t>(index.size());
      if (nutts == 0) {
        const std::vector<int32_t> *blob;
        int rc = get_mel_blob(plan, 1.0f, &blob);
        if (rc != SNB_OK) return fail(rc);
        blobs = *blob;
        b->nblobs = 1;
        b->mel_max_runs = (*blob)[plan->params.mel_run_off + plan->params.B + 2];
      }
      o_mel = add_section(blobs.data(), blobs.size() * sizeof(int32_t));
      has_mel = true;
    }
    // ---- per-utterance mel index; fused path: first tile of every utterance
    //      (the tile table itself is expanded on the device, see below) ----
    o_uttmel = add_section(utt_mel.data(), nutts * sizeof(int32_t));
    has_uttmel = true;
    if (plan->fast_path) {
      const int T = plan->tile_frames;
      std::vector<int64_t> tile_first(nutts + 1, 0);
      for (int64_t u = 0; u < nutts; ++u) {
        const int64_t nf = b->frame_offsets[u + 1] - b->frame_offsets[u];
        tile_first[u + 1] = tile_first[u] + (nf + T - 1) / T;
      }
      b->ntiles = tile_first[nutts];
      o_tfirst = add_section(tile_first.data(), tile_first.size() * sizeof(int64_t));
      has_tiles = true;
    }
  }
  // device-only tail of the blob: the expanded tile table
  const size_t o_tiles = (stage.size() + 15) / 16 * 16;
  const size_t blob_bytes = o_tiles + (has_tiles ? static_cast<size_t>(b->ntiles) * sizeof(TileDesc) : 0) + 16;
  unsigned char *d = nullptr;
  cudaError_t e;
  if (on_stream) {
    // stream-ordered: pooled device blob, pooled pinned staging, one async copy
    // queued on the caller's stream; nothing here waits for the device
    int dev = 0;
    cudaGetDevice(&dev);
    PoolBlock host;
    e = device_pool().acquire(blob_bytes, dev, &b->dev_block);
    if (e == cudaSuccess) e = pinned_pool().acquire(stage.size() + 16, dev, &host);
    if (e == cudaSuccess) {
      b->pooled = true;
      b->streams.push_back(stream);
      d = static_cast<unsigned char *>(b->dev_block.p);
      std::memcpy(host.p, stage.data(), stage.size());
      if (!stage.empty()) e = cudaMemcpyAsync(d, host.p, stage.size(), cudaMemcpyHostToDevice, stream);
      cudaEvent_t ev = nullptr;
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventRecord(ev, stream);
      if (e != cudaSuccess) cudaStreamSynchronize(stream);       // staging must not be reused early
      if (ev && e == cudaSuccess) host.pending.push_back(ev);
      else if (ev) cudaEventDestroy(ev);
      pinned_pool().release(host);
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(set_error(SNB_ERR_CUDA, "batch upload failed: %s", cudaGetErrorString(e)));
    }
  } else {
    e = cudaMalloc(&d, blob_bytes);
    if (e == cudaSuccess) e = upload(d, stage.data(), stage.size());
    if (e != cudaSuccess) {
      if (d) cudaFree(d);
      return fail(set_error(SNB_ERR_CUDA, "batch upload failed: %s", cudaGetErrorString(e)));
    }
  }
  b->d_blob = d;
  b->d_sample_begin = reinterpret_cast<int64_t *>(d + o_begin);
  b->d_sample_len = reinterpret_cast<int64_t *>(d + o_len);
  b->d_frame_offsets = reinterpret_cast<int64_t *>(d + o_foff);
  if (has_down) b->d_down_offsets = reinterpret_cast<int64_t *>(d + o_down);
  if (has_mel) b->d_mel_blobs = reinterpret_cast<int32_t *>(d + o_mel);
  if (has_uttmel) b->d_utt_mel = reinterpret_cast<int32_t *>(d + o_uttmel);
  if (has_tiles && b->ntiles > 0) {
    b->d_tile_first = reinterpret_cast<int64_t *>(d + o_tfirst);
    b->d_tiles = reinterpret_cast<TileDesc *>(d + o_tiles);
    // (expanded on the device by the first launch that needs it, on ITS stream:
    // a kernel queued here, on the upload stream of a chunked pipeline, would
    // stall the next chunk's PCM copy behind whatever the SMs are running)
  }
  *out = b;
  return SNB_OK;
}

// Expands the tile table once, on the stream of the first compute call; later
// calls on other streams wait for that expansion through an event.
static int ensure_tiles(const snb_plan *plan, const snb_batch *batch, cudaStream_t stream) {
  if (!batch->d_tiles || batch->ntiles == 0) return SNB_OK;
  std::lock_guard<std::mutex> lock(batch->stream_mu);
  if (!batch->tiles_ready) {
    const FeatParams &p = plan->params;
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>(batch->nutts, 8192));
    expand_tiles_kernel<<<grid, 64, 0, stream>>>(batch->d_tiles, batch->d_tile_first, batch->d_sample_begin,
                                                 batch->d_sample_len, batch->d_frame_offsets, batch->d_utt_mel,
                                                 batch->nutts, plan->tile_frames, p.S, p.W, p.fo.snip_edges);
    SNB_LAUNCH_CHECK();
    SNB_CUDA_CHECK(cudaEventCreateWithFlags(&batch->tiles_event, cudaEventDisableTiming));
    SNB_CUDA_CHECK(cudaEventRecord(batch->tiles_event, stream));
    batch->tiles_stream = stream;
    batch->tiles_ready = true;
  } else if (stream != batch->tiles_stream) {
    SNB_CUDA_CHECK(cudaStreamWaitEvent(stream, batch->tiles_event, 0));
  }
  return SNB_OK;
}

extern "C" void snb_batch_destroy(snb_batch *b) {
  if (!b) return;
  if (b->tiles_event) cudaEventDestroy(b->tiles_event);
  if (b->pooled) {
    // recycle the blob once everything queued so far on the streams that used
    // the batch has drained; an unrecordable stream falls back to cudaFree
    bool ok = true;
    for (cudaStream_t s : b->streams) {
      cudaEvent_t ev = nullptr;
      cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventRecord(ev, s);
      if (e != cudaSuccess) {
        if (ev) cudaEventDestroy(ev);
        cudaGetLastError();
        ok = false;
        break;
      }
      b->dev_block.pending.push_back(ev);
    }
    if (ok) {
      device_pool().release(b->dev_block);
    } else {
      for (cudaEvent_t ev : b->dev_block.pending) cudaEventDestroy(ev);
      if (b->dev_block.p) cudaFree(b->dev_block.p);
    }
  } else if (b->d_blob) {
    cudaFree(b->d_blob);
  }
  delete b;
}
extern "C" int64_t snb_batch_num_utts(const snb_batch *b) { return b->nutts; }
extern "C" int64_t snb_batch_total_frames(const snb_batch *b) { return b->total_frames; }
extern "C" const int64_t *snb_batch_frame_offsets(const snb_batch *b) { return b->frame_offsets.data(); }
extern "C" const int64_t *snb_batch_frame_offsets_device(const snb_batch *b) { return b->d_frame_offsets; }

// ---------------------------------------------------------------------------
// host: launch
// ---------------------------------------------------------------------------
// the max-dynamic-smem attribute is per function: only ever raise it
template <typename K>
static int ensure_smem(K kernel, size_t bytes, std::atomic<size_t> *current) {
  size_t cur = current->load();
  while (bytes > cur) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(bytes));
    if (e != cudaSuccess)
      return set_error(SNB_ERR_CUDA, "cannot reserve %zu bytes of shared memory: %s", bytes,
                       cudaGetErrorString(e));
    if (current->compare_exchange_weak(cur, bytes)) break;
  }
  return SNB_OK;
}
static std::atomic<size_t> g_fast_smem{0}, g_fast_smem4{0}, g_fast_smem_w400{0}, g_fast_smem4_w400{0},
    g_gen_smem{0};

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

static int compute_features_impl(const snb_plan *plan, const snb_batch *batch, const int16_t *d_pcm,
                                 const float *d_wave, int64_t capacity, uint64_t seed, void *d_out,
                                 int64_t ld_out, void *stream_, bool foreign_batch = false) {
  if (!plan || !batch || plan->kind != 0 || (batch->plan != plan && !foreign_batch))
    return set_error(SNB_ERR_VALUE, "plan/batch mismatch");
  if (batch->total_frames == 0) return SNB_OK;
  if ((!d_pcm && !d_wave) || !d_out) return set_error(SNB_ERR_VALUE, "null device buffer");
  if (capacity < batch->total_samples)
    return set_error(SNB_ERR_VALUE, "pcm buffer smaller than the batch (%lld < %lld samples)",
                     (long long)capacity, (long long)batch->total_samples);
  const FeatParams &p = plan->params;
  if (ld_out < p.dim) return set_error(SNB_ERR_VALUE, "ld_out %lld < dim %d", (long long)ld_out, p.dim);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  batch->note_stream(stream);
  if (plan->fast_path && !d_wave) {
    FastArgs a;
    a.p = p;
    fast_layout(plan, &a.sm);
    {
      int rc = ensure_tiles(batch->plan, batch, stream);     // (batch->plan: same framing for a foreign batch)
      if (rc != SNB_OK) return rc;
    }
    a.tiles = batch->d_tiles;
    a.ntiles = batch->ntiles;
    a.sample_begin = batch->d_sample_begin;
    a.sample_len = batch->d_sample_len;
    a.frame_offsets = batch->d_frame_offsets;
    a.mel_blobs = batch->d_mel_blobs;
    a.pcm = d_pcm;
    a.total_samples = capacity;
    a.out = d_out;
    a.ld_out = ld_out;
    a.seed = seed;
    static const bool no_tma = getenv("SNB_NO_TMA") != nullptr;
    a.use_tma = (!no_tma && (reinterpret_cast<uintptr_t>(d_pcm) % 16 == 0)) ? 1 : 0;
    // a segment of n bins meets at most (n + 6) / 8 + 1 chunks of 8 bins; the
    // widest possible segment is the whole spectrum
    a.mel_max_runs = batch->mel_max_runs > 0 ? batch->mel_max_runs : 33;
    auto launch = [&](auto kernel, std::atomic<size_t> *state) -> int {
      int rc = ensure_smem(kernel, a.sm.total, state);
      if (rc != SNB_OK) return rc;
      int per_sm = 1;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kFastThreads, a.sm.total);
      if (per_sm < 1) per_sm = 1;
      const int64_t grid = std::min<int64_t>(batch->ntiles, static_cast<int64_t>(num_sms()) * per_sm);
      kernel<<<static_cast<unsigned>(grid), kFastThreads, a.sm.total, stream>>>(a);
      return SNB_OK;
    };
    int rc;
    static const bool no_w400 = getenv("SNB_FUSED_W400") && atoi(getenv("SNB_FUSED_W400")) == 0;
    if (p.W == 400 && !no_w400)
      rc = (plan->fused_occ == 4) ? launch(fused_features_512_kernel<4, 400>, &g_fast_smem4_w400)
                                  : launch(fused_features_512_kernel<3, 400>, &g_fast_smem_w400);
    else
      rc = (plan->fused_occ == 4) ? launch(fused_features_512_kernel<4, 0>, &g_fast_smem4)
                                  : launch(fused_features_512_kernel<3, 0>, &g_fast_smem);
    if (rc != SNB_OK) return rc;
    SNB_LAUNCH_CHECK();
    return SNB_OK;
  }
  if (d_wave && p.B > 0)
    return set_error(SNB_ERR_UNSUPPORTED, "float32 input is only supported for the energy kind");
  GenArgs g;
  g.p = p;
  g.log2n = -1;
  if ((p.N & (p.N - 1)) == 0) {
    g.log2n = 0;
    while ((1 << g.log2n) < p.N) ++g.log2n;
  }
  const snb_feat_opts &xo = p.xo;
  const int tables = (xo.kind == SNB_FEAT_MFCC ? xo.num_ceps * p.B : 0) + xo.num_ceps +
                     (xo.kind == SNB_FEAT_PLP ? (xo.lpc_order + 1) * (p.B + 2) : 0);
  g.tables_floats = align_up(tables, 4);
  g.warp_floats = align_up(2 * p.N + 3 * (p.B + 2) + xo.lpc_order + 2 + 8, 4);
  g.sample_begin = batch->d_sample_begin;
  g.sample_len = batch->d_sample_len;
  g.frame_offsets = batch->d_frame_offsets;
  g.utt_mel_idx = batch->d_utt_mel;
  g.mel_blobs = batch->d_mel_blobs;
  g.pcm = d_pcm;
  g.pcm_f32 = d_wave;
  g.nutts = batch->nutts;
  g.total_frames = batch->total_frames;
  g.out = d_out;
  g.ld_out = ld_out;
  g.seed = seed;
  // (also reached with float input on a plan tiled for the fast path: energy)
  const size_t smem = static_cast<size_t>(g.tables_floats + kGenWarps * g.warp_floats) * 4;
  {
    int rc = ensure_smem(generic_features_kernel, smem, &g_gen_smem);
    if (rc != SNB_OK) return rc;
  }
  const int64_t want = (batch->total_frames + kGenWarps - 1) / kGenWarps;
  const int64_t grid = std::min<int64_t>(want, static_cast<int64_t>(num_sms()) * 8);
  generic_features_kernel<<<static_cast<unsigned>(grid), kGenWarps * 32, smem, stream>>>(g);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int64_t snb_feature_workspace_bytes(const snb_plan *plan, const snb_batch *batch) {
  if (!plan || !batch || plan->kind != 0 || !plan->rasta_mel_plan) return 0;
  return (batch->total_frames + 1) * static_cast<int64_t>(plan->params.B + 1) * 4 + 256;
}

extern "C" int snb_compute_features_ws(const snb_plan *plan, const snb_batch *batch, const int16_t *d_pcm,
                                       int64_t pcm_capacity, uint64_t seed, void *d_out, int64_t ld_out,
                                       void *d_workspace, int64_t workspace_bytes, void *stream_) {
  if (!plan || plan->kind != 0 || !plan->rasta_mel_plan)
    return compute_features_impl(plan, batch, d_pcm, nullptr, pcm_capacity, seed, d_out, ld_out, stream_);
  if (!batch || batch->plan != plan) return set_error(SNB_ERR_VALUE, "plan/batch mismatch");
  if (batch->total_frames == 0) return SNB_OK;
  if (!d_workspace || workspace_bytes < snb_feature_workspace_bytes(plan, batch))
    return set_error(SNB_ERR_VALUE, "RASTA-PLP needs a workspace of snb_feature_workspace_bytes() bytes");
  const FeatParams &p = plan->params;
  if (ld_out < p.dim) return set_error(SNB_ERR_VALUE, "ld_out %lld < dim %d", (long long)ld_out, p.dim);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float *mel = static_cast<float *>(d_workspace);
  const int64_t ld_mel = p.B + 1;
  // 1. [log-energy | linear mel energies] with the batch of the PLP plan (same
  //    framing, tiles and mel blobs as the internal filterbank plan)
  int rc = compute_features_impl(plan->rasta_mel_plan, batch, d_pcm, nullptr, pcm_capacity, seed, mel,
                                 ld_mel, stream_, /*foreign_batch=*/true);
  if (rc != SNB_OK) return rc;
  // 2. the frame-recursive filter, in place
  const int64_t nthreads = batch->nutts * p.B;
  rasta_kernel<<<static_cast<unsigned>((nthreads + 127) / 128), 128, 0, stream>>>(
      mel, ld_mel, 1, p.B, batch->d_frame_offsets, batch->nutts);
  SNB_LAUNCH_CHECK();
  // 3. PLP tail from the filtered energies
  PlpMelArgs a;
  a.p = p;
  a.mel = mel;
  a.ld_mel = ld_mel;
  a.frame_offsets = batch->d_frame_offsets;
  a.utt_mel_idx = batch->d_utt_mel;
  a.mel_blobs = batch->d_mel_blobs;
  a.nutts = batch->nutts;
  a.total_frames = batch->total_frames;
  a.out = static_cast<float *>(d_out);
  a.ld_out = ld_out;
  const snb_feat_opts &xo = p.xo;
  a.tables_floats = align_up(xo.num_ceps + (xo.lpc_order + 1) * (p.B + 2), 4);
  a.warp_floats = align_up(3 * (p.B + 2) + xo.lpc_order + 2 + 8, 4);
  if (plan->fast_path && batch->nblobs > 1)
    return set_error(SNB_ERR_UNSUPPORTED, "RASTA-PLP with several VTLN warps in one batch");
  const size_t smem = static_cast<size_t>(a.tables_floats + kGenWarps * a.warp_floats) * 4;
  static std::atomic<size_t> plp_smem{48 * 1024};
  rc = ensure_smem(plp_from_mel_kernel, smem, &plp_smem);
  if (rc != SNB_OK) return rc;
  const int64_t want = (batch->total_frames + kGenWarps - 1) / kGenWarps;
  const int64_t grid = std::min<int64_t>(want, static_cast<int64_t>(num_sms()) * 8);
  plp_from_mel_kernel<<<static_cast<unsigned>(grid), kGenWarps * 32, smem, stream>>>(a);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_compute_features(const snb_plan *plan, const snb_batch *batch, const int16_t *d_pcm,
                                    int64_t pcm_capacity, uint64_t seed, void *d_out, int64_t ld_out,
                                    void *stream) {
  if (plan && plan->kind == 0 && plan->rasta_mel_plan)
    return set_error(SNB_ERR_VALUE, "RASTA-PLP plans need snb_compute_features_ws()");
  return compute_features_impl(plan, batch, d_pcm, nullptr, pcm_capacity, seed, d_out, ld_out, stream);
}

extern "C" int snb_compute_features_f32(const snb_plan *plan, const snb_batch *batch, const float *d_wave,
                                        int64_t capacity, uint64_t seed, void *d_out, int64_t ld_out,
                                        void *stream) {
  return compute_features_impl(plan, batch, nullptr, d_wave, capacity, seed, d_out, ld_out, stream);
}
