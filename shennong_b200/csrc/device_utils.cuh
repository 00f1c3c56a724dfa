// device_utils.cuh -- small device helpers shared by the kernels
#ifndef SNB_DEVICE_UTILS_CUH_
#define SNB_DEVICE_UTILS_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

namespace snb {

#define SNB_FULL_MASK 0xffffffffu

// all-reduce sum over aligned lane groups of size G (xor butterfly: every
// lane ends with the bit-identical total)
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(SNB_FULL_MASK, v, o);
  return v;
}
template <int G>
__device__ __forceinline__ double group_sum_f64(double v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(SNB_FULL_MASK, v, o);
  return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + TMA 1-D bulk copy (cp.async.bulk -> SASS UBLKCP) -----------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(count)
               : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *dst_smem, const void *src_gmem,
                                              uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

// ---- cp.async (LDGSTS): 16-byte global -> shared copies that occupy no registers
__device__ __forceinline__ void cp_async_16(void *dst_smem, const void *src_gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- counter-based noise for dither ----------------------------------------
// splitmix64 finaliser: two independent 32-bit uniforms per (seed, frame, pair)
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
// Box-Muller: two N(0,1) samples from one 64-bit hash
__device__ __forceinline__ void gauss_pair(uint64_t seed, uint64_t frame,
                                           uint32_t pair, float *g0, float *g1) {
  const uint64_t h = mix64(seed + 0x9e3779b97f4a7c15ull * (frame * 4096ull + pair + 1ull));
  const uint32_t a = static_cast<uint32_t>(h), b = static_cast<uint32_t>(h >> 32);
  const float u1 = (static_cast<float>(a >> 8) + 1.0f) * (1.0f / 16777216.0f);  // (0,1]
  const float u2 = static_cast<float>(b >> 8) * (1.0f / 16777216.0f);           // [0,1)
  const float r = sqrtf(-2.0f * __logf(u1));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  *g0 = r * c;
  *g1 = r * s;
}

// fused fast path: one 32-bit hash per PAIR of samples (16 bits per uniform:
// |g| <= 4.86), see dither_pair
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du;
  x ^= x >> 15; x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
// hash32 is a bijection: the keys of the first 2^32 frames of a batch are all
// different (no birthday collisions between frames)
__device__ __forceinline__ uint32_t frame_noise_key(uint64_t seed, uint64_t frame) {
  const uint32_t lo = static_cast<uint32_t>(seed), hi = static_cast<uint32_t>(seed >> 32);
  return hash32(hash32(static_cast<uint32_t>(frame) ^ lo) + static_cast<uint32_t>(frame >> 32)) ^ hi;
}

// Two dither samples (dither * N(0,1)) for the fused kernel: Box-Muller on the
// MUFU unit only.  One 32-bit hash per pair: its high half, spliced into the
// mantissa of 1.0f, gives u1 = (k + 0.5) / 65536 without an int->float
// conversion; the low half is the angle.  c = -2 ln2 dither^2 folds the change
// of base of lg2, Box-Muller's -2 and the dither amplitude into one multiply
// (r = dither sqrt(-2 ln u1) = sqrt(c lg2 u1)).  No operand can be denormal,
// so the .ftz approximations are used as they are (the generic logf / rsqrtf
// expansions spend 7 instructions on denormal fix-ups).
__device__ __forceinline__ float2 dither_pair(uint32_t key, uint32_t pair, float c) {
  // (xor, not add: with key + pair * G two frames whose keys differ by m * G
  // would share their whole sequence shifted by m pairs -- likely among 1e7
  // frames; with xor only isolated samples can coincide)
  const uint32_t h = hash32(key ^ (pair * 0x9e3779b9u));
  const float u1 = __uint_as_float(0x3f800040u | ((h >> 9) & 0x007fff80u)) - 1.0f;   // (0, 1)
  float l, r, sn, cs;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u1));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l * c));
  const float ang = static_cast<float>(h & 0xffffu) * (6.283185307179586f / 65536.0f);
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(ang));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(ang));
  return make_float2(r * cs, r * sn);
}

// ---- packed fp32 pairs (sm_100a add/mul/fma.f32x2 -> SASS FADD2/FMUL2/FFMA2) ----
// One instruction for two floats held in an aligned 64-bit register pair: the
// same flop rate as the scalar forms at half the issue slots (measured:
// tools/microbench/f32x2.cu).  pk / upk are register renames, not moves.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float lo, float hi) {
  f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void upk(f32x2 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}

// the same operations with a compile-time choice between the packed
// instruction and two scalar ones on the halves (P = false)
template <bool P> __device__ __forceinline__ f32x2 add2t(f32x2 a, f32x2 b) {
  if (P) return add2(a, b);
  float a0, a1, b0, b1; upk(a, a0, a1); upk(b, b0, b1); return pk(a0 + b0, a1 + b1);
}
template <bool P> __device__ __forceinline__ f32x2 sub2t(f32x2 a, f32x2 b) {
  if (P) return sub2(a, b);
  float a0, a1, b0, b1; upk(a, a0, a1); upk(b, b0, b1); return pk(a0 - b0, a1 - b1);
}
template <bool P> __device__ __forceinline__ f32x2 mul2t(f32x2 a, f32x2 b) {
  if (P) return mul2(a, b);
  float a0, a1, b0, b1; upk(a, a0, a1); upk(b, b0, b1); return pk(a0 * b0, a1 * b1);
}
template <bool P> __device__ __forceinline__ f32x2 fma2t(f32x2 a, f32x2 b, f32x2 c) {
  if (P) return fma2(a, b, c);
  float a0, a1, b0, b1, c0, c1; upk(a, a0, a1); upk(b, b0, b1); upk(c, c0, c1);
  return pk(fmaf(a0, b0, c0), fmaf(a1, b1, c1));
}

// radix-4 butterfly on packed complex values (re, im); pm = (1, -1), mp = (-1, 1).
// (a1 - a3) * (-i) is formed directly in swapped order by two scalar
// subtractions, so that the two outputs that need it are one FFMA2 each.
template <bool P>
__device__ __forceinline__ void bfly4p(f32x2 a0, f32x2 a1, f32x2 a2, f32x2 a3, f32x2 &y0, f32x2 &y1,
                                       f32x2 &y2, f32x2 &y3, f32x2 pm, f32x2 mp) {
  const f32x2 t0 = add2t<P>(a0, a2), t1 = sub2t<P>(a0, a2), t2 = add2t<P>(a1, a3);
  float a1r, a1i, a3r, a3i;
  upk(a1, a1r, a1i); upk(a3, a3r, a3i);
  const float dr = a1r - a3r, di = a1i - a3i;        // d = a1 - a3
  y0 = add2t<P>(t0, t2); y2 = sub2t<P>(t0, t2);
  if (P) {
    const f32x2 dsw = pk(di, dr);                    // formed directly in swapped order
    y1 = fma2(dsw, pm, t1);                          // t1 - i d
    y3 = fma2(dsw, mp, t1);                          // t1 + i d
  } else {
    float t1r, t1i;
    upk(t1, t1r, t1i);
    y1 = pk(t1r + di, t1i - dr);
    y3 = pk(t1r - di, t1i + dr);
  }
}

// in-register 16-point DFT (forward, e^{-2 pi i nk/16}), radix 4x4,
// natural-order output, on packed complex values: 104 instructions (40 FADD2 +
// 16 FFMA2 + 48 scalar) instead of the 160 of the scalar form (P = false),
// bit-identical results
template <bool P>
__device__ __forceinline__ void fft16p(f32x2 (&x)[16]) {
  const float C1 = 0.92387953251128674f;  // cos(pi/8)
  const float S1 = 0.38268343236508977f;  // sin(pi/8)
  const float R2 = 0.70710678118654752f;  // sqrt(1/2)
  const f32x2 pm = pk(1.0f, -1.0f), mp = pk(-1.0f, 1.0f);
  f32x2 b[16];
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2)
    bfly4p<P>(x[n2], x[4 + n2], x[8 + n2], x[12 + n2], b[n2 * 4], b[n2 * 4 + 1], b[n2 * 4 + 2], b[n2 * 4 + 3], pm, mp);
  {
    float r, i;
    upk(b[5], r, i);  b[5] = pk(r * C1 + i * S1, i * C1 - r * S1);
    upk(b[6], r, i);  b[6] = pk((r + i) * R2, (i - r) * R2);
    upk(b[7], r, i);  b[7] = pk(r * S1 + i * C1, i * S1 - r * C1);
    upk(b[9], r, i);  b[9] = pk((r + i) * R2, (i - r) * R2);
    upk(b[10], r, i); b[10] = pk(i, -r);
    upk(b[11], r, i); b[11] = pk((i - r) * R2, -(r + i) * R2);
    upk(b[13], r, i); b[13] = pk(r * S1 + i * C1, i * S1 - r * C1);
    upk(b[14], r, i); b[14] = pk((i - r) * R2, -(r + i) * R2);
    upk(b[15], r, i); b[15] = pk(-(r * C1 + i * S1), -(i * C1 - r * S1));
  }
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
    bfly4p<P>(b[k1], b[4 + k1], b[8 + k1], b[12 + k1], x[k1], x[k1 + 4], x[k1 + 8], x[k1 + 12], pm, mp);
}

}  // namespace snb
#endif
