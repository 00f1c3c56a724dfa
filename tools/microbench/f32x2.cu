// Microbenchmark: issue cost of packed fp32 (FFMA2/FADD2) against scalar FFMA on sm_100a
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu && ./f32x2
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

constexpr int ITERS = 4096;

// mode 0: 16 scalar FFMA per iteration; mode 1: 8 FFMA2 (same flops); mode 2: 16 FFMA + 8 IADD-like ALU ops;
// mode 3: 8 FFMA2 + 8 ALU ops
template <int MODE>
__global__ void kern(float *out, float s) {
  float a[16];
  u64 p[8];
  unsigned x[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) { p[i] = pk(a[2 * i], a[2 * i + 1]); x[i] = threadIdx.x + i; }
  const u64 ss = pk(s, s * 0.5f), tt = pk(0.25f, 0.125f);
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], s, 0.25f);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], ss, tt);
    }
    if (MODE >= 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) x[i] = (x[i] ^ (x[i] >> 3)) + it;   // 2 ALU ops each
    }
  }
  float acc = 0.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc += a[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) { float u, v; upk(p[i], u, v); acc += u + v + x[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
static void run(const char *name, float *d) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * 8, block = 256;
  kern<MODE><<<grid, block>>>(d, 0.999f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) kern<MODE><<<grid, block>>>(d, 0.999f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 5;
  const double fma = double(grid) * block * ITERS * 16.0;
  printf("%-28s %8.3f ms  %7.2f TFLOP/s fp32 (%s)\n", name, ms, 2.0 * fma / (ms * 1e-3) * 1e-12,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float *d;
  cudaMalloc(&d, 148 * 8 * 256 * 4);
  run<0>("16 FFMA", d);
  run<1>("8 FFMA2", d);
  run<2>("16 FFMA + 16 ALU", d);
  run<3>("8 FFMA2 + 16 ALU", d);
  return 0;
}
