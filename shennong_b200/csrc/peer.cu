// peer.cu -- collection of the result rows over NVLink peer memory.
//
// utterances shard over the GPUs of one box with no collective on
// the data path; the only exchange is the all-gather of the finished rows at
// collection time (the reference collects joblib results in the parent process,
// shennong/pipeline.py:541-567).  Here every rank owns a result buffer that
// its peers map through CUDA IPC, and a rank PUSHES its finished rows into the
// buffers of all ranks with plain stores that travel over NVLink / NVSwitch:
//
//   snb_peer_buffer_create / open / close / destroy   buffers + IPC handles
//   snb_gather_rows      one kernel: a block of rows -> every peer buffer
//
// One process per GPU; handles are exchanged by the host side (torch.distributed
// object all-gather, shennong_b200/distributed.py).  The stores of a launch are
// visible to the owner of the destination once the launch has completed and
// the ranks have synchronised (stream order + the step's barrier).
#include <cstring>

#include "snb_internal.h"

namespace snb {

constexpr int kMaxPeers = 16;

struct ScatterArgs {
  const float4 *src;
  int64_t n4;                   // float4 elements
  float4 *dst[kMaxPeers];
  int32_t ndst;
};

// Every element is read once (L2 / HBM) and stored to each destination.  The
// CTAs are SMALL on purpose (128 threads, 32 registers): the launch runs on a
// second stream next to the feature kernel of the following chunk, whose
// persistent CTAs leave 4 096 registers and 34 KB of shared memory free on
// every SM -- a 512-thread CTA did not fit there and the collection ran after
// the compute instead of under it (bench at N = 2: 14.9 ms per step against
// 12.2 for the chunked compute alone).  Posted 16-byte stores, two independent
// elements in flight per thread and destination.
__global__ void __launch_bounds__(128, 16) peer_scatter_kernel(const ScatterArgs a) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + stride < a.n4; i += 2 * stride) {
    const float4 v0 = __ldcs(a.src + i), v1 = __ldcs(a.src + i + stride);
    for (int p = 0; p < a.ndst; ++p) {
      float4 *d = a.dst[p];
      d[i] = v0; d[i + stride] = v1;
    }
  }
  for (; i < a.n4; i += stride) {
    const float4 v = __ldcs(a.src + i);
    for (int p = 0; p < a.ndst; ++p) a.dst[p][i] = v;
  }
}

}  // namespace snb

using namespace snb;

extern "C" int snb_peer_buffer_create(int64_t bytes, void **d_ptr, snb_peer_handle *handle) {
  if (!d_ptr || !handle || bytes <= 0) return set_error(SNB_ERR_VALUE, "bad argument");
  *d_ptr = nullptr;
  void *p = nullptr;
  SNB_CUDA_CHECK(cudaMalloc(&p, static_cast<size_t>(bytes)));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    cudaFree(p);
    return set_error(SNB_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
  }
  static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(snb_peer_handle), "handle size");
  std::memset(handle, 0, sizeof(*handle));
  std::memcpy(handle, &h, sizeof(h));
  *d_ptr = p;
  return SNB_OK;
}

extern "C" int snb_peer_buffer_open(const snb_peer_handle *handle, void **d_ptr) {
  if (!d_ptr || !handle) return set_error(SNB_ERR_VALUE, "bad argument");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  SNB_CUDA_CHECK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return SNB_OK;
}

extern "C" int snb_peer_buffer_close(void *d_ptr) {
  if (d_ptr) SNB_CUDA_CHECK(cudaIpcCloseMemHandle(d_ptr));
  return SNB_OK;
}

extern "C" int snb_peer_buffer_destroy(void *d_ptr) {
  if (d_ptr) SNB_CUDA_CHECK(cudaFree(d_ptr));
  return SNB_OK;
}

extern "C" int snb_gather_rows(const float *d_src, int64_t nfloats, float *const *dst, int32_t ndst,
                               int64_t dst_offset_floats, int32_t ctas, void *stream) {
  if (nfloats == 0 || ndst == 0) return SNB_OK;
  if (!d_src || !dst || ndst < 0 || ndst > kMaxPeers) return set_error(SNB_ERR_VALUE, "bad argument");
  if ((nfloats & 3) || (dst_offset_floats & 3) || (reinterpret_cast<uintptr_t>(d_src) & 15))
    return set_error(SNB_ERR_VALUE, "row blocks must be multiples of 16 bytes");
  ScatterArgs a;
  a.src = reinterpret_cast<const float4 *>(d_src);
  a.n4 = nfloats / 4;
  a.ndst = ndst;
  for (int p = 0; p < ndst; ++p) {
    if (!dst[p] || (reinterpret_cast<uintptr_t>(dst[p]) & 15)) return set_error(SNB_ERR_VALUE, "bad destination");
    a.dst[p] = reinterpret_cast<float4 *>(dst[p] + dst_offset_floats);
  }
  if (ctas <= 0) ctas = 296;
  // An SM keeps one L1 / shared-memory split while CTAs are resident: ask for
  // the split of the feature kernel (maximum shared memory) so that these CTAs
  // can join its SMs instead of waiting for them to drain.
  static std::atomic<bool> carveout_set{false};
  if (!carveout_set.exchange(true)) {
    cudaFuncSetAttribute(peer_scatter_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    cudaGetLastError();
  }
  peer_scatter_kernel<<<static_cast<unsigned>(ctas), 128, 0, static_cast<cudaStream_t>(stream)>>>(a);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}
