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

#include "device_utils.cuh"
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

// The same push with the TMA unit instead of the load/store pipes: ONE warp
// per CTA, lane 0 moves 8 KB pieces global -> shared (cp.async.bulk, mbarrier
// completion) -> every destination (cp.async.bulk shared -> global, one bulk
// group per piece) through a ring of kBulkStages buffers.  The CTA executes a
// few dozen instructions per piece, holds 32 registers per thread and 24 KB of
// shared memory: it fits next to the resident CTAs of the feature kernel and
// takes no issue slots from them, the data never passes through registers.
constexpr int kBulkStages = 3;
constexpr int kBulkPiece = 8192;          // bytes

struct BulkArgs {
  const char *src;
  int64_t nbytes;                         // multiple of 16
  char *dst[kMaxPeers];
  int32_t ndst;
};

__device__ __forceinline__ void bulk_copy_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}

__global__ void __launch_bounds__(32) peer_bulk_push_kernel(const BulkArgs a) {
  extern __shared__ __align__(128) unsigned char s_ring[];
  __shared__ __align__(8) uint64_t s_bar[kBulkStages];
  if (threadIdx.x != 0) return;
  for (int s = 0; s < kBulkStages; ++s) mbar_init(&s_bar[s], 1);
  fence_barrier_init();
  const int64_t npieces = (a.nbytes + kBulkPiece - 1) / kBulkPiece;
  const int64_t stride = gridDim.x;
  // prologue: the first kBulkStages - 1 loads
  int64_t next = blockIdx.x;              // next piece to load
  int lstage = 0;
  for (int s = 0; s < kBulkStages - 1 && next < npieces; ++s, next += stride) {
    const int64_t off = next * kBulkPiece;
    const uint32_t bytes = static_cast<uint32_t>(min(static_cast<int64_t>(kBulkPiece), a.nbytes - off));
    mbar_arrive_expect_tx(&s_bar[lstage], bytes);
    bulk_copy_g2s(s_ring + lstage * kBulkPiece, a.src + off, bytes, &s_bar[lstage]);
    lstage = (lstage + 1 == kBulkStages) ? 0 : lstage + 1;
  }
  uint32_t parity = 0;
  int stage = 0;
  for (int64_t piece = blockIdx.x; piece < npieces; piece += stride) {
    const int64_t off = piece * kBulkPiece;
    const uint32_t bytes = static_cast<uint32_t>(min(static_cast<int64_t>(kBulkPiece), a.nbytes - off));
    while (!mbar_try_wait(&s_bar[stage], (parity >> stage) & 1u)) __nanosleep(64);   // (no busy spin next to
    parity ^= 1u << stage;                                                           //  the feature kernel's warps)
    for (int p = 0; p < a.ndst; ++p) bulk_copy_s2g(a.dst[p] + off, s_ring + stage * kBulkPiece, bytes);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    // refill the buffer whose stores were committed one piece ago: only the
    // group just committed may still be reading shared memory
    if (next < npieces) {
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      const int64_t noff = next * kBulkPiece;
      const uint32_t nb = static_cast<uint32_t>(min(static_cast<int64_t>(kBulkPiece), a.nbytes - noff));
      mbar_arrive_expect_tx(&s_bar[lstage], nb);
      bulk_copy_g2s(s_ring + lstage * kBulkPiece, a.src + noff, nb, &s_bar[lstage]);
      lstage = (lstage + 1 == kBulkStages) ? 0 : lstage + 1;
      next += stride;
    }
    stage = (stage + 1 == kBulkStages) ? 0 : stage + 1;
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");    // writes performed before the CTA exits
}

// copy-engine form: per-destination streams and events of the calling device
struct CeState {
  cudaStream_t streams[kMaxPeers] = {};
  cudaEvent_t done[kMaxPeers] = {};
  cudaEvent_t ready = nullptr;
  bool init = false;
};
static std::mutex g_ce_mutex;
static std::map<int, CeState> g_ce;

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

extern "C" int snb_gather_rows_bulk(const float *d_src, int64_t nfloats, float *const *dst, int32_t ndst,
                                    int64_t dst_offset_floats, int32_t ctas, void *stream) {
  if (nfloats == 0 || ndst == 0) return SNB_OK;
  if (!d_src || !dst || ndst < 0 || ndst > kMaxPeers) return set_error(SNB_ERR_VALUE, "bad argument");
  if ((nfloats & 3) || (dst_offset_floats & 3) || (reinterpret_cast<uintptr_t>(d_src) & 15))
    return set_error(SNB_ERR_VALUE, "row blocks must be multiples of 16 bytes");
  BulkArgs a;
  a.src = reinterpret_cast<const char *>(d_src);
  a.nbytes = nfloats * 4;
  a.ndst = 0;
  for (int p = 0; p < ndst; ++p) {
    if (!dst[p] || (reinterpret_cast<uintptr_t>(dst[p]) & 15)) return set_error(SNB_ERR_VALUE, "bad destination");
    if (dst[p] + dst_offset_floats == d_src) continue;     // rows produced in place
    a.dst[a.ndst++] = reinterpret_cast<char *>(dst[p] + dst_offset_floats);
  }
  if (a.ndst == 0) return SNB_OK;
  if (ctas <= 0) ctas = 148;
  static std::atomic<bool> attr_set{false};
  if (!attr_set.exchange(true)) {
    cudaFuncSetAttribute(peer_bulk_push_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    cudaGetLastError();
  }
  peer_bulk_push_kernel<<<static_cast<unsigned>(ctas), 32, kBulkStages * kBulkPiece,
                          static_cast<cudaStream_t>(stream)>>>(a);
  SNB_LAUNCH_CHECK();
  return SNB_OK;
}

extern "C" int snb_gather_rows_ce(const float *d_src, int64_t nfloats, float *const *dst, int32_t ndst,
                                  int64_t dst_offset_floats, void *stream) {
  if (nfloats == 0 || ndst == 0) return SNB_OK;
  if (!d_src || !dst || ndst < 0 || ndst > kMaxPeers) return set_error(SNB_ERR_VALUE, "bad argument");
  int dev = 0;
  SNB_CUDA_CHECK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_ce_mutex);
  CeState &st = g_ce[dev];
  if (!st.init) {
    for (int p = 0; p < kMaxPeers; ++p) {
      SNB_CUDA_CHECK(cudaStreamCreateWithFlags(&st.streams[p], cudaStreamNonBlocking));
      SNB_CUDA_CHECK(cudaEventCreateWithFlags(&st.done[p], cudaEventDisableTiming));
    }
    SNB_CUDA_CHECK(cudaEventCreateWithFlags(&st.ready, cudaEventDisableTiming));
    st.init = true;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SNB_CUDA_CHECK(cudaEventRecord(st.ready, s));
  const size_t bytes = static_cast<size_t>(nfloats) * 4;
  for (int p = 0; p < ndst; ++p) {
    if (!dst[p]) return set_error(SNB_ERR_VALUE, "bad destination");
    if (dst[p] + dst_offset_floats == d_src) continue;     // rows produced in place
    SNB_CUDA_CHECK(cudaStreamWaitEvent(st.streams[p], st.ready, 0));
    SNB_CUDA_CHECK(cudaMemcpyAsync(dst[p] + dst_offset_floats, d_src, bytes, cudaMemcpyDeviceToDevice,
                                   st.streams[p]));
    SNB_CUDA_CHECK(cudaEventRecord(st.done[p], st.streams[p]));
    SNB_CUDA_CHECK(cudaStreamWaitEvent(s, st.done[p], 0));
  }
  return SNB_OK;
}
