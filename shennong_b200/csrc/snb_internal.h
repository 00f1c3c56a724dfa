// snb_internal.h -- shared declarations of libsnb.so (not part of the ABI)
#ifndef SNB_INTERNAL_H_
#define SNB_INTERNAL_H_

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/snb.h"

namespace snb {

// ---- error handling --------------------------------------------------------
int set_error(int code, const char *fmt, ...);
extern std::atomic<int64_t> g_launch_count;

#define SNB_CUDA_CHECK(expr)                                              \
  do {                                                                    \
    cudaError_t err__ = (expr);                                           \
    if (err__ != cudaSuccess)                                             \
      return snb::set_error(SNB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, \
                            cudaGetErrorString(err__), __FILE__, __LINE__); \
  } while (0)

#define SNB_LAUNCH_CHECK()                                                  \
  do {                                                                      \
    snb::g_launch_count.fetch_add(1, std::memory_order_relaxed);            \
    cudaError_t err__ = cudaGetLastError();                                 \
    if (err__ != cudaSuccess)                                               \
      return snb::set_error(SNB_ERR_CUDA, "kernel launch failed: %s (%s:%d)", \
                            cudaGetErrorString(err__), __FILE__, __LINE__);  \
  } while (0)

// Host -> device upload that is COMPLETE on return for every stream: a plain
// cudaMemcpy from pageable memory may return once the data is staged, with the
// DMA only ordered in the legacy stream -- kernels launched on non-blocking
// streams (PyTorch's) could then read the destination too early.
cudaError_t upload(void *dst, const void *src, size_t bytes);
// the calling thread's private non-blocking stream used by upload()
cudaError_t upload_stream(cudaStream_t *out);

// ---- host-side tables (tables.cc) ------------------------------------------
int32_t window_size(const snb_frame_opts &o);
int32_t window_shift(const snb_frame_opts &o);
int32_t padded_window_size(const snb_frame_opts &o);
int64_t num_frames(int64_t nsamples, const snb_frame_opts &o);
int64_t first_sample_of_frame(int64_t frame, const snb_frame_opts &o);
void window_function(const snb_frame_opts &o, std::vector<float> *out);

struct MelBanksHost {
  int32_t num_bins = 0, num_fft_bins = 0;
  std::vector<int32_t> first, size;     // per mel bin
  std::vector<float> weights;           // concatenated ranges
  std::vector<int32_t> offset;          // start of each range in weights
  std::vector<float> center_freqs;
  // segment view: segment s in [0, B] holds the FFT bins whose mel value lies
  // in (center[s-1], center[s]]; such a bin feeds the rising side of
  // triangle s (weight up) and the falling side of triangle s-1 (weight down)
  std::vector<int32_t> seg_first, seg_size;   // [B+1]
  std::vector<float> up, down;                // [num_fft_bins]
  std::vector<int32_t> seg_of;                // [num_fft_bins] segment of a bin, -1: outside every triangle
};
// returns SNB_OK or SNB_ERR_OPTION (sets error)
int build_mel_banks(const snb_frame_opts &fo, const snb_mel_opts &mo,
                    float vtln_warp, MelBanksHost *out);
void build_dct(int32_t num_ceps, int32_t num_bins, std::vector<float> *out);
void build_lifter(int32_t num_ceps, float q, std::vector<float> *out);
void build_equal_loudness(const std::vector<float> &center_freqs,
                          std::vector<float> *out);
void build_idft_bases(int32_t n_bases, int32_t dimension,
                      std::vector<float> *out);

// ---- device-side views -------------------------------------------------------
// mel "blob": one contiguous float/int block per VTLN warp value (words):
//   [ loudness[B] | chunk_w[32*20] | chunk_meta[32] | run_first[B+2] | longest run list |   <- fused kernel (smem)
//     first[B] | size[B] | offset[B] | weights[wcap] ]                   <- generic kernel (global memory)
// chunk_*: FFT bins dealt to the 32 lanes of a warp in contiguous chunks of 8
// (features.cu, "mel energies"); only built when the FFT has 256 bins.
struct FeatTables {
  const float *window;     // [W]
  const float2 *tw_half;   // [N/2]  e^{-2 pi i m/(N/2)}  (fast path: W256^m)
  const float2 *tw_full;   // [N/2]  e^{-2 pi i k/N}      (fast path: W512^k)
  const float *dct;        // [num_ceps, B]
  const float *lifter;     // [num_ceps]
  const float *idft;       // [lpc_order+1, B+2]
  const float2 *tw_dft;    // [N] generic non-pow2 DFT twiddles
};

struct FeatParams {
  snb_frame_opts fo;
  snb_feat_opts xo;
  int32_t W, S, N;            // window, shift, padded (fft) size
  int32_t B;                  // mel bins (0 if none)
  int32_t dim;                // output columns
  int32_t mel_blob_stride;    // in 4-byte words
  int32_t mel_wcap;           // weights capacity per blob
  int32_t mel_chunk_off;      // word offset (multiple of 4) of chunk_w
  int32_t mel_meta_off;       // word offset of chunk_meta
  int32_t mel_run_off;        // word offset of run_first
  int32_t mel_fast_words;     // words the fused kernel stages (multiple of 4)
  int32_t mel_gen_off;        // word offset of first[] (generic-path sections)
  int32_t need_raw_energy, need_post_energy;
  float log_energy_floor;
  float eps_energy;           // FLT_EPSILON (Kaldi) or DBL_EPSILON (plp.py)
  FeatTables t;
};

struct PitchTables;          // pitch.cu

struct alignas(16) TileDesc {   // one CTA work item: frames [f0, f0+nf) of utt (32 bytes)
  int32_t utt;
  int32_t f0;
  int32_t nf;
  int32_t mel_idx;
  int64_t row0;                // first output row of the tile
  int64_t g0;                  // index in the packed PCM buffer of the tile's first sample when its
                               // whole span lies inside the utterance (bulk-copyable), else -1
};

}  // namespace snb

struct snb_plan {
  int kind;                  // 0 feature plan, 1 pitch plan
  int device;
  // feature plan
  snb_frame_opts fo;
  snb_mel_opts mo;
  snb_feat_opts xo;
  bool has_mel = false;
  bool fast_path = false;
  snb::FeatParams params;
  int32_t tile_frames = 32;
  int32_t fused_occ = 3;     // CTAs per SM the fused kernel variant is compiled for
  void *d_tables = nullptr;  // one allocation holding all fixed tables
  size_t smem_bytes = 0;
  // VTLN mel-blob cache (host side), keyed by warp bits
  mutable std::mutex mu;
  mutable std::map<uint32_t, std::vector<int32_t>> mel_blobs;
  // RASTA-PLP: internal plan producing linear mel energies (+ log-energy)
  snb_plan *rasta_mel_plan = nullptr;
  // pitch plan
  snb_pitch_opts po;
  snb::PitchTables *pitch = nullptr;
};

namespace snb {
// A block of device (or pinned host) memory on loan from a MemPool; `pending`
// are the events that must complete before the block may be handed out again
struct PoolBlock {
  void *p = nullptr;
  size_t cap = 0;
  int device = 0;
  std::vector<cudaEvent_t> pending;
};

// Recycles the small per-batch allocations so that the steady state of a
// chunked pipeline performs no cudaMalloc / cudaFree / cudaHostAlloc (cudaFree
// synchronises the whole device).  Thread-safe.
class MemPool {
 public:
  explicit MemPool(bool pinned) : pinned_(pinned) {}
  cudaError_t acquire(size_t bytes, int device, PoolBlock *out);
  void release(PoolBlock block);       // takes ownership of block.pending
 private:
  static bool ready(PoolBlock *b);
  void free_block(PoolBlock *b);
  bool pinned_;
  std::mutex mu_;
  std::vector<PoolBlock> free_;
};
MemPool &device_pool();
MemPool &pinned_pool();
}  // namespace snb

struct snb_batch {
  const snb_plan *plan;
  // stream-ordered batches (snb_batch_create_on_stream): pooled blob, and the
  // streams whose queued work must drain before the blob is recycled
  bool pooled = false;
  snb::PoolBlock dev_block;
  mutable std::mutex stream_mu;
  mutable std::vector<cudaStream_t> streams;
  void note_stream(cudaStream_t s) const {
    if (!pooled) return;
    std::lock_guard<std::mutex> lock(stream_mu);
    for (cudaStream_t t : streams) if (t == s) return;
    streams.push_back(s);
  }
  int64_t nutts = 0, total_frames = 0, total_samples = 0;
  std::vector<int64_t> sample_begin, sample_len, frame_offsets;
  void *d_blob = nullptr;            // single device allocation; the pointers below alias it
  int64_t *d_sample_begin = nullptr, *d_sample_len = nullptr, *d_frame_offsets = nullptr;
  snb::TileDesc *d_tiles = nullptr;      // fused path: expanded on the device from d_tile_first
  int64_t ntiles = 0;
  int64_t *d_tile_first = nullptr;       // [nutts + 1] first tile of every utterance
  mutable bool tiles_ready = false;      // guarded by stream_mu (lazy expansion at the first launch)
  mutable cudaEvent_t tiles_event = nullptr;
  mutable cudaStream_t tiles_stream = nullptr;
  int32_t *d_utt_mel = nullptr;          // [nutts] mel blob of every utterance
  int32_t *d_mel_blobs = nullptr;  // [nblobs, mel_blob_stride]
  int32_t nblobs = 0;
  int32_t mel_max_runs = 0;        // longest per-segment run list over the blobs (fused mel stage)
  // pitch
  std::vector<int64_t> down_offsets;   // per-utt offsets in downsampled signal
  int64_t *d_down_offsets = nullptr;
  int64_t total_down = 0;
};

namespace snb {
// launchers implemented in the .cu files
int launch_features(const snb_plan *plan, const snb_batch *batch,
                    const int16_t *d_pcm, uint64_t seed, void *d_out,
                    int64_t ld_out, cudaStream_t stream);
int feature_plan_finalize(snb_plan *plan);   // picks path, smem size
int pitch_plan_init(snb_plan *plan);
void pitch_plan_free(snb_plan *plan);
int pitch_batch_init(const snb_plan *plan, snb_batch *batch);
}  // namespace snb

#endif
