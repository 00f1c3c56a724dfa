/* snb.h -- C ABI of libsnb.so, the B200 (sm_100a) speech-feature engine.
 *
 * This is the drop-in boundary for shennong's frame-based processor hot path.
 * In the reference the same boundary is the pykaldi FFI (CLIF-wrapped Kaldi
 * C++): every entry point below names the reference call site it replaces
 * (paths relative to bootphon/shennong @6febf5c).  The reference-side binding
 * a maintainer would add is a ctypes stub: see INTEGRATION.md and
 * shennong_b200/_lib.py.
 *
 * Conventions
 *   - plain C, no exceptions; every call returns an int status (SNB_OK == 0,
 *     negative = error) and snb_last_error() gives a thread-local message.
 *     SNB_ERR_OPTION is what Kaldi reports with KALDI_ERR (Python side raises
 *     RuntimeError, like pykaldi does); SNB_ERR_VALUE maps to ValueError.
 *   - the library never allocates or frees caller data: PCM, features and
 *     statistics live in caller-owned DEVICE buffers (PyTorch tensors'
 *     data_ptr() in the Python host).  Ragged batches are described by HOST
 *     offset arrays handed to snb_batch_create(), which owns the small device
 *     copies (offsets, tile table, per-utterance mel-table index).
 *   - plans are immutable after creation and may be shared between threads;
 *     all compute calls are asynchronous on the given cudaStream_t (passed as
 *     void*; NULL = default stream) and re-entrant per (plan, stream).
 *   - the caller selects the device (cudaSetDevice / torch.cuda.set_device)
 *     before creating plans/batches; objects are bound to that device.
 */
#ifndef SNB_H_
#define SNB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNB_VERSION 100

enum {
  SNB_OK = 0,
  SNB_ERR_OPTION = -1,   /* options Kaldi rejects (KALDI_ERR -> RuntimeError) */
  SNB_ERR_VALUE = -2,    /* bad argument (-> ValueError) */
  SNB_ERR_CUDA = -3,     /* CUDA runtime failure */
  SNB_ERR_UNSUPPORTED = -4
};

enum { SNB_WIN_HAMMING = 0, SNB_WIN_HANNING = 1, SNB_WIN_POVEY = 2,
       SNB_WIN_RECTANGULAR = 3, SNB_WIN_BLACKMAN = 4 };

enum { SNB_FEAT_SPECTROGRAM = 0, SNB_FEAT_FBANK = 1, SNB_FEAT_MFCC = 2,
       SNB_FEAT_PLP = 3, SNB_FEAT_ENERGY = 4 };

/* kaldi.feat.window.FrameExtractionOptions as filled by
 * shennong/processor/base.py:122-262 (milliseconds stored as float32). */
typedef struct snb_frame_opts {
  float samp_freq;
  float frame_shift_ms;
  float frame_length_ms;
  float dither;
  float preemph_coeff;
  float blackman_coeff;
  int32_t remove_dc_offset;
  int32_t window_type;            /* SNB_WIN_* */
  int32_t round_to_power_of_two;
  int32_t snip_edges;
} snb_frame_opts;

/* kaldi.feat.mel.MelBanksOptions (shennong/processor/base.py:288-374) */
typedef struct snb_mel_opts {
  int32_t num_bins;
  float low_freq;
  float high_freq;
  float vtln_low;
  float vtln_high;
} snb_mel_opts;

/* union of kaldi.feat.{mfcc.MfccOptions, fbank.FbankOptions,
 * spectrogram.SpectrogramOptions, plp.PlpOptions} and the energy options of
 * shennong/processor/energy.py:57-107 */
typedef struct snb_feat_opts {
  int32_t kind;                   /* SNB_FEAT_* */
  int32_t num_ceps;
  int32_t use_energy;
  float energy_floor;
  int32_t raw_energy;
  float cepstral_lifter;
  int32_t htk_compat;
  int32_t use_log_fbank;
  int32_t use_power;
  int32_t lpc_order;
  float compress_factor;
  float cepstral_scale;
  int32_t rasta;
  int32_t energy_compression;     /* 0 off, 1 log, 2 sqrt */
} snb_feat_opts;

/* kaldi.feat.pitch.PitchExtractionOptions (shennong/processor/pitch_kaldi.py:86-252) */
typedef struct snb_pitch_opts {
  float samp_freq, frame_shift_ms, frame_length_ms, preemph_coeff;
  float min_f0, max_f0, soft_min_f0, penalty_factor, lowpass_cutoff;
  float resample_freq, delta_pitch, nccf_ballast;
  int32_t lowpass_filter_width, upsample_filter_width;
  int32_t snip_edges;
} snb_pitch_opts;

/* kaldi.feat.pitch.ProcessPitchOptions (pitch_kaldi.py:321-489) */
typedef struct snb_pitch_post_opts {
  float pitch_scale, pov_scale, pov_offset, delta_pitch_scale;
  float delta_pitch_noise_stddev;
  int32_t normalization_left_context, normalization_right_context;
  int32_t delta_window, delay;
  int32_t add_pov_feature, add_normalized_log_pitch, add_delta_pitch,
      add_raw_log_pitch;
} snb_pitch_post_opts;

typedef struct snb_plan snb_plan;     /* options + immutable device tables */
typedef struct snb_batch snb_batch;   /* ragged batch descriptor */

/* ---- library ---------------------------------------------------------- */
int snb_version(void);
const char *snb_last_error(void);
/* number of kernels launched by this library since load (bench.py's
 * gpu_launches) */
int64_t snb_launch_count(void);

/* ---- host-side framing helpers ---------------------------------------- */
/* replaces kaldi.feat.window.FrameExtractionOptions.window_size()/
 * window_shift()/padded_window_size() (shennong/frames.py:101-108) */
int32_t snb_window_size(const snb_frame_opts *o);
int32_t snb_window_shift(const snb_frame_opts *o);
int32_t snb_padded_window_size(const snb_frame_opts *o);
/* replaces kaldi.feat.window.num_frames(n, opts, flush=True)
 * (shennong/frames.py:137-138, processor/energy.py:154); <0 on error */
int64_t snb_num_frames(int64_t nsamples, const snb_frame_opts *o);
/* replaces kaldi.feat.window.first_sample_of_frame (processor/plp.py:218) */
int64_t snb_first_sample_of_frame(int32_t frame, const snb_frame_opts *o);
/* replaces kaldi.feat.window.FeatureWindowFunction.from_options(opt).window
 * (shennong/window.py:107-114); writes window_size floats to host memory */
int snb_window_function(const snb_frame_opts *o, float *out, int32_t capacity);
/* replaces kaldi.feat.mel.MelBanks(mel_opts, frame_opts, vtln_warp)
 * (processor/plp.py:491-494): dense host copy for inspection/tests:
 * weights [num_bins, padded/2], center_freqs [num_bins] */
int snb_mel_banks_host(const snb_frame_opts *fo, const snb_mel_opts *mo,
                       float vtln_warp, float *weights, float *center_freqs);

/* ---- plans -------------------------------------------------------------- */
/* replaces kaldi.feat.{mfcc.Mfcc, fbank.Fbank, spectrogram.Spectrogram}(opts)
 * construction (processor/base.py:430, spectrogram.py:139), the PLP buffers
 * of processor/plp.py:443-480 and the energy setup of energy.py:148-165.
 * mo may be NULL for spectrogram/energy. */
int snb_feature_plan_create(const snb_frame_opts *fo, const snb_mel_opts *mo,
                            const snb_feat_opts *xo, snb_plan **out);
/* replaces kaldi.feat.pitch.PitchExtractionOptions use at pitch_kaldi.py:298 */
int snb_pitch_plan_create(const snb_pitch_opts *po, snb_plan **out);
void snb_plan_destroy(snb_plan *plan);
/* output columns of the plan (FeaturesProcessor.ndims) */
int32_t snb_plan_dim(const snb_plan *plan);
/* 1 when the tcgen-free fused fast path (512-point FFT) serves this plan */
int32_t snb_plan_uses_fast_path(const snb_plan *plan);

/* ---- ragged batches ----------------------------------------------------- */
/* sample_begin / sample_len: HOST int64[nutts], utterance u is
 * pcm[sample_begin[u] .. sample_begin[u] + sample_len[u]) of the packed int16
 * buffer.  Any begin works: the fused kernel bulk-copies (TMA) the 16-byte
 * aligned piece that covers a tile whenever the buffer itself is 16-byte
 * aligned; even begins keep the paired 32-bit sample loads.  The per-tile
 * table is expanded on the device by the first compute call, on its stream.
 * vtln_warps: HOST float[nutts] or NULL (all 1.0) -- the vtln_warp argument
 * of MelFeaturesProcessor.process (processor/base.py:376-406). */
int snb_batch_create(const snb_plan *plan, const int64_t *sample_begin,
                     const int64_t *sample_len, int64_t nutts,
                     const float *vtln_warps, snb_batch **out);
/* Stream-ordered variant for chunked pipelines: the descriptor upload is
 * queued on `stream` (no host wait, no cudaMalloc/cudaFree in the steady state:
 * device blob and pinned staging come from internal pools) and
 * snb_batch_destroy() recycles the blob once the work queued -- up to the
 * destroy call -- on `stream` and on every stream given to a compute call with
 * this batch has drained.  Contract: kernels that read the batch (including
 * post-processing calls given snb_batch_frame_offsets_device()) are queued on
 * one of those streams before the batch is destroyed. */
int snb_batch_create_on_stream(const snb_plan *plan, const int64_t *sample_begin,
                               const int64_t *sample_len, int64_t nutts,
                               const float *vtln_warps, void *stream,
                               snb_batch **out);
void snb_batch_destroy(snb_batch *batch);
int64_t snb_batch_num_utts(const snb_batch *batch);
int64_t snb_batch_total_frames(const snb_batch *batch);
/* HOST int64[nutts+1] frame offsets (rows of the packed output) */
const int64_t *snb_batch_frame_offsets(const snb_batch *batch);
/* DEVICE int64[nutts+1] copy of the same, for the post-processing calls */
const int64_t *snb_batch_frame_offsets_device(const snb_batch *batch);

/* ---- feature extraction ------------------------------------------------- */
/* replaces <Computer>(opts).compute(SubVector(int16 signal), vtln_warp)
 * (processor/base.py:427-431, spectrogram.py:137-140), the PLP python loop
 * (processor/plp.py:510-626) and the energy loop (energy.py:168-183), for a
 * whole batch in one fused launch.
 *   d_pcm : DEVICE int16[pcm_capacity] (16-byte aligned base for the TMA
 *           path); pcm_capacity = readable samples, >= max(begin + len)
 *   d_out : DEVICE float32[total_frames, ld_out] (energy kind: float64),
 *           plan dim columns written starting at column 0 of d_out.
 *   seed  : dither noise seed (ignored when dither == 0) */
int snb_compute_features(const snb_plan *plan, const snb_batch *batch,
                         const int16_t *d_pcm, int64_t pcm_capacity,
                         uint64_t seed, void *d_out, int64_t ld_out,
                         void *stream);
/* RASTA-PLP (shennong/processor/plp.py:64-146, 582-585) needs a DEVICE scratch
 * for the mel energies the frame-recursive filter runs on: this many bytes
 * (0 for every other plan) ... */
int64_t snb_feature_workspace_bytes(const snb_plan *plan,
                                    const snb_batch *batch);
/* ... handed to this variant of snb_compute_features (d_workspace may be NULL
 * when snb_feature_workspace_bytes() is 0) */
int snb_compute_features_ws(const snb_plan *plan, const snb_batch *batch,
                            const int16_t *d_pcm, int64_t pcm_capacity,
                            uint64_t seed, void *d_out, int64_t ld_out,
                            void *d_workspace, int64_t workspace_bytes,
                            void *stream);
/* same for float32 PCM (only the energy kind: shennong's EnergyProcessor does
 * NOT cast the signal to int16, energy.py:158) */
int snb_compute_features_f32(const snb_plan *plan, const snb_batch *batch,
                             const float *d_wave, int64_t capacity,
                             uint64_t seed, void *d_out, int64_t ld_out,
                             void *stream);

/* ---- post-processing (all on packed [total_frames, ld] matrices) -------- */
/* replaces kaldi.feat.functions.compute_deltas (postprocessor/delta.py:130) */
int snb_compute_deltas(const float *d_in, int64_t ld_in, int32_t dim,
                       const int64_t *d_frame_offsets, int64_t nutts,
                       int64_t total_frames, int32_t order, int32_t window,
                       float *d_out, int64_t ld_out, void *stream);
/* replaces kaldi.transform.cmvn.Cmvn.accumulate (postprocessor/cmvn.py:217-219)
 * per utterance: d_utt_stats DEVICE float64[nutts, 2, dim+1] is OVERWRITTEN.
 * d_weights: DEVICE float32[total_frames] or NULL. */
int snb_cmvn_accumulate(const float *d_feats, int64_t ld, int32_t dim,
                        const int64_t *d_frame_offsets, int64_t nutts,
                        const float *d_weights, double *d_utt_stats,
                        void *stream);
/* deterministic group reduction (per-speaker CMVN, pipeline_manager.py:75-85):
 * group g sums utterances d_group_utts[d_group_ptr[g] .. d_group_ptr[g+1]) in
 * order, ADDING into d_group_stats float64[ngroups, 2, dim+1]. */
int snb_cmvn_reduce_groups(const double *d_utt_stats, int32_t dim,
                           const int64_t *d_group_ptr,
                           const int64_t *d_group_utts, int64_t ngroups,
                           double *d_group_stats, void *stream);
/* float32 normalisation table from float64 stats, exactly ApplyCmvn /
 * ApplyCmvnReverse's per-dimension arithmetic (transform/cmvn.cc, reached from
 * cmvn.py:273-278): d_norm float32[ngroups, 2, dim], row 0 = offset, row 1 =
 * scale, so that y = x * scale + offset.  Groups whose count is < 1 get NaN
 * (the Python side raises ValueError like cmvn.py:254-257). */
int snb_cmvn_norm_from_stats(const double *d_stats, int64_t ngroups,
                             int32_t dim, int32_t norm_vars, int32_t reverse,
                             float *d_norm, void *stream);
/* replaces Cmvn.apply (cmvn.py:273-278).  d_utt_group: DEVICE int32[nutts]
 * index into d_norm (NULL: utterance u uses d_norm[u]). */
int snb_cmvn_apply(const float *d_in, int64_t ld_in, int32_t dim,
                   const int64_t *d_frame_offsets, int64_t nutts,
                   int64_t total_frames, const float *d_norm,
                   const int32_t *d_utt_group, float *d_out, int64_t ld_out,
                   void *stream);
/* fused pass 2 of the pipeline (pipeline.py:624-643): CMVN apply (skipped
 * when d_norm is NULL) followed by deltas, reading the base features once. */
int snb_cmvn_apply_deltas(const float *d_in, int64_t ld_in, int32_t dim,
                          const int64_t *d_frame_offsets, int64_t nutts,
                          int64_t total_frames, const float *d_norm,
                          const int32_t *d_utt_group, int32_t order,
                          int32_t window, float *d_out, int64_t ld_out,
                          void *stream);
/* replaces kaldi.feat.functions.sliding_window_cmn (cmvn.py:492) */
int snb_sliding_window_cmn(const float *d_in, int64_t ld_in, int32_t dim,
                           const int64_t *d_frame_offsets, int64_t nutts,
                           int64_t total_frames, int32_t center,
                           int32_t cmn_window, int32_t min_window,
                           int32_t normalize_variance, float *d_out,
                           int64_t ld_out, void *stream);
/* replaces kaldi.ivector.compute_vad_energy (postprocessor/vad.py:183);
 * column 0 of d_feats is the log-energy; d_out DEVICE float32[total_frames]
 * holding 0/1 (directly usable as CMVN weights, pipeline.py:588-596) */
int snb_vad_energy(const float *d_feats, int64_t ld,
                   const int64_t *d_frame_offsets, int64_t nutts,
                   int64_t total_frames, float energy_threshold,
                   float energy_mean_scale, int32_t frames_context,
                   float proportion_threshold, float *d_out, void *stream);
/* float64 -> float32 column copy (energy features -> VAD input) */
int snb_convert_f64_to_f32(const double *d_in, float *d_out, int64_t n,
                           void *stream);

/* ---- pitch --------------------------------------------------------------- */
/* frames compute_kaldi_pitch returns for nsamples (pitch_kaldi.py:298) */
int64_t snb_pitch_num_frames(int64_t nsamples, const snb_pitch_opts *po);
/* utterances the tracker follows at once (one warp each): batches of a
 * multiple of it keep every round of the Viterbi kernel full */
int64_t snb_pitch_wave_utts(const snb_plan *plan);
/* the same for n utterance lengths at once (HOST arrays): a corpus is planned
 * (row offsets of every utterance) before its first chunk is uploaded */
void snb_pitch_num_frames_array(const int64_t *nsamples, int64_t n,
                                const snb_pitch_opts *po, int64_t *out);
/* bytes of DEVICE scratch snb_compute_pitch needs for this batch */
int64_t snb_pitch_workspace_bytes(const snb_plan *plan, const snb_batch *batch);
/* replaces kaldi.feat.pitch.compute_kaldi_pitch(opts, wave)
 * (pitch_kaldi.py:296-299): d_out DEVICE float32[total_frames, ld_out],
 * columns (NCCF, pitch Hz). */
int snb_compute_pitch(const snb_plan *plan, const snb_batch *batch,
                      const int16_t *d_pcm, void *d_workspace,
                      int64_t workspace_bytes, float *d_out, int64_t ld_out,
                      void *stream);
int32_t snb_process_pitch_dim(const snb_pitch_post_opts *o);
/* replaces kaldi.feat.pitch.process_pitch(opts, raw) (pitch_kaldi.py:536-537)
 * d_out DEVICE float32[total_frames + nutts * delay, ld_out]: with delay = d
 * (>= 0; Kaldi asserts on a negative one) an utterance of F > 0 frames gives
 * F + d rows, row t = features of frame max(0, t - d), written from row
 * frame_offsets[u] + u * d (d = 0: the input geometry).  With
 * d_out_frame_offsets (DEVICE int64[nutts+1], may be NULL) the rows of
 * utterance u go to out_frame_offsets[u] instead and only the first
 * out_frame_offsets[u+1] - out_frame_offsets[u] of them are written: pasting
 * pitch next to features that have one or two frames less is the trim of
 * Features.concatenate(tolerance=2) (features.py:350-437, pipeline.py:639-641);
 * max_frames_per_utt = longest utterance of the batch in frames (launch
 * geometry); seed drives the delta-pitch noise (ignored when its stddev is 0) */
int snb_process_pitch(const snb_pitch_post_opts *o, const float *d_raw,
                      int64_t ld_raw, const int64_t *d_frame_offsets,
                      const int64_t *d_out_frame_offsets,
                      int64_t nutts, int64_t total_frames,
                      int64_t max_frames_per_utt, uint64_t seed, float *d_out,
                      int64_t ld_out, void *stream);

/* ---- collection: all-gather of the finished rows over NVLink ----
 * The reference collects the per-utterance results of its joblib workers in
 * the parent process (pipeline.py:541-567, processor/base.py:97-107).  One
 * process per GPU here: every rank owns a result buffer, its peers map it
 * through CUDA IPC, and a rank PUSHES its finished row blocks into the buffers
 * of all ranks with one kernel of plain stores over NVLink / NVSwitch. */
typedef struct snb_peer_handle { unsigned char bytes[64]; } snb_peer_handle;
/* device buffer of `bytes` + the handle other processes of the node open it with */
int snb_peer_buffer_create(int64_t bytes, void **d_ptr, snb_peer_handle *handle);
/* maps a peer's buffer in this process (enables peer access); close before exit */
int snb_peer_buffer_open(const snb_peer_handle *handle, void **d_ptr);
int snb_peer_buffer_close(void *d_ptr);
int snb_peer_buffer_destroy(void *d_ptr);
/* copies nfloats (multiple of 4) contiguous floats at d_src to
 * dst[p] + dst_offset_floats for p < ndst (HOST array of DEVICE pointers, own
 * buffer and mapped peer buffers alike; ndst <= 16), by `ctas` CTAs of 128
 * threads (<= 0: 296)
 * on `stream`.  The rows are visible to a peer once this launch has completed
 * and the ranks have synchronised. */
int snb_gather_rows(const float *d_src, int64_t nfloats, float *const *dst,
                    int32_t ndst, int64_t dst_offset_floats, int32_t ctas,
                    void *stream);

/* The same collection without the load/store pipes of the SMs.
 * snb_gather_rows_bulk: one-warp CTAs (<= 0: 148) whose lane 0 drives the TMA
 * unit: 8 KB pieces global -> shared -> every destination
 * (cp.async.bulk both ways), 24 KB of shared memory per CTA; meant to run on
 * a second stream UNDER the feature kernel of the next chunk.
 * snb_gather_rows_ce: one cudaMemcpyAsync per destination on internal
 * per-destination streams (copy engines), forked from and joined to `stream`.
 * A destination equal to the source (rows produced in place in the own
 * buffer) is skipped by both. */
int snb_gather_rows_bulk(const float *d_src, int64_t nfloats, float *const *dst,
                         int32_t ndst, int64_t dst_offset_floats, int32_t ctas,
                         void *stream);
int snb_gather_rows_ce(const float *d_src, int64_t nfloats, float *const *dst,
                       int32_t ndst, int64_t dst_offset_floats, void *stream);

/* ---- input step on the host ------------------------------------
 * The reference decodes every utterance with scipy.io.wavfile in its joblib
 * workers (audio.py:243-286) and slices [tstart, tstop] from the decoded array
 * (audio.py:520-561, utterances.py:171-176).  The payload of a RIFF/WAVE file
 * with PCM encoding, one channel and 16 bits IS the int16 signal of this path:
 * these two calls find it and read it into caller-owned (pinned) memory on
 * `nthreads` native threads.
 * snb_wav_scan_batch: data_offset[i] (bytes, -1 when file i is anything else
 * than mono 16-bit PCM or cannot be read), nsamples[i], rate[i]. */
int snb_wav_scan_batch(const char *const *paths, int64_t n, int64_t *data_offset,
                       int64_t *nsamples, int32_t *rate, int32_t nthreads);
/* reads nbytes[i] bytes at byte offsets[i] of file paths[i] into dst[i];
 * SNB_ERR_VALUE (and *first_failed = smallest failing index) on a short read */
int snb_read_segments(const char *const *paths, const int64_t *offsets,
                      const int64_t *nbytes, void *const *dst, int64_t n,
                      int32_t nthreads, int64_t *first_failed);

/* ---- sample-rate conversion on the device -----------------------
 * The reference resamples one utterance at a time on the host with sox or
 * scipy (audio.py:358-423).  Here: Kaldi's LinearResample (resample.cc, the
 * algorithm of kaldi::ResampleWaveform and of this path's pitch extractor) as
 * a polyphase filter over a packed int16 batch already on the device.
 * lowpass_cutoff <= 0: 0.99 * 0.5 * min(rate_in, rate_out); num_zeros <= 0: 6. */
typedef struct snb_resampler snb_resampler;
int snb_resampler_create(int32_t rate_in, int32_t rate_out, float lowpass_cutoff,
                         int32_t num_zeros, snb_resampler **out);
void snb_resampler_destroy(snb_resampler *r);
/* samples an utterance of `nsamples` gives (the resampler is flushed) */
int64_t snb_resampler_num_out(const snb_resampler *r, int64_t nsamples);
/* utterance u: d_pcm[d_begin[u] .. + d_len[u]) -> d_out[d_out_begin[u] .. +
 * snb_resampler_num_out(d_len[u])), as float32 (d_out_f32) and / or int16
 * truncated toward zero and saturated (d_out_i16) -- either may be NULL.
 * `max_out` bounds the outputs of one utterance, d_counts is an int64 scratch
 * of nutts elements (receives the output counts).  nutts <= 65535. */
int snb_resample_batch(const snb_resampler *r, const int16_t *d_pcm,
                       const int64_t *d_begin, const int64_t *d_len,
                       const int64_t *d_out_begin, int64_t nutts, int64_t max_out,
                       float *d_out_f32, int16_t *d_out_i16, int64_t *d_counts,
                       void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SNB_H_ */
