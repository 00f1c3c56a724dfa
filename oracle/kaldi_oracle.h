/* kaldi_oracle.h -- CPU oracle for the shennong frame-based feature hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (shennong_b200/) may
 * include, link or call this.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * The reference (bootphon/shennong @6febf5c) delegates the arithmetic of this
 * path to Kaldi through pykaldi (conda package shennong-pykaldi, NOT version
 * pinned: /root/reference/environment.yml:7) and Kaldi is not vendored in
 * /root/reference.  This file restates Kaldi's published algorithms
 * (src/feat/{feature-window,feature-functions,mel-computations,feature-mfcc,
 * feature-fbank,feature-spectrogram,feature-plp,pitch-functions,resample}.cc,
 * src/transform/cmvn.cc, src/ivector/voice-activity-detection.cc) anchored on
 * the reference's own call sites and its in-tree Python restatement of
 * ExtractWindow/ProcessWindow/PlpComputer (shennong/processor/plp.py:149-260,
 * 510-626).  Parity pinning: see oracle/README.md (fbank/mfcc/spectrogram are
 * pinned against torchaudio.compliance.kaldi golden vectors; PLP, RASTA-PLP
 * and energy against the outputs of the reference's own plp.py / energy.py run
 * over a pykaldi shim; Kaldi pitch is "parity unpinned").
 */
#ifndef KALDI_ORACLE_H_
#define KALDI_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_WIN_HAMMING = 0, ORC_WIN_HANNING = 1, ORC_WIN_POVEY = 2,
       ORC_WIN_RECTANGULAR = 3, ORC_WIN_BLACKMAN = 4 };

enum { ORC_FEAT_SPECTROGRAM = 0, ORC_FEAT_FBANK = 1, ORC_FEAT_MFCC = 2,
       ORC_FEAT_PLP = 3, ORC_FEAT_ENERGY = 4 };

/* kaldi FrameExtractionOptions (shennong/processor/base.py:122-262) */
typedef struct {
  float samp_freq;
  float frame_shift_ms;
  float frame_length_ms;
  float dither;
  float preemph_coeff;
  float blackman_coeff;
  int32_t remove_dc_offset;
  int32_t window_type;
  int32_t round_to_power_of_two;
  int32_t snip_edges;
} orc_frame_opts;

/* kaldi MelBanksOptions (shennong/processor/base.py:288-374) */
typedef struct {
  int32_t num_bins;
  float low_freq;
  float high_freq;
  float vtln_low;
  float vtln_high;
} orc_mel_opts;

/* union of Mfcc/Fbank/Spectrogram/Plp/Energy options */
typedef struct {
  int32_t kind;           /* ORC_FEAT_* */
  int32_t num_ceps;       /* mfcc, plp */
  int32_t use_energy;     /* mfcc, fbank, plp */
  float energy_floor;
  int32_t raw_energy;
  float cepstral_lifter;  /* mfcc, plp */
  int32_t htk_compat;
  int32_t use_log_fbank;  /* fbank */
  int32_t use_power;      /* fbank */
  int32_t lpc_order;      /* plp */
  float compress_factor;  /* plp */
  float cepstral_scale;   /* plp */
  int32_t rasta;          /* plp (shennong/processor/plp.py:64-146) */
  int32_t energy_compression; /* energy: 0 off, 1 log, 2 sqrt */
} orc_feat_opts;

typedef struct {
  float samp_freq, frame_shift_ms, frame_length_ms, preemph_coeff;
  float min_f0, max_f0, soft_min_f0, penalty_factor, lowpass_cutoff;
  float resample_freq, delta_pitch, nccf_ballast;
  int32_t lowpass_filter_width, upsample_filter_width;
  int32_t snip_edges;
  int32_t recompute_frame;
} orc_pitch_opts;

typedef struct {
  float pitch_scale, pov_scale, pov_offset, delta_pitch_scale;
  float delta_pitch_noise_stddev;
  int32_t normalization_left_context, normalization_right_context;
  int32_t delta_window, delay;
  int32_t add_pov_feature, add_normalized_log_pitch, add_delta_pitch,
      add_raw_log_pitch;
} orc_pitch_post_opts;

/* framing */
int32_t orc_window_size(const orc_frame_opts *o);
int32_t orc_window_shift(const orc_frame_opts *o);
int32_t orc_padded_window_size(const orc_frame_opts *o);
int64_t orc_num_frames(int64_t nsamples, const orc_frame_opts *o);
int64_t orc_first_sample_of_frame(int32_t frame, const orc_frame_opts *o);
void orc_window_function(const orc_frame_opts *o, float *out /*[window_size]*/);

/* mel banks: dense [num_bins, padded/2] weights, centre freqs [num_bins].
 * Returns 0, or -1 for options Kaldi rejects with KALDI_ERR. */
int32_t orc_mel_banks(const orc_frame_opts *fo, const orc_mel_opts *mo,
                      float vtln_warp, float *weights, float *center_freqs);

/* output dimension of a feature kind, or -1 on invalid options */
int32_t orc_feat_dim(const orc_frame_opts *fo, const orc_mel_opts *mo,
                     const orc_feat_opts *xo);

/* wave is the int16 PCM already converted to float (energy: raw scale).
 * out is [num_frames, dim] float32 (energy kind: float64 stored in out64).
 * Returns number of frames, or <0 on option errors. */
int64_t orc_compute_features(const float *wave, int64_t nsamples,
                             const orc_frame_opts *fo, const orc_mel_opts *mo,
                             const orc_feat_opts *xo, float vtln_warp,
                             float *out, double *out64);

/* batch version for the CPU baseline: OpenMP over utterances */
int64_t orc_compute_features_batch(const int16_t *pcm,
                                   const int64_t *sample_offsets,
                                   const int64_t *frame_offsets, int64_t nutts,
                                   const orc_frame_opts *fo,
                                   const orc_mel_opts *mo,
                                   const orc_feat_opts *xo, float *out,
                                   int32_t nthreads);

/* post-processing */
void orc_compute_deltas(const float *in, int64_t nframes, int32_t dim,
                        int32_t order, int32_t window, float *out);
void orc_cmvn_accumulate(const float *feats, int64_t nframes, int32_t dim,
                         const float *weights /*or NULL*/,
                         double *stats /*[2, dim+1], updated*/);
int32_t orc_cmvn_apply(const double *stats, int32_t dim, int32_t norm_vars,
                       int32_t reverse, float *feats, int64_t nframes);
void orc_sliding_window_cmn(const float *in, int64_t nframes, int32_t dim,
                            int32_t center, int32_t cmn_window,
                            int32_t min_window, int32_t normalize_variance,
                            float *out);
void orc_vad_energy(const float *feats, int64_t nframes, int32_t dim,
                    float energy_threshold, float energy_mean_scale,
                    int32_t frames_context, float proportion_threshold,
                    float *out /*[nframes] 0/1*/);

/* pitch */
/* Kaldi LinearResample of a whole signal, flushed (kaldi::ResampleWaveform);
 * cutoff <= 0 / num_zeros <= 0 select Kaldi's defaults */
int64_t orc_resample_num_out(int64_t n_in, int32_t rate_in, int32_t rate_out);
void orc_resample(const float *in, int64_t n_in, int32_t rate_in, int32_t rate_out,
                  float cutoff, int32_t num_zeros, float *out);
int64_t orc_pitch_num_frames(int64_t nsamples, const orc_pitch_opts *o);
int32_t orc_pitch_num_lags(const orc_pitch_opts *o);
int64_t orc_compute_kaldi_pitch(const float *wave, int64_t nsamples,
                                const orc_pitch_opts *o,
                                float *out /*[nframes,2]*/);
int32_t orc_process_pitch_dim(const orc_pitch_post_opts *o);
int64_t orc_process_pitch(const float *raw /*[nframes,2]*/, int64_t nframes,
                          const orc_pitch_post_opts *o,
                          float *out /*[nframes+delay, dim]*/);

/* full pipeline batch for the CPU baseline (config 3):
 * mfcc/fbank/plp -> per-utterance cmvn -> deltas */
int64_t orc_pipeline_batch(const int16_t *pcm, const int64_t *sample_offsets,
                           const int64_t *frame_offsets, int64_t nutts,
                           const orc_frame_opts *fo, const orc_mel_opts *mo,
                           const orc_feat_opts *xo, int32_t do_cmvn,
                           int32_t norm_vars, int32_t delta_order,
                           int32_t delta_window, float *out, int32_t nthreads);

#ifdef __cplusplus
}
#endif
#endif
