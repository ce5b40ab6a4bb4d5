/*
 * kws.h -- C ABI of libkws.so: the B200-native keyword-spotting hot path of
 * see--/speech_recognition (AudioProcessor waveform stage + STFT/log-mel/MFCC,
 * raw-waveform Depthwise1D network forward with TTA, pseudo-label selection,
 * voting).  Every entry point replaces a TensorFlow/Keras/NumPy call site of the
 * reference (file:line relative to the reference repository root); INTEGRATION.md
 * shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++ / torch types cross the boundary
 *   - every function returns 0 on success, a negative KWS_E* code on failure;
 *     kws_last_error() gives the message (per handle; handle==NULL -> the last
 *     kws_create failure of the calling thread)
 *   - pointers are CUDA DEVICE pointers unless the name ends in _h (host memory)
 *   - the caller owns every buffer it passes in; the handle owns weights, bases
 *     and workspace; one handle per (device, stream); a handle is not thread-safe
 *   - all work is enqueued on the cudaStream_t passed as `stream` (NULL = legacy
 *     default stream); device-pointer entry points never synchronise, *_host entry
 *     points return after their result is in the host buffer
 *   - there is NO CPU fallback: without a CUDA device kws_create fails
 */
#ifndef KWS_H_
#define KWS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KWS_ABI_VERSION 2

#define KWS_OK            0
#define KWS_EINVAL       -1   /* bad argument                                   */
#define KWS_ECUDA        -2   /* CUDA runtime error (message has the detail)    */
#define KWS_ENOMEM       -3   /* allocation failed                              */
#define KWS_ESTATE       -4   /* call order (e.g. forward before model_load)    */
#define KWS_EUNSUPPORTED -5   /* shape / architecture outside this path         */

#define KWS_SAMPLES 16000     /* desired_samples: 1 s @ 16 kHz, model.py:1798   */
#define KWS_MAX_VIEWS 16
#define KWS_MAX_MODELS 4

/* feature kinds (AudioProcessor.output_representation, input_data.py:441-449) */
#define KWS_FEAT_SPEC   0     /* spectrogram_  [B, frames, 257]  input_data.py:366 */
#define KWS_FEAT_LOGMEL 1     /* log(mel+1e-6) [B, frames, M]    input_data.py:378 */
#define KWS_FEAT_MFCC   2     /* mfcc_[..., :K][B, frames, K]    input_data.py:379-381 */

/* arithmetic of the dense contractions */
#define KWS_PREC_FP32   0     /* fp32 CUDA-core GEMMs: 1e-4 parity tier          */
#define KWS_PREC_TC     1     /* tcgen05 tensor-core GEMMs, fp32 accumulate in TMEM: network with fp16
                                 operands / fp16 activations (1e-2 parity tier), depthwise FIR accumulated in
                                 fp32; STFT with split-fp16 operands (keeps the fp32 tier's 1e-4)           */

#define KWS_ARCH_195 195      /* model.py:775-838 (also exp 206)                  */
#define KWS_ARCH_106 106      /* 32-class variant from the logs_106 graph         */
#define KWS_ARCH_TIME_SLICED 716   /* conv_1d_time_sliced_model(filter_mult=1), model.py:716-772: conv1d_1 with 32
                                      filters, 13 depthwise-separable blocks, GlobalAveragePooling1D -> Dense(256)
                                      -> ReLU6 -> Dense(num_classes) head; num_classes = dense_2/kernel.shape[1] */

#define KWS_ARCH_STEFFENET 1663    /* steffeNet, model.py:1663-1726 (Conv1D k75 s50 stem, residual depthwise-separable blocks
                                      up to 1536 channels, max || average pooling head).  Runs on the fp32 CUDA-core kernels in
                                      both precision tiers (channel counts beyond the tensor-core kernels' tiling). */

typedef struct kws_handle kws_t;

/* A host fp32 tensor addressed by its Keras variable name, e.g.
 * "conv1d_1/kernel", "batch_normalization_3/moving_variance",
 * "depthwise_conv2d_2/depthwise_kernel", "dense_1/bias"
 * (keras.models.load_model, make_submission.py:64-71). */
typedef struct {
  const char*  name;
  const float* data;     /* host, C-contiguous, Keras layout */
  int64_t      numel;
} kws_tensor_h;

/* ---- lifetime ---------------------------------------------------------- */
/* max_rows = clip-views processed per internal chunk (workspace is sized from it). */
int  kws_create(kws_t** h, int device, int max_rows);
void kws_destroy(kws_t* h);
const char* kws_last_error(const kws_t* h);
int  kws_abi_version(void);
int  kws_set_precision(kws_t* h, int precision);          /* KWS_PREC_*  (default TC) */
/* Tensor-core tier only: run conv1d_1 and the first depthwise-separable block (model.py:805-812) as ONE
 * kernel (default 1).  0 keeps them as two launches -- same results, more HBM traffic; used for A/B. */
int  kws_set_fusion(kws_t* h, int on);
/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
int64_t kws_launch_count(const kws_t* h);

/* Per-kernel-class device timing for roofline reports: when enabled every launch is bracketed
 * by CUDA events on its stream; kws_timing_read() synchronises, sums the elapsed milliseconds
 * and launch counts per class and resets.  Classes: 0 augment, 1 DFT/STFT GEMM, 2 mel+DCT,
 * 3 slice_conv1, 4 depthwise+pointwise blocks, 5 head, 6 other (n_classes >= 7); with n_classes >= 18,
 * slots 7..17 additionally hold the 11 blocks individually (tensor-core tier). */
int kws_timing_enable(kws_t* h, int on);
int kws_timing_read(kws_t* h, double* ms_per_class, int64_t* count_per_class, int n_classes);

/* ---- stage 1a: waveform augmentation  (input_data.py:338-359, utils.py:56-73) ---- */
/* background_data list of AudioProcessor (input_data.py:274-309) as one concatenated
 * device array; file_offsets_h[n_files+1] are the start indices of each wav. */
int kws_set_noise_bank(kws_t* h, const float* bank, const int64_t* file_offsets_h, int n_files);
/* out[b,t] = fl(bg[b,t]*bg_vol[b]) + fl(wav[b,(t-shift[b]) mod 16000]*fg_vol[b]);
 * bg[b] = bank[file_offsets[bg_file[b]] + bg_off[b] ...] or zeros when bg_file[b] < 0;
 * clamp != 0 applies the exp-106-era clip_by_value(-1,1) (input_data.py:356).
 * Replaces sess.run(background_clamp_) per clip (input_data.py:517-519). */
int kws_augment(kws_t* h, const float* wav, const int32_t* shift, const int32_t* bg_file,
                const int32_t* bg_off, const float* bg_vol, const float* fg_vol,
                float* out, int B, int clamp, void* stream);
/* Same with 16-bit PCM input decoded in the load: x = float(pcm) / divisor, divisor = 32768
 * (DecodeWav, input_data.py:334-336) or 32767 (make_submission_on_rpi.py:97). */
int kws_augment_pcm16(kws_t* h, const int16_t* pcm, float divisor, const int32_t* shift,
                      const int32_t* bg_file, const int32_t* bg_off, const float* bg_vol,
                      const float* fg_vol, float* out, int B, int clamp, void* stream);

/* ---- speed-TTA view: phase-vocoder time stretch  (create_tta_set.py:10-22, used by make_submission.py:131-140) ---- */
/* out[b] = np.int16(librosa.effects.time_stretch(np.float32(pcm[b]) / 32767, rate)[-16000:] * 32767): STFT 2048 / 512
 * (periodic Hann, centered, reflect-padded) -> phase vocoder -> ISTFT, the last 16000 samples, 16-bit PCM in and out
 * [B,16000] (the reference reads and writes WAV files here).  0 < rate <= 1 (the reference uses 0.9).  librosa is not
 * vendored by the reference: the algorithm is the published librosa 0.5.x one (parity unpinned, oracle/stretch.py). */
int kws_time_stretch_pcm16(kws_t* h, const int16_t* pcm, int B, double rate, int16_t* out, void* stream);
int kws_time_stretch_host_pcm16(kws_t* h, const int16_t* pcm_h, int B, double rate, int16_t* out_h);

/* ---- stage 1b: STFT -> |.| -> mel -> log -> DCT  (input_data.py:361-381) ---- */
/* model_settings keys of prepare_model_settings (model.py:1785-1829): window /
 * stride samples, dct_coefficient_count (= n_mel), num_log_mel_features (= n_keep). */
int kws_frontend_config(kws_t* h, int window_size_samples, int window_stride_samples,
                        int n_mel, int n_keep, float lower_edge_hertz, float upper_edge_hertz,
                        int sample_rate);
/* Native contrib_audio flavour of the same stage (audio.py:15-23, exp-106 graph nodes AudioSpectrogram / Mfcc):
 * contrib_audio.audio_spectrogram(window_size, stride, magnitude_squared=True) + contrib_audio.mfcc(
 * dct_coefficient_count; TF defaults lower 20 Hz, upper 4000 Hz, 40 filterbank channels).  After this call
 * kws_features(kind = SPEC) returns the POWER spectrogram, LOGMEL log(max(mel, 1e-12)) of the HTK-style bank
 * applied to sqrt(power), MFCC its sqrt(2/N)-normalised DCT-II.  TF's C++ kernels are un-vendored: the
 * algorithm follows spectrogram.cc / mfcc_mel_filterbank.cc / mfcc_dct.cc as published (parity unpinned). */
int kws_frontend_config_contrib(kws_t* h, int window_size, int stride, int sample_rate, float lower_hz,
                                float upper_hz, int filterbank_channels, int dct_coefficient_count);
int kws_frontend_frames(const kws_t* h);   /* spectrogram_length (98) */
/* Replaces sess.run(spectrogram_ | mfcc_) (input_data.py:520-531). */
int kws_features(kws_t* h, const float* wav, int B, int kind, float* out, void* stream);

/* ---- stage 2: network forward (+ TTA mean + argmax)  (model.py:775-838,
 *      make_submission.py:120-146) ---- */
int kws_model_load(kws_t* h, int slot, int arch, const kws_tensor_h* tensors_h, int n_tensors);
int kws_model_classes(const kws_t* h, int slot);
/* probs_mean[b,:] = (sum_v softmax_v) / n_views in the order given; argmax = first max.
 * view v of clip b is  view_gain[v] * roll(wav[b], view_shift[v])  (np.roll semantics,
 * make_submission.py:126-130).  Either output pointer may be NULL. */
int kws_forward(kws_t* h, int slot, const float* wav, int B, const int32_t* view_shift_h,
                const float* view_gain_h, int n_views, float* probs_mean, int32_t* argmax,
                void* stream);

/* Parity aid: run the network up to `layer` (0 = conv1d_1+BN+ReLU6, i = block i, 1..11) and
 * write that activation as fp32 [B*n_views, T_layer, C_layer] (needs B*n_views <= max_rows). */
int kws_debug_activation(kws_t* h, int slot, const float* wav, int B, const int32_t* view_shift_h,
                         const float* view_gain_h, int n_views, int layer, float* out, void* stream);

/* ---- driver math on the device ---- */
/* 32 -> 12 conversion (convert_from_see_v3_bugfix.py:76-110, freeze_graph_32_classes.py:55-69):
 * out[j] = max_{c: class_map[c]==j} p[c]; re-softmax exp(x)/sum (no max subtraction);
 * probs_u8 = trunc(p*255).  probs_out / probs_u8 may be NULL. */
int kws_convert_classes(kws_t* h, const float* probs, int B, int C_in, const int32_t* class_map_h,
                        int C_out, float* probs_out, uint8_t* probs_u8, void* stream);
/* create_pseudo_with_thresh.py:17-18,40-43: label = first argmax of the uint8 row,
 * keep = !(float32(max)/255 < thresh). */
int kws_select(kws_t* h, const uint8_t* probs_u8, int B, int C, double thresh,
               int32_t* label, uint8_t* keep, void* stream);
/* majority_vote.py:26-56 on integer labels [M,B] (M <= 16): most frequent label, ties to the
 * label first seen in the earliest submission; below min_count fall back to labels[0,b].
 * REPR_106_pseudo.py:12 unanimity == clear[b] with M=3, min_count=3. */
int kws_vote(kws_t* h, const int32_t* labels, int M, int B, int min_count,
             int32_t* voted, uint8_t* clear, void* stream);

/* ---- host-buffer entry points: what the reference-side binding calls ---- */
/* All of them run a chunked three-stream pipeline (H2D of chunk k+1 and D2H of chunk k-1 under the kernels of
 * chunk k) on the handle's own streams and return when the results are in the host buffers.  Host buffers may be
 * pinned (cudaHostAlloc / cudaHostRegister: copied by DMA straight from / to the caller's memory) or pageable (a
 * plain np.ndarray / malloc): pageable buffers are staged through pinned slots owned by the handle by a few host
 * threads (KWS_STAGE_THREADS, default 4), so the overlap holds for them too.  KWS_STAGING_AUTO detects which per
 * call (cudaPointerGetAttributes); the other two modes force the choice (A/B measurements). */
#define KWS_STAGING_AUTO   0
#define KWS_STAGING_ALWAYS 1
#define KWS_STAGING_NEVER  2
int kws_set_host_staging(kws_t* h, int mode);
/* Model.predict(x) + TTA (make_submission.py:120-146): wav_h [B,16000] host fp32 in,
 * probs_h [B,C] / argmax_h [B] host out. */
int kws_predict_host(kws_t* h, int slot, const float* wav_h, int B, const int32_t* view_shift_h,
                     const float* view_gain_h, int n_views, float* probs_h, int32_t* argmax_h);
/* Same with the clips in the wire format of the reference's WAV files: 16-bit PCM [B,16000], decoded on the device
 * as float(pcm) / divisor (32768: DecodeWav, input_data.py:334-336; 32767: the scipy paths,
 * make_submission_on_rpi.py:97, create_pseudo_with_thresh.py:48).  Half the host->device bytes of the fp32 form. */
int kws_predict_host_pcm16(kws_t* h, int slot, const int16_t* pcm_h, float divisor, int B,
                           const int32_t* view_shift_h, const float* view_gain_h, int n_views,
                           float* probs_h, int32_t* argmax_h);
/* AudioProcessor.get_data body (input_data.py:457-536) for pre-drawn parameters: host
 * waveforms + parameter arrays in, host representation out.  kind = -1 -> 'raw'. */
int kws_get_data_host(kws_t* h, const float* wav_h, const int32_t* shift_h, const int32_t* bg_file_h,
                      const int32_t* bg_off_h, const float* bg_vol_h, const float* fg_vol_h,
                      int B, int clamp, int kind, float* out_h);
/* The whole north-star hot path on one batch: augment -> log-mel/MFCC -> TTA forward.
 * Any of feat_h / probs_h / argmax_h may be NULL. */
int kws_pipeline_host(kws_t* h, int slot, const float* wav_h, const int32_t* shift_h,
                      const int32_t* bg_file_h, const int32_t* bg_off_h, const float* bg_vol_h,
                      const float* fg_vol_h, int B, int feat_kind, const int32_t* view_shift_h,
                      const float* view_gain_h, int n_views, float* feat_h, float* probs_h,
                      int32_t* argmax_h);

/* kws_pipeline_host with 16-bit PCM clips (see kws_predict_host_pcm16); the decode is fused into the augment
 * kernel's load (input_data.py:334-341).  Parameter arrays all NULL = decode only. */
int kws_pipeline_host_pcm16(kws_t* h, int slot, const int16_t* pcm_h, float divisor, const int32_t* shift_h,
                            const int32_t* bg_file_h, const int32_t* bg_off_h, const float* bg_vol_h,
                            const float* fg_vol_h, int B, int feat_kind, const int32_t* view_shift_h,
                            const float* view_gain_h, int n_views, float* feat_h, float* probs_h,
                            int32_t* argmax_h);

#ifdef __cplusplus
}
#endif
#endif /* KWS_H_ */
