// Internal declarations shared by the libkws.so translation units.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/kws.h"

namespace kws {

struct SteffeNet;
constexpr int L = KWS_SAMPLES;           // samples per clip
constexpr int MAX_BLOCKS = 13;           // depthwise-separable blocks after conv1d_1: 11 (exp 195 / 206 / 106) or 13 (conv_1d_time_sliced)
constexpr int TIMED_BLOCKS = 11;         // blocks with their own kws_timing_read slot
constexpr int NUM_SMS_B200 = 148;

struct ViewTable {                        // TTA views, passed to kernels by value
  int   n;
  int   shift[KWS_MAX_VIEWS];             // np.roll shift
  float gain[KWS_MAX_VIEWS];
};

struct LayerDesc {                        // one depthwise-separable block (model.py:34-52)
  int cin, cout, stride, pad_left, t_in, t_out;
};

struct Model {
  bool loaded = false;
  int arch = 0, classes = 0, c0 = 0, t0 = 0, t_last = 0, c_last = 0;
  bool dense1_bias = false, pool_max_avg = false;
  int n_blocks = 0;
  // head_kind 0: attention pooling head (model.py:819-830; exp 106: attention-weighted mean);
  //           1: GlobalAveragePooling1D -> Dense(hidden, no bias) -> ReLU6 -> Dense(classes, softmax)  (model.py:759-765)
  int head_kind = 0, hidden = 0;
  int c0_tc = 0;                          // conv1d_1 width of the tensor-core images: c0 rounded up to 64 (zero filters)
  LayerDesc layers[MAX_BLOCKS];
  // fp32 device weights (Keras layouts), all inside one allocation `blob`
  float* blob = nullptr;
  float* w_conv1 = nullptr;               // [120, c0]
  float* bn_scale[MAX_BLOCKS + 1];        // s = rsqrt(var+eps)*gamma
  float* bn_shift[MAX_BLOCKS + 1];        // beta - mean*s
  float* w_dw[MAX_BLOCKS];                // [3, cin]
  float* w_pw[MAX_BLOCKS];                // [cin, cout]
  float* tc_shift0 = nullptr;             // conv1d_1 BN shift padded to c0_tc (== bn_shift[0] when c0 is a multiple of 64)
  struct SteffeNet* steffe = nullptr;     // arch KWS_ARCH_STEFFENET: its own layer program (steffenet.cu)
  float* hidden_ws = nullptr;             // head_kind 1: [max rows][hidden] activations of the hidden Dense layer
  size_t hidden_ws_rows = 0;
  float* w_d1 = nullptr;                  // [t_last*c_last, t_last]  (head_kind 1: [c_last, hidden])
  float* b_d1 = nullptr;                  // [t_last] (zeros when the arch has no bias)
  float* w_d2 = nullptr;                  // [feat, classes]
  // tensor-core operand images (fp16, pre-swizzled UMMA K-major SW128 slabs)
  void* tc_blob = nullptr;
  __half* tc_conv1 = nullptr;             // [2 slabs][c0 rows][64] swizzled, 80 folded taps
  __half* tc_pw[MAX_BLOCKS];              // [cin/64 slabs][cout rows][64] swizzled, BN scale folded in
  __half* tc_dw[MAX_BLOCKS];              // depthwise taps [3][cin] fp16
  size_t max_act_elems = 0;               // max over layers of T*C (per clip-view)
};

struct Frontend {
  bool configured = false;
  int win = 0, hop = 0, n_fft = 0, n_bins = 0, frames = 0, n_mel = 0, n_keep = 0;
  // flavour: 0 = tf.contrib.signal chain of input_data.py:361-381 (magnitude, log(x + 1e-6));
  //          1 = native contrib_audio ops of audio.py:15-23 (power spectrogram, log(max(x, 1e-12)))
  int flavour = 0;
  float* blob = nullptr;
  float* dft_basis = nullptr;             // [win, 2*n_bins] interleaved (cos*w, -sin*w)
  float* mel_w = nullptr;                 // [n_bins, n_mel]
  float* dct_w = nullptr;                 // [n_mel, n_keep]
  // sparse mel (each bin feeds <= 2 filters)
  int*   mel_idx = nullptr;               // [n_bins][2]
  float* mel_val = nullptr;               // [n_bins][2]
  // tensor-core tier (tc_frontend.cu): split-fp16 basis blocks, banded-mel table, padded DCT
  void*  tc_blob = nullptr;
  bool   tc_ok = false;                   // this configuration is served by the tcgen05 kernel
  const uint8_t* tc_basis = nullptr;
  const float* tc_bin_tab = nullptr;
  const float* tc_dct = nullptr;
  const float* tc_hann = nullptr;         // [512] window, zero padded
  int tc_m_split = 0;
};

}  // namespace kws

namespace kws { class CopyPool; void copy_pool_destroy(CopyPool*); }

struct kws_handle {
  int device = 0;
  int max_rows = 0;
  int precision = KWS_PREC_TC;
  bool fuse_conv1_block1 = true;
  uint32_t smem_attr_done = 0;            // per-handle (= per-device) bits: MaxDynamicSharedMemorySize set for kernel k
  int num_sms = kws::NUM_SMS_B200;
  std::string err;
  int64_t launches = 0;
  // noise bank (caller-owned device memory)
  const float* bank = nullptr;
  int64_t bank_len = 0;
  int n_files = 0;
  int64_t* file_offsets_d = nullptr;      // device copy [n_files+1]
  kws::Frontend fe;
  kws::Model models[KWS_MAX_MODELS];
  // workspace
  void* act[2] = {nullptr, nullptr};      // ping-pong activations
  size_t act_bytes = 0;
  float* spec_ws = nullptr;               // frontend intermediates (fp32 path)
  size_t spec_ws_bytes = 0;
  float* mel_ws = nullptr;
  size_t mel_ws_bytes = 0;
  // host-entry staging (host_pipeline.cu): device slots, and pinned host slots for pageable caller buffers
  void* stage_d = nullptr; size_t stage_bytes = 0;
  void* pin_in[2] = {nullptr, nullptr};  size_t pin_in_bytes[2] = {0, 0};
  void* pin_out[2] = {nullptr, nullptr}; size_t pin_out_bytes[2] = {0, 0};
  kws::CopyPool* copy_pool = nullptr;     // worker threads of the pageable -> pinned staging copies
  int staging_mode = 0;                   // KWS_STAGING_*
  cudaEvent_t ev_user = nullptr;          // last use of the handle's workspace on a caller stream (device entry points)
  cudaStream_t own_stream = nullptr;      // compute stream of the host entry points
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
  // speed-TTA time stretch (stretch.cu): tables of the last rate + per-CTA STFT scratch, host-entry staging
  void* stretch_ws = nullptr; double stretch_rate = 0.0; int stretch_n_out = 0; bool stretch_attr_done = false;
  void* stretch_stage = nullptr; size_t stretch_stage_bytes = 0;
  // cuTensorMapEncodeTiled results keyed by (buffer, shape, box): the activation buffers and chunk sizes repeat from
  // call to call, so every launch after the first of a shape finds its tensor maps here (tc_net.cu)
  struct TmapEntry { const void* act; int c; long long d1, d2; int box_rows, swizzle, box_ch; alignas(64) unsigned char map[128]; };
  std::vector<TmapEntry> tmap_cache;
  // optional per-kernel-class device timing (bench.py roofline): event pairs around launches
  bool timing = false;
  struct TimedLaunch { int cls; cudaEvent_t e0, e1; };
  std::vector<TimedLaunch> timed;
  std::vector<cudaEvent_t> event_pool;
};

namespace kws {

int fail(kws_handle* h, int code, const std::string& msg);
bool debug_sync();   // KWS_DEBUG_SYNC=1: synchronise after every launch so a fault names its kernel
int ensure_bytes(kws_handle* h, void** p, size_t* cur, size_t need, bool pinned = false);

#define KWS_CUDA(h, expr)                                                              \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess)                                                             \
      return kws::fail((h), KWS_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

#define KWS_LAUNCH_CHECK(h)                                                            \
  do {                                                                                 \
    (h)->launches++;                                                                   \
    cudaError_t _e = cudaGetLastError();                                               \
    if (_e == cudaSuccess && kws::debug_sync()) _e = cudaDeviceSynchronize();          \
    if (_e != cudaSuccess)                                                             \
      return kws::fail((h), KWS_ECUDA, std::string("kernel launch at ") + __FILE__ + ":" + \
                                           std::to_string(__LINE__) + ": " + cudaGetErrorString(_e)); \
  } while (0)

// kernel classes for kws_timing_read
// classes 7..17 = the 11 depthwise+pointwise blocks individually (also summed into KC_BLOCKS)
enum { KC_AUGMENT = 0, KC_DFT = 1, KC_MELDCT = 2, KC_CONV1 = 3, KC_BLOCKS = 4, KC_HEAD = 5, KC_OTHER = 6, KC_COUNT = 7,
       KC_BLOCK0 = 7, KC_COUNT_EXT = 18 };
void timer_begin(kws_handle* h, int cls, cudaStream_t st);
void timer_end(kws_handle* h, cudaStream_t st);
#define KWS_T0(h, cls, st) do { if ((h)->timing) kws::timer_begin((h), (cls), (st)); } while (0)
#define KWS_T1(h, st) do { if ((h)->timing) kws::timer_end((h), (st)); } while (0)

// Device entry points that touch the handle's workspace (activations, front-end intermediates) record this event on
// the caller's stream; the host entry points, which run on the handle's own streams, wait for it first.
int mark_user_stream(kws_handle* h, cudaStream_t st);

// ---- dispatch helpers of api.cu used by host_pipeline.cu ----
int make_views(kws_handle* h, const int32_t* shift_h, const float* gain_h, int n, ViewTable* vt);
int forward_dispatch(kws_handle* h, int slot, const float* wav, int B, const ViewTable& vt, float* probs,
                     int32_t* argmax, cudaStream_t st);
int features_dispatch(kws_handle* h, const float* wav, int B, int kind, float* out, cudaStream_t st);
size_t feat_dim(const kws_handle* h, int kind);

// ---- launchers implemented in the .cu files ----
int launch_augment(kws_handle* h, const float* wav, const int16_t* pcm, float pcm_scale,
                   const int32_t* shift, const int32_t* bg_file, const int32_t* bg_off,
                   const float* bg_vol, const float* fg_vol, float* out, int B, int clamp,
                   cudaStream_t st);
int launch_dense_softmax_tta(kws_handle* h, const float* hid, int hidden, int n_views, int n_clips, const float* w2, int classes,
                             float* probs_mean, int32_t* argmax, cudaStream_t st);   // Dense(classes, softmax) per view + TTA mean + argmax
// steffeNet (model.py:1663-1726), fp32 CUDA-core kernels (steffenet.cu)
struct SteffeNet;
int steffe_build(kws_handle* h, Model& m, const kws_tensor_h* t, int n);
void steffe_free(SteffeNet* s);
int launch_forward_steffe(kws_handle* h, Model& m, const float* wav, int B, const ViewTable& vt, float* probs_mean, int32_t* argmax,
                          cudaStream_t st);
int launch_time_stretch(kws_handle* h, const int16_t* pcm, int B, double rate, float divisor, int16_t* out, cudaStream_t st);
int frontend_build(kws_handle* h, int win, int hop, int n_mel, int n_keep, float f_lo, float f_hi,
                   int sample_rate, int flavour = 0);
int launch_features_f32(kws_handle* h, const float* wav, int B, int kind, float* out, cudaStream_t st);
int launch_features_tc(kws_handle* h, const float* wav, int B, int kind, float* out, cudaStream_t st);
int model_build(kws_handle* h, int slot, int arch, const kws_tensor_h* t, int n);
// dbg_layer >= 0: stop after layer dbg_layer (0 = conv1d_1, i = block i) of the FIRST chunk and
// write that activation as fp32 [rows, T, C] to dbg_out (development / parity aid).
int launch_forward_f32(kws_handle* h, Model& m, const float* wav, int B, const ViewTable& vt,
                       float* probs_mean, int32_t* argmax, cudaStream_t st, int dbg_layer = -1,
                       float* dbg_out = nullptr);
int launch_forward_tc(kws_handle* h, Model& m, const float* wav, int B, const ViewTable& vt,
                      float* probs_mean, int32_t* argmax, cudaStream_t st, int dbg_layer = -1,
                      float* dbg_out = nullptr);
int launch_to_float(kws_handle* h, const void* src, bool src_half, float* dst, size_t n, cudaStream_t st);
int launch_head(kws_handle* h, Model& m, const void* act, bool act_half, int n_clips, int n_views,
                float* probs_mean, int32_t* argmax, cudaStream_t st);
int launch_convert(kws_handle* h, const float* probs, int B, int C_in, const int32_t* class_map_h,
                   int C_out, float* probs_out, uint8_t* probs_u8, cudaStream_t st);
int launch_select(kws_handle* h, const uint8_t* probs_u8, int B, int C, double thresh,
                  int32_t* label, uint8_t* keep, cudaStream_t st);
int launch_vote(kws_handle* h, const int32_t* labels, int M, int B, int min_count, int32_t* voted,
                uint8_t* clear, cudaStream_t st);

}  // namespace kws
