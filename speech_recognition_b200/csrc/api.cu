// C ABI of libkws.so (include/kws.h): argument validation, handle lifetime and dispatch to the kernel
// launchers (the host-buffer entry points live in host_pipeline.cu).  No CPU fallback exists anywhere below.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>

#include "common.cuh"

namespace kws {

static thread_local std::string g_create_error;

bool debug_sync() {
  static const bool on = [] { const char* e = getenv("KWS_DEBUG_SYNC"); return e && e[0] == '1'; }();
  return on;
}

int fail(kws_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_create_error = msg;
  return code;
}

int ensure_bytes(kws_handle* h, void** p, size_t* cur, size_t need, bool pinned) {
  if (*cur >= need && *p) return KWS_OK;
  if (*p) { if (pinned) cudaFreeHost(*p); else cudaFree(*p); *p = nullptr; *cur = 0; }
  cudaError_t e = pinned ? cudaMallocHost(p, need) : cudaMalloc(p, need);
  if (e != cudaSuccess) {
    *p = nullptr;
    return fail(h, KWS_ENOMEM, std::string("allocation of ") + std::to_string(need) + " bytes failed: " +
                                   cudaGetErrorString(e));
  }
  *cur = need;
  return KWS_OK;
}

void timer_begin(kws_handle* h, int cls, cudaStream_t st) {
  auto get = [&]() {
    cudaEvent_t e;
    if (!h->event_pool.empty()) { e = h->event_pool.back(); h->event_pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
  };
  kws_handle::TimedLaunch t{cls, get(), get()};
  cudaEventRecord(t.e0, st);
  h->timed.push_back(t);
}
void timer_end(kws_handle* h, cudaStream_t st) { cudaEventRecord(h->timed.back().e1, st); }

int make_views(kws_handle* h, const int32_t* shift_h, const float* gain_h, int n, ViewTable* vt) {
  if (n <= 0 || n > KWS_MAX_VIEWS) return fail(h, KWS_EINVAL, "n_views must be in 1..16");
  vt->n = n;
  for (int i = 0; i < KWS_MAX_VIEWS; ++i) { vt->shift[i] = 0; vt->gain[i] = 1.0f; }
  for (int i = 0; i < n; ++i) {
    vt->shift[i] = shift_h ? shift_h[i] : 0;
    vt->gain[i] = gain_h ? gain_h[i] : 1.0f;
  }
  return KWS_OK;
}

int forward_dispatch(kws_handle* h, int slot, const float* wav, int B, const ViewTable& vt,
                            float* probs, int32_t* argmax, cudaStream_t st) {
  if (slot < 0 || slot >= KWS_MAX_MODELS || !h->models[slot].loaded)
    return fail(h, KWS_ESTATE, "kws_forward before kws_model_load for this slot");
  if (B < 0) return fail(h, KWS_EINVAL, "negative batch");
  if (B == 0) return KWS_OK;
  if (!wav) return fail(h, KWS_EINVAL, "null waveform pointer");
  if (h->models[slot].arch == KWS_ARCH_STEFFENET) return launch_forward_steffe(h, h->models[slot], wav, B, vt, probs, argmax, st);
  return h->precision == KWS_PREC_FP32 ? launch_forward_f32(h, h->models[slot], wav, B, vt, probs, argmax, st)
                                       : launch_forward_tc(h, h->models[slot], wav, B, vt, probs, argmax, st);
}

int features_dispatch(kws_handle* h, const float* wav, int B, int kind, float* out, cudaStream_t st) {
  if (!h->fe.configured) return fail(h, KWS_ESTATE, "kws_features before kws_frontend_config");
  if (kind < KWS_FEAT_SPEC || kind > KWS_FEAT_MFCC) return fail(h, KWS_EINVAL, "unknown feature kind");
  if (B < 0) return fail(h, KWS_EINVAL, "negative batch");
  if (B == 0) return KWS_OK;
  if (!wav || !out) return fail(h, KWS_EINVAL, "null pointer");
  return h->precision == KWS_PREC_FP32 ? launch_features_f32(h, wav, B, kind, out, st)
                                       : launch_features_tc(h, wav, B, kind, out, st);
}

int mark_user_stream(kws_handle* h, cudaStream_t st) {
  if (!h->ev_user) KWS_CUDA(h, cudaEventCreateWithFlags(&h->ev_user, cudaEventDisableTiming));
  KWS_CUDA(h, cudaEventRecord(h->ev_user, st));
  return KWS_OK;
}

size_t feat_dim(const kws_handle* h, int kind) {
  const Frontend& fe = h->fe;
  const int d = kind == KWS_FEAT_SPEC ? fe.n_bins : (kind == KWS_FEAT_LOGMEL ? fe.n_mel : fe.n_keep);
  return static_cast<size_t>(fe.frames) * d;
}

}  // namespace kws

using namespace kws;

extern "C" {

int kws_abi_version(void) { return KWS_ABI_VERSION; }

const char* kws_last_error(const kws_t* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int64_t kws_launch_count(const kws_t* h) { return h ? h->launches : 0; }

int kws_create(kws_t** out, int device, int max_rows) {
  if (!out) return fail(nullptr, KWS_EINVAL, "null handle pointer");
  *out = nullptr;
  if (max_rows <= 0) return fail(nullptr, KWS_EINVAL, "max_rows must be positive");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(nullptr, KWS_ECUDA, std::string("no CUDA device: libkws has no CPU fallback (") +
                                        cudaGetErrorString(e) + ")");
  if (device < 0 || device >= n) return fail(nullptr, KWS_EINVAL, "device index out of range");
  if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(nullptr, KWS_ECUDA, cudaGetErrorString(e));
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(nullptr, KWS_ECUDA, cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, KWS_EUNSUPPORTED, std::string("libkws is built for sm_100a (B200); found ") + prop.name);
  kws_handle* h = new (std::nothrow) kws_handle();
  if (!h) return fail(nullptr, KWS_ENOMEM, "out of host memory");
  h->device = device; h->max_rows = max_rows; h->num_sms = prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
    delete h; return fail(nullptr, KWS_ECUDA, cudaGetErrorString(e));
  }
  *out = h;
  return KWS_OK;
}

void kws_destroy(kws_t* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (int i = 0; i < KWS_MAX_MODELS; ++i) {
    if (h->models[i].blob) cudaFree(h->models[i].blob);
    if (h->models[i].tc_blob) cudaFree(h->models[i].tc_blob);
    if (h->models[i].hidden_ws) cudaFree(h->models[i].hidden_ws);
    if (h->models[i].steffe) steffe_free(h->models[i].steffe);
  }
  if (h->fe.blob) cudaFree(h->fe.blob);
  if (h->fe.tc_blob) cudaFree(h->fe.tc_blob);
  if (h->file_offsets_d) cudaFree(h->file_offsets_d);
  for (int i = 0; i < 2; ++i) if (h->act[i]) cudaFree(h->act[i]);
  if (h->spec_ws) cudaFree(h->spec_ws);
  if (h->mel_ws) cudaFree(h->mel_ws);
  if (h->stage_d) cudaFree(h->stage_d);
  for (int i = 0; i < 2; ++i) {
    if (h->pin_in[i]) cudaFreeHost(h->pin_in[i]);
    if (h->pin_out[i]) cudaFreeHost(h->pin_out[i]);
  }
  if (h->stretch_ws) cudaFree(h->stretch_ws);
  if (h->stretch_stage) cudaFree(h->stretch_stage);
  if (h->copy_pool) copy_pool_destroy(h->copy_pool);
  if (h->ev_user) cudaEventDestroy(h->ev_user);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  if (h->h2d_stream) cudaStreamDestroy(h->h2d_stream);
  if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
  for (int i = 0; i < 2; ++i) {
    if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]);
    if (h->ev_comp[i]) cudaEventDestroy(h->ev_comp[i]);
    if (h->ev_d2h[i]) cudaEventDestroy(h->ev_d2h[i]);
  }
  for (auto& t : h->timed) { cudaEventDestroy(t.e0); cudaEventDestroy(t.e1); }
  for (auto e : h->event_pool) cudaEventDestroy(e);
  delete h;
}

int kws_timing_enable(kws_t* h, int on) {
  if (!h) return KWS_EINVAL;
  h->timing = on != 0;
  return KWS_OK;
}

int kws_timing_read(kws_t* h, double* ms_per_class, int64_t* count_per_class, int n_classes) {
  if (!h) return KWS_EINVAL;
  if (n_classes < KC_COUNT || !ms_per_class || !count_per_class) return fail(h, KWS_EINVAL, "need 7 class slots");
  for (int i = 0; i < n_classes; ++i) { ms_per_class[i] = 0.0; count_per_class[i] = 0; }
  for (auto& t : h->timed) {
    KWS_CUDA(h, cudaEventSynchronize(t.e1));
    float ms = 0.f;
    KWS_CUDA(h, cudaEventElapsedTime(&ms, t.e0, t.e1));
    int cls = t.cls;
    if (cls >= KC_BLOCK0) {                                    // per-block slot when the caller asked for them
      if (cls < n_classes) { ms_per_class[cls] += ms; count_per_class[cls]++; }
      cls = KC_BLOCKS;
    }
    ms_per_class[cls] += ms; count_per_class[cls]++;
    h->event_pool.push_back(t.e0); h->event_pool.push_back(t.e1);
  }
  h->timed.clear();
  return KWS_OK;
}

int kws_set_precision(kws_t* h, int precision) {
  if (!h) return KWS_EINVAL;
  if (precision != KWS_PREC_FP32 && precision != KWS_PREC_TC) return fail(h, KWS_EINVAL, "unknown precision");
  h->precision = precision;
  return KWS_OK;
}

int kws_set_fusion(kws_t* h, int on) {
  if (!h) return KWS_EINVAL;
  h->fuse_conv1_block1 = on != 0;
  return KWS_OK;
}

int kws_set_noise_bank(kws_t* h, const float* bank, const int64_t* file_offsets_h, int n_files) {
  if (!h) return KWS_EINVAL;
  if (n_files < 0 || (n_files > 0 && (!bank || !file_offsets_h))) return fail(h, KWS_EINVAL, "bad noise bank");
  if (reinterpret_cast<uintptr_t>(bank) % 16) return fail(h, KWS_EINVAL, "noise bank must be 16-byte aligned");
  KWS_CUDA(h, cudaSetDevice(h->device));
  if (h->file_offsets_d) { cudaFree(h->file_offsets_d); h->file_offsets_d = nullptr; }
  h->bank = bank; h->n_files = n_files; h->bank_len = 0;
  if (n_files == 0) return KWS_OK;
  for (int i = 0; i < n_files; ++i)
    if (file_offsets_h[i + 1] - file_offsets_h[i] <= KWS_SAMPLES)
      return fail(h, KWS_EINVAL, "every background file must be longer than one clip (input_data.py:485-486)");
  h->bank_len = file_offsets_h[n_files];
  KWS_CUDA(h, cudaMalloc(&h->file_offsets_d, (n_files + 1) * sizeof(int64_t)));
  KWS_CUDA(h, cudaMemcpy(h->file_offsets_d, file_offsets_h, (n_files + 1) * sizeof(int64_t), cudaMemcpyHostToDevice));
  return KWS_OK;
}

static int augment_common(kws_t* h, const float* wav, const int16_t* pcm, float divisor, const int32_t* shift,
                          const int32_t* bg_file, const int32_t* bg_off, const float* bg_vol,
                          const float* fg_vol, float* out, int B, int clamp, void* stream) {
  if (!h) return KWS_EINVAL;
  if (B < 0) return fail(h, KWS_EINVAL, "negative batch");
  if (B == 0) return KWS_OK;
  if ((!wav && !pcm) || !shift || !bg_file || !bg_off || !bg_vol || !fg_vol || !out)
    return fail(h, KWS_EINVAL, "null pointer");
  const void* in = wav ? static_cast<const void*>(wav) : static_cast<const void*>(pcm);
  if (reinterpret_cast<uintptr_t>(in) % 16 || reinterpret_cast<uintptr_t>(out) % 16)
    return fail(h, KWS_EINVAL, "waveform buffers must be 16-byte aligned");
  return launch_augment(h, wav, pcm, divisor, shift, bg_file, bg_off, bg_vol, fg_vol, out, B, clamp,
                        static_cast<cudaStream_t>(stream));
}

int kws_augment(kws_t* h, const float* wav, const int32_t* shift, const int32_t* bg_file,
                const int32_t* bg_off, const float* bg_vol, const float* fg_vol, float* out, int B,
                int clamp, void* stream) {
  return augment_common(h, wav, nullptr, 1.0f, shift, bg_file, bg_off, bg_vol, fg_vol, out, B, clamp, stream);
}

int kws_augment_pcm16(kws_t* h, const int16_t* pcm, float divisor, const int32_t* shift,
                      const int32_t* bg_file, const int32_t* bg_off, const float* bg_vol,
                      const float* fg_vol, float* out, int B, int clamp, void* stream) {
  if (h && !(divisor > 0.0f)) return fail(h, KWS_EINVAL, "divisor must be positive");
  return augment_common(h, nullptr, pcm, divisor, shift, bg_file, bg_off, bg_vol, fg_vol, out, B, clamp, stream);
}

int kws_time_stretch_pcm16(kws_t* h, const int16_t* pcm, int B, double rate, int16_t* out, void* stream) {
  if (!h) return KWS_EINVAL;
  if (B < 0) return fail(h, KWS_EINVAL, "negative batch");
  if (B == 0) return KWS_OK;
  if (!pcm || !out) return fail(h, KWS_EINVAL, "null pointer");
  return launch_time_stretch(h, pcm, B, rate, 32767.0f, out, static_cast<cudaStream_t>(stream));
}

int kws_time_stretch_host_pcm16(kws_t* h, const int16_t* pcm_h, int B, double rate, int16_t* out_h) {
  if (!h) return KWS_EINVAL;
  if (B < 0) return fail(h, KWS_EINVAL, "negative batch");
  if (B == 0) return KWS_OK;
  if (!pcm_h || !out_h) return fail(h, KWS_EINVAL, "null pointer");
  KWS_CUDA(h, cudaSetDevice(h->device));
  const int chunk = std::min(B, 8192);
  const size_t clip_bytes = static_cast<size_t>(L) * sizeof(int16_t);
  int rc = ensure_bytes(h, &h->stretch_stage, &h->stretch_stage_bytes, 2 * chunk * clip_bytes);
  if (rc) return rc;
  int16_t* d_in = static_cast<int16_t*>(h->stretch_stage);
  int16_t* d_out = d_in + static_cast<size_t>(chunk) * L;
  cudaStream_t st = h->own_stream;
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int nb = std::min(chunk, B - b0);
    KWS_CUDA(h, cudaMemcpyAsync(d_in, pcm_h + static_cast<size_t>(b0) * L, nb * clip_bytes, cudaMemcpyHostToDevice, st));
    rc = launch_time_stretch(h, d_in, nb, rate, 32767.0f, d_out, st);
    if (rc) { cudaStreamSynchronize(st); return rc; }
    KWS_CUDA(h, cudaMemcpyAsync(out_h + static_cast<size_t>(b0) * L, d_out, nb * clip_bytes, cudaMemcpyDeviceToHost, st));
    KWS_CUDA(h, cudaStreamSynchronize(st));
  }
  return KWS_OK;
}

int kws_frontend_config(kws_t* h, int win, int hop, int n_mel, int n_keep, float f_lo, float f_hi, int sample_rate) {
  if (!h) return KWS_EINVAL;
  KWS_CUDA(h, cudaSetDevice(h->device));
  return frontend_build(h, win, hop, n_mel, n_keep, f_lo, f_hi, sample_rate);
}

int kws_frontend_config_contrib(kws_t* h, int window_size, int stride, int sample_rate, float lower_hz,
                                float upper_hz, int filterbank_channels, int dct_coefficient_count) {
  if (!h) return KWS_EINVAL;
  KWS_CUDA(h, cudaSetDevice(h->device));
  return frontend_build(h, window_size, stride, filterbank_channels, dct_coefficient_count, lower_hz, upper_hz,
                        sample_rate, /*flavour=*/1);
}

int kws_frontend_frames(const kws_t* h) { return (h && h->fe.configured) ? h->fe.frames : 0; }

int kws_features(kws_t* h, const float* wav, int B, int kind, float* out, void* stream) {
  if (!h) return KWS_EINVAL;
  const int rc = features_dispatch(h, wav, B, kind, out, static_cast<cudaStream_t>(stream));
  return rc ? rc : (B > 0 ? mark_user_stream(h, static_cast<cudaStream_t>(stream)) : KWS_OK);
}

int kws_model_load(kws_t* h, int slot, int arch, const kws_tensor_h* tensors_h, int n) {
  if (!h) return KWS_EINVAL;
  if (!tensors_h || n <= 0) return fail(h, KWS_EINVAL, "no tensors");
  KWS_CUDA(h, cudaSetDevice(h->device));
  return model_build(h, slot, arch, tensors_h, n);
}

int kws_model_classes(const kws_t* h, int slot) {
  if (!h || slot < 0 || slot >= KWS_MAX_MODELS || !h->models[slot].loaded) return 0;
  return h->models[slot].classes;
}

int kws_forward(kws_t* h, int slot, const float* wav, int B, const int32_t* view_shift_h,
                const float* view_gain_h, int n_views, float* probs_mean, int32_t* argmax, void* stream) {
  if (!h) return KWS_EINVAL;
  ViewTable vt;
  int rc = make_views(h, view_shift_h, view_gain_h, n_views, &vt);
  if (rc) return rc;
  rc = forward_dispatch(h, slot, wav, B, vt, probs_mean, argmax, static_cast<cudaStream_t>(stream));
  return rc ? rc : (B > 0 ? mark_user_stream(h, static_cast<cudaStream_t>(stream)) : KWS_OK);
}

int kws_debug_activation(kws_t* h, int slot, const float* wav, int B, const int32_t* view_shift_h,
                         const float* view_gain_h, int n_views, int layer, float* out, void* stream) {
  if (!h) return KWS_EINVAL;
  ViewTable vt;
  int rc = make_views(h, view_shift_h, view_gain_h, n_views, &vt);
  if (rc) return rc;
  if (slot < 0 || slot >= KWS_MAX_MODELS || !h->models[slot].loaded) return fail(h, KWS_ESTATE, "model not loaded");
  if (h->models[slot].arch == KWS_ARCH_STEFFENET) return fail(h, KWS_EUNSUPPORTED, "kws_debug_activation is not available for steffeNet");
  if (layer < 0 || layer > h->models[slot].n_blocks || !out || !wav || B <= 0) return fail(h, KWS_EINVAL, "bad arguments");
  if (B * n_views > h->max_rows) return fail(h, KWS_EINVAL, "debug activation needs B*n_views <= max_rows");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  rc = h->precision == KWS_PREC_FP32
           ? launch_forward_f32(h, h->models[slot], wav, B, vt, nullptr, nullptr, st, layer, out)
           : launch_forward_tc(h, h->models[slot], wav, B, vt, nullptr, nullptr, st, layer, out);
  return rc ? rc : mark_user_stream(h, st);
}

int kws_convert_classes(kws_t* h, const float* probs, int B, int C_in, const int32_t* class_map_h,
                        int C_out, float* probs_out, uint8_t* probs_u8, void* stream) {
  if (!h) return KWS_EINVAL;
  if (B < 0 || !class_map_h || (B > 0 && !probs)) return fail(h, KWS_EINVAL, "bad arguments");
  return launch_convert(h, probs, B, C_in, class_map_h, C_out, probs_out, probs_u8, static_cast<cudaStream_t>(stream));
}

int kws_select(kws_t* h, const uint8_t* probs_u8, int B, int C, double thresh, int32_t* label,
               uint8_t* keep, void* stream) {
  if (!h) return KWS_EINVAL;
  if (B < 0 || (B > 0 && !probs_u8)) return fail(h, KWS_EINVAL, "bad arguments");
  return launch_select(h, probs_u8, B, C, thresh, label, keep, static_cast<cudaStream_t>(stream));
}

int kws_vote(kws_t* h, const int32_t* labels, int M, int B, int min_count, int32_t* voted,
             uint8_t* clear, void* stream) {
  if (!h) return KWS_EINVAL;
  if (B < 0 || (B > 0 && !labels)) return fail(h, KWS_EINVAL, "bad arguments");
  return launch_vote(h, labels, M, B, min_count, voted, clear, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
