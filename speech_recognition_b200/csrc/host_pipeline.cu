// Host-buffer entry points of libkws.so (include/kws.h): what the reference-side binding calls instead of
// sess.run(...) per clip (input_data.py:517-531) and model.predict(...) per batch (make_submission.py:120-146).
//
// One staging allocation on the device holds two slots of waveforms, parameters and results; H2D copies,
// kernels and D2H copies of consecutive chunks run on three streams ordered by per-slot events, so the copies
// hide under the kernels.  Waveforms arrive as fp32 [B,16000] or -- the wire format of the reference's WAV
// files (input_data.py:334-336) -- as 16-bit PCM, which halves the bytes that cross PCIe; the decode is fused
// into the augment kernel's load.
//
// Caller buffers may be pinned (cudaHostAlloc / cudaHostRegister: the DMA engine reads them directly) or
// pageable (a plain np.ndarray).  cudaMemcpyAsync on pageable memory degrades to a synchronous staged copy and
// would serialise the three streams, so pageable buffers go through the handle's own pinned slots instead: a
// small pool of host threads copies chunk k+1 into a pinned slot while the GPU works on chunk k, and results
// are drained from pinned slots the same way.  Which path a buffer takes is detected per call
// (cudaPointerGetAttributes) unless kws_set_host_staging forces it.
#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

#include "common.cuh"

namespace kws {

// ------------------------------------------------------------------------------------------
// pageable -> pinned copies on a few host threads (one memcpy stream per thread saturates ~10 GB/s of
// the ~30 GB/s a PCIe gen5 x16 link moves)
// ------------------------------------------------------------------------------------------
class CopyPool {
 public:
  explicit CopyPool(int workers) {
    jobs_.resize(workers);
    for (int i = 0; i < workers; ++i) th_.emplace_back([this, i] { run(i); });
  }
  ~CopyPool() {
    { std::lock_guard<std::mutex> g(m_); stop_ = true; }
    cv_work_.notify_all();
    for (auto& t : th_) t.join();
  }
  void copy(void* dst, const void* src, size_t bytes) {
    const size_t parts = th_.size() + 1;
    if (th_.empty() || bytes < (1u << 20)) { memcpy(dst, src, bytes); return; }
    const size_t piece = ((bytes + parts - 1) / parts + 4095) & ~static_cast<size_t>(4095);
    {
      std::lock_guard<std::mutex> g(m_);
      for (size_t i = 0; i < th_.size(); ++i) {
        const size_t o = std::min(bytes, (i + 1) * piece), e = std::min(bytes, (i + 2) * piece);
        jobs_[i] = Job{static_cast<char*>(dst) + o, static_cast<const char*>(src) + o, e - o};
      }
      pending_ = static_cast<int>(th_.size());
      ++gen_;
    }
    cv_work_.notify_all();
    memcpy(dst, src, std::min(bytes, piece));
    std::unique_lock<std::mutex> g(m_);
    cv_done_.wait(g, [this] { return pending_ == 0; });
  }

 private:
  struct Job { char* d; const char* s; size_t n; };
  void run(int idx) {
    uint64_t seen = 0;
    std::unique_lock<std::mutex> g(m_);
    for (;;) {
      cv_work_.wait(g, [&] { return stop_ || gen_ != seen; });
      if (stop_) return;
      seen = gen_;
      const Job j = jobs_[idx];
      g.unlock();
      if (j.n) memcpy(j.d, j.s, j.n);
      g.lock();
      if (--pending_ == 0) cv_done_.notify_one();
    }
  }
  std::vector<std::thread> th_;
  std::vector<Job> jobs_;
  std::mutex m_;
  std::condition_variable cv_work_, cv_done_;
  int pending_ = 0;
  uint64_t gen_ = 0;
  bool stop_ = false;
};

void copy_pool_destroy(CopyPool* p) { delete p; }

namespace {

bool is_pinned(const void* p) {
  if (!p) return true;
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

struct StageLayout {
  size_t wav, aug, shift, bgf, bgo, bgv, fgv, feat, probs, amax, total;
};

StageLayout stage_layout(int nb, int classes, size_t out_dim) {
  StageLayout s{};
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 255) / 256 * 256; return r; };
  s.wav = take(static_cast<size_t>(nb) * L * 4);            // fp32 waveforms, or int16 PCM in the first half
  s.aug = take(static_cast<size_t>(nb) * L * 4);
  s.shift = take(nb * 4); s.bgf = take(nb * 4); s.bgo = take(nb * 4); s.bgv = take(nb * 4); s.fgv = take(nb * 4);
  s.feat = take(static_cast<size_t>(nb) * out_dim * 4);
  s.probs = take(static_cast<size_t>(nb) * classes * 4);
  s.amax = take(nb * 4);
  s.total = o;
  return s;
}

struct PipelineArgs {
  int slot;
  const float* wav_h; const int16_t* pcm_h; float divisor;
  const int32_t* shift_h; const int32_t* bg_file_h; const int32_t* bg_off_h; const float* bg_vol_h; const float* fg_vol_h;
  int B, feat_kind;
  const int32_t* view_shift_h; const float* view_gain_h; int n_views;
  float* feat_h; float* probs_h; int32_t* argmax_h;
};

int pipeline(kws_handle* h, const PipelineArgs& a) {
  if (!h) return KWS_EINVAL;
  const int B = a.B;
  if (B < 0) return fail(h, KWS_EINVAL, "negative batch");
  if (B == 0) return KWS_OK;
  const bool pcm = a.pcm_h != nullptr;
  if (!a.wav_h && !pcm) return fail(h, KWS_EINVAL, "null waveform pointer");
  if (pcm && !(a.divisor > 0.0f)) return fail(h, KWS_EINVAL, "divisor must be positive");
  KWS_CUDA(h, cudaSetDevice(h->device));
  const bool do_aug = a.shift_h || a.bg_file_h || a.bg_off_h || a.bg_vol_h || a.fg_vol_h;
  if (do_aug && !(a.shift_h && a.bg_file_h && a.bg_off_h && a.bg_vol_h && a.fg_vol_h))
    return fail(h, KWS_EINVAL, "augmentation parameters must be all present or all NULL");
  const bool do_feat = a.feat_kind >= 0;
  const bool do_fwd = a.n_views > 0;
  const bool raw_out = !do_feat && !do_fwd && a.feat_h;        // 'raw' representation: the (augmented) waveform itself
  if (do_feat && !h->fe.configured) return fail(h, KWS_ESTATE, "front end not configured");
  ViewTable vt{};
  int classes = 0;
  if (do_fwd) {
    int rc = make_views(h, a.view_shift_h, a.view_gain_h, a.n_views, &vt);
    if (rc) return rc;
    if (a.slot < 0 || a.slot >= KWS_MAX_MODELS || !h->models[a.slot].loaded) return fail(h, KWS_ESTATE, "model not loaded");
    classes = h->models[a.slot].classes;
  }
  const size_t fdim = do_feat ? feat_dim(h, a.feat_kind) : (raw_out ? static_cast<size_t>(L) : 0);
  const size_t esize = pcm ? sizeof(int16_t) : sizeof(float);
  const void* in_h = pcm ? static_cast<const void*>(a.pcm_h) : static_cast<const void*>(a.wav_h);

  // Chunks of one forward pass worth of clip-views (max_rows), so that with two staging slots
  // the H2D copy of chunk k+1 (copy stream) and the D2H copy of chunk k-1 (second copy stream)
  // run under the kernels of chunk k (compute stream); events order the three streams per slot.
  // Clips per staged chunk (KWS_HOST_CHUNK overrides).  r02 A/B on one 8-GPU box, 16,384 clips x 8 views per call: 8,192-clip
  // chunks are 2 % faster than 4,096-clip ones on one or two GPUs (larger launches) but slower from four GPUs on, where
  // the ranks share the host's memory bandwidth and the exposed first upload / last download of a call grows with the
  // chunk (3.65 M against 3.70 M clips/s on eight); copy-bound calls (3 views or fewer) lose 12 % to the larger chunk.
  static const int chunk_env = [] { const char* e = getenv("KWS_HOST_CHUNK"); return e && atoi(e) > 0 ? atoi(e) : 0; }();
  const int chunk_cap = chunk_env ? chunk_env : 4096;
  const int chunk = std::max(1, std::min(std::min(B, chunk_cap), do_fwd ? std::max(1, h->max_rows / a.n_views) : 2048));
  const StageLayout lay = stage_layout(chunk, std::max(classes, 1), fdim);
  int rc = ensure_bytes(h, &h->stage_d, &h->stage_bytes, 2 * lay.total);
  if (rc) return rc;
  if (!h->h2d_stream) {
    KWS_CUDA(h, cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking));
    KWS_CUDA(h, cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      KWS_CUDA(h, cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming));
      KWS_CUDA(h, cudaEventCreateWithFlags(&h->ev_comp[i], cudaEventDisableTiming));
      KWS_CUDA(h, cudaEventCreateWithFlags(&h->ev_d2h[i], cudaEventDisableTiming));
    }
  }
  cudaStream_t st = h->own_stream, s_in = h->h2d_stream, s_out = h->d2h_stream;
  // the handle's workspace may still be in use by a device entry point on the caller's stream
  if (h->ev_user) KWS_CUDA(h, cudaStreamWaitEvent(st, h->ev_user, 0));

  // ---- pageable caller buffers go through pinned slots ----
  const bool want_out = a.feat_h || a.probs_h || a.argmax_h;
  bool stage_in, stage_out;
  if (h->staging_mode == KWS_STAGING_ALWAYS) { stage_in = true; stage_out = want_out; }
  else if (h->staging_mode == KWS_STAGING_NEVER) { stage_in = stage_out = false; }
  else {
    stage_in = !is_pinned(in_h);
    stage_out = want_out && !(is_pinned(a.feat_h) && is_pinned(a.probs_h) && is_pinned(a.argmax_h));
  }
  const size_t par_bytes = do_aug ? 5 * static_cast<size_t>(chunk) * 4 : 0;
  const size_t pin_in_need = static_cast<size_t>(chunk) * L * esize + par_bytes;
  const size_t out_feat = (a.feat_h ? static_cast<size_t>(chunk) * fdim * 4 : 0);
  const size_t out_probs = (do_fwd && a.probs_h ? static_cast<size_t>(chunk) * classes * 4 : 0);
  const size_t pin_out_need = out_feat + out_probs + static_cast<size_t>(chunk) * 4;
  if (stage_in || stage_out) {
    if (!h->copy_pool) {
      static const int workers = [] { const char* e = getenv("KWS_STAGE_THREADS"); return e ? std::max(0, atoi(e) - 1) : 3; }();
      h->copy_pool = new CopyPool(workers);
    }
    for (int i = 0; i < 2; ++i) {
      if (stage_in && (rc = ensure_bytes(h, &h->pin_in[i], &h->pin_in_bytes[i], pin_in_need, true))) return rc;
      if (stage_out && (rc = ensure_bytes(h, &h->pin_out[i], &h->pin_out_bytes[i], pin_out_need, true))) return rc;
    }
  }
  struct Pending { int b0, nb; bool live; } pend[2] = {{0, 0, false}, {0, 0, false}};
  auto drain = [&](int s) -> int {                     // pinned result slot s -> caller buffers (host blocks on its D2H)
    if (!pend[s].live) return KWS_OK;
    KWS_CUDA(h, cudaEventSynchronize(h->ev_d2h[s]));
    const char* src = static_cast<const char*>(h->pin_out[s]);
    const int b0 = pend[s].b0, nb = pend[s].nb;
    if (a.feat_h) h->copy_pool->copy(a.feat_h + static_cast<size_t>(b0) * fdim, src, static_cast<size_t>(nb) * fdim * 4);
    if (do_fwd && a.probs_h) memcpy(a.probs_h + static_cast<size_t>(b0) * classes, src + out_feat, static_cast<size_t>(nb) * classes * 4);
    if (do_fwd && a.argmax_h) memcpy(a.argmax_h + b0, src + out_feat + out_probs, static_cast<size_t>(nb) * 4);
    pend[s].live = false;
    return KWS_OK;
  };
  // an error inside the loop must not leave copies into / out of caller memory in flight
  auto bail = [&](int code) { cudaStreamSynchronize(s_in); cudaStreamSynchronize(st); cudaStreamSynchronize(s_out); return code; };
#define KWS_PIPE(expr)                                                                                \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess) return bail(fail(h, KWS_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e))); \
  } while (0)

  int k = 0;
  // Chunk schedule: the first H2D copy and the last D2H copy are the only ones that cannot hide under kernels,
  // so the call starts with a short chunk, doubles it (a chunk's H2D copy is ~1.8x faster than the kernels of the
  // chunk before it, so it stays almost hidden) up to the forward's chunk size -- large chunks run the network
  // ~8 % faster than small ones -- and ends with a short tail.  (Measured on schedules 1/4..1/16, x1.5..x2.5:
  // all within 380-400k clips/s at 4096 clips per call; this one was the best.)
  const int first = std::max(std::min(chunk, 64), std::min(chunk / 16, B / 8));
  const int tail = std::max(std::min(chunk, 64), std::min(std::min(chunk / 4, 512), B / 8));
  const bool ramp = B > 2 * first + tail;
  int next = ramp ? first : chunk;
  for (int b0 = 0, nb = 0; b0 < B; b0 += nb, ++k) {
    const int rem = B - b0;
    nb = std::min(next, rem);
    if (ramp && rem > tail && rem <= next + tail) nb = rem - tail;          // leave a short tail
    next = std::min(chunk, 2 * next);
    const int sk = k & 1;
    char* base = static_cast<char*>(h->stage_d) + sk * lay.total;
    float* d_wav = reinterpret_cast<float*>(base + lay.wav);
    float* d_aug = reinterpret_cast<float*>(base + lay.aug);
    int32_t* d_par[5] = {reinterpret_cast<int32_t*>(base + lay.shift), reinterpret_cast<int32_t*>(base + lay.bgf),
                         reinterpret_cast<int32_t*>(base + lay.bgo), reinterpret_cast<int32_t*>(base + lay.bgv),
                         reinterpret_cast<int32_t*>(base + lay.fgv)};
    const void* par_h[5] = {a.shift_h, a.bg_file_h, a.bg_off_h, a.bg_vol_h, a.fg_vol_h};
    // ---- host: results of chunk k-2 leave their pinned slot, inputs of chunk k enter theirs ----
    if (stage_out && (rc = drain(sk))) return bail(rc);
    const char* src = static_cast<const char*>(in_h) + static_cast<size_t>(b0) * L * esize;
    const size_t in_bytes = static_cast<size_t>(nb) * L * esize;
    const char* par_src[5];
    for (int i = 0; i < 5; ++i) par_src[i] = do_aug ? static_cast<const char*>(par_h[i]) + static_cast<size_t>(b0) * 4 : nullptr;
    if (stage_in) {
      if (k >= 2) KWS_PIPE(cudaEventSynchronize(h->ev_h2d[sk]));            // the H2D copy of chunk k-2 has read the slot
      char* pin = static_cast<char*>(h->pin_in[sk]);
      h->copy_pool->copy(pin, src, in_bytes);
      src = pin;
      for (int i = 0; i < 5 && do_aug; ++i) {
        char* dst = pin + static_cast<size_t>(chunk) * L * esize + static_cast<size_t>(i) * chunk * 4;
        memcpy(dst, par_src[i], static_cast<size_t>(nb) * 4);
        par_src[i] = dst;
      }
    }
    // ---- copy stream: inputs of chunk k (the slot was last read by the kernels and -- 'raw' output -- the D2H copy of chunk k-2) ----
    if (k >= 2) {
      KWS_PIPE(cudaStreamWaitEvent(s_in, h->ev_comp[sk], 0));
      KWS_PIPE(cudaStreamWaitEvent(s_in, h->ev_d2h[sk], 0));
    }
    KWS_PIPE(cudaMemcpyAsync(d_wav, src, in_bytes, cudaMemcpyHostToDevice, s_in));
    if (do_aug)
      for (int i = 0; i < 5; ++i) KWS_PIPE(cudaMemcpyAsync(d_par[i], par_src[i], static_cast<size_t>(nb) * 4, cudaMemcpyHostToDevice, s_in));
    KWS_PIPE(cudaEventRecord(h->ev_h2d[sk], s_in));
    // ---- compute stream: kernels of chunk k (its result buffers were last read by D2H of chunk k-2) ----
    KWS_PIPE(cudaStreamWaitEvent(st, h->ev_h2d[sk], 0));
    if (k >= 2) KWS_PIPE(cudaStreamWaitEvent(st, h->ev_d2h[sk], 0));
    const float* x = d_wav;
    if (do_aug || pcm) {                                  // PCM without augmentation: identity parameters (decode only)
      rc = launch_augment(h, pcm ? nullptr : d_wav, pcm ? reinterpret_cast<const int16_t*>(d_wav) : nullptr, pcm ? a.divisor : 1.0f,
                          do_aug ? d_par[0] : nullptr, do_aug ? d_par[1] : nullptr, do_aug ? d_par[2] : nullptr,
                          do_aug ? reinterpret_cast<const float*>(d_par[3]) : nullptr,
                          do_aug ? reinterpret_cast<const float*>(d_par[4]) : nullptr, d_aug, nb, 0, st);
      if (rc) return bail(rc);
      x = d_aug;
    }
    float* d_feat = reinterpret_cast<float*>(base + lay.feat);
    float* d_probs = reinterpret_cast<float*>(base + lay.probs);
    int32_t* d_amax = reinterpret_cast<int32_t*>(base + lay.amax);
    if (do_feat && (rc = features_dispatch(h, x, nb, a.feat_kind, d_feat, st))) return bail(rc);
    if (do_fwd && (rc = forward_dispatch(h, a.slot, x, nb, vt, d_probs, d_amax, st))) return bail(rc);
    KWS_PIPE(cudaEventRecord(h->ev_comp[sk], st));
    // ---- second copy stream: results of chunk k ----
    KWS_PIPE(cudaStreamWaitEvent(s_out, h->ev_comp[sk], 0));
    char* pin_o = stage_out ? static_cast<char*>(h->pin_out[sk]) : nullptr;
    if (a.feat_h && (do_feat || raw_out)) {
      void* dst = stage_out ? static_cast<void*>(pin_o) : static_cast<void*>(a.feat_h + static_cast<size_t>(b0) * fdim);
      KWS_PIPE(cudaMemcpyAsync(dst, do_feat ? d_feat : x, static_cast<size_t>(nb) * fdim * 4, cudaMemcpyDeviceToHost, s_out));
    }
    if (do_fwd && a.probs_h) {
      void* dst = stage_out ? static_cast<void*>(pin_o + out_feat) : static_cast<void*>(a.probs_h + static_cast<size_t>(b0) * classes);
      KWS_PIPE(cudaMemcpyAsync(dst, d_probs, static_cast<size_t>(nb) * classes * 4, cudaMemcpyDeviceToHost, s_out));
    }
    if (do_fwd && a.argmax_h) {
      void* dst = stage_out ? static_cast<void*>(pin_o + out_feat + out_probs) : static_cast<void*>(a.argmax_h + b0);
      KWS_PIPE(cudaMemcpyAsync(dst, d_amax, static_cast<size_t>(nb) * 4, cudaMemcpyDeviceToHost, s_out));
    }
    KWS_PIPE(cudaEventRecord(h->ev_d2h[sk], s_out));
    if (stage_out) pend[sk] = Pending{b0, nb, true};
  }
  if (stage_out) {
    if ((rc = drain(k & 1))) return bail(rc);             // older chunk first
    if ((rc = drain((k + 1) & 1))) return bail(rc);
  }
  KWS_PIPE(cudaStreamSynchronize(s_out));
  KWS_PIPE(cudaStreamSynchronize(st));
#undef KWS_PIPE
  return KWS_OK;
}

}  // namespace
}  // namespace kws

using namespace kws;

extern "C" {

int kws_set_host_staging(kws_t* h, int mode) {
  if (!h) return KWS_EINVAL;
  if (mode != KWS_STAGING_AUTO && mode != KWS_STAGING_ALWAYS && mode != KWS_STAGING_NEVER)
    return fail(h, KWS_EINVAL, "unknown staging mode");
  h->staging_mode = mode;
  return KWS_OK;
}

int kws_pipeline_host(kws_t* h, int slot, const float* wav_h, const int32_t* shift_h,
                      const int32_t* bg_file_h, const int32_t* bg_off_h, const float* bg_vol_h,
                      const float* fg_vol_h, int B, int feat_kind, const int32_t* view_shift_h,
                      const float* view_gain_h, int n_views, float* feat_h, float* probs_h,
                      int32_t* argmax_h) {
  return pipeline(h, PipelineArgs{slot, wav_h, nullptr, 1.0f, shift_h, bg_file_h, bg_off_h, bg_vol_h, fg_vol_h, B, feat_kind,
                                  view_shift_h, view_gain_h, n_views, feat_h, probs_h, argmax_h});
}

int kws_pipeline_host_pcm16(kws_t* h, int slot, const int16_t* pcm_h, float divisor, const int32_t* shift_h,
                            const int32_t* bg_file_h, const int32_t* bg_off_h, const float* bg_vol_h,
                            const float* fg_vol_h, int B, int feat_kind, const int32_t* view_shift_h,
                            const float* view_gain_h, int n_views, float* feat_h, float* probs_h,
                            int32_t* argmax_h) {
  if (h && B > 0 && !pcm_h) return fail(h, KWS_EINVAL, "null PCM pointer");
  return pipeline(h, PipelineArgs{slot, nullptr, pcm_h, divisor, shift_h, bg_file_h, bg_off_h, bg_vol_h, fg_vol_h, B, feat_kind,
                                  view_shift_h, view_gain_h, n_views, feat_h, probs_h, argmax_h});
}

int kws_predict_host(kws_t* h, int slot, const float* wav_h, int B, const int32_t* view_shift_h,
                     const float* view_gain_h, int n_views, float* probs_h, int32_t* argmax_h) {
  if (h && n_views <= 0) return fail(h, KWS_EINVAL, "n_views must be positive");
  return kws_pipeline_host(h, slot, wav_h, nullptr, nullptr, nullptr, nullptr, nullptr, B, -1, view_shift_h,
                           view_gain_h, n_views, nullptr, probs_h, argmax_h);
}

int kws_predict_host_pcm16(kws_t* h, int slot, const int16_t* pcm_h, float divisor, int B, const int32_t* view_shift_h,
                           const float* view_gain_h, int n_views, float* probs_h, int32_t* argmax_h) {
  if (h && n_views <= 0) return fail(h, KWS_EINVAL, "n_views must be positive");
  return kws_pipeline_host_pcm16(h, slot, pcm_h, divisor, nullptr, nullptr, nullptr, nullptr, nullptr, B, -1,
                                 view_shift_h, view_gain_h, n_views, nullptr, probs_h, argmax_h);
}

int kws_get_data_host(kws_t* h, const float* wav_h, const int32_t* shift_h, const int32_t* bg_file_h,
                      const int32_t* bg_off_h, const float* bg_vol_h, const float* fg_vol_h, int B,
                      int clamp, int kind, float* out_h) {
  if (h && clamp) return fail(h, KWS_EUNSUPPORTED, "clamp is only available through kws_augment");
  if (h && !out_h) return fail(h, KWS_EINVAL, "null output pointer");
  return kws_pipeline_host(h, 0, wav_h, shift_h, bg_file_h, bg_off_h, bg_vol_h, fg_vol_h, B, kind, nullptr,
                           nullptr, 0, out_h, nullptr, nullptr);
}

}  // extern "C"
