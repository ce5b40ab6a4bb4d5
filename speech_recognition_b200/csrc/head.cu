// K7 head -- attention pooling + classifier + TTA mean + argmax, one warp per clip-view.
// Replaces, per view (reference model.py:819-830):
//   Flatten -> Dense(T, softmax) -> x * a[:, None] -> GlobalMaxPool1D || GlobalAveragePooling1D
//   -> Dense(classes, softmax)            (exp 195/206)
//   Flatten -> Dense(T, softmax, no bias) -> mean_t(x * a) -> Dense(32, softmax)   (exp 106)
// and across views (make_submission.py:137-146): probs = (p_0 + p_1 + ...) / n_views in view
// order, argmax with first-index tie rule.  Memory/latency-bound CUDA-core work with
// warp-shuffle reductions; the weights (166 KB + 48 KB) are shared-memory resident.
#include <algorithm>

#include "common.cuh"
#include "gemm_f32.cuh"

namespace kws {

namespace {

constexpr int HEAD_T = 9;                 // time steps entering the head (both shipped archs)
constexpr int HEAD_MAX_CLASSES = 32;

__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(__half v) { return __half2float(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Persistent kernel: the dense_1 (166 KB) and dense_2 (48 KB, transposed) weights live in shared
// memory for the whole launch -- re-fetching them per clip-view from L2 is what bounds a
// one-CTA-per-clip formulation -- and every warp owns IPW consecutive views of one clip at a time and reuses
// every weight it reads for all of them.  A CTA iteration covers floor(16 * IPW / n_views) clips;
// the per-view probabilities meet in shared memory and are averaged in view order.
//
// r02: a lane owns 8 CONSECUTIVE channels (one 16-byte load of the fp16 activation) instead of every 32nd element
// (a 2-byte load: r01 spent the kernel waiting for 576 of them per lane and 4 views, 0.11 of the HBM roofline).
// The flattened activation index is i = 256 k + 8 lane + e; dense_1's rows are re-laid out in shared memory as
// [k][e][j 0..3 | j 4..7 | j 8][lane] so that the nine weights of an element are two conflict-free LDS.128 and one
// LDS.32.  The next block's activations are in flight (as raw 16-byte words) while the current block is multiplied.
constexpr int HEAD_WARPS = 16;
constexpr int HEAD_BLK = 256;                       // activation elements per block: 32 lanes x 8
constexpr int HEAD_WROW = 2 * 128 + 32;             // floats of one (k, e) weight row group: j0-3 | j4-7 | j8, each x 32 lanes
constexpr int HEAD_MAX_KB = 2;                      // channel blocks of 256 in the pooling: C <= 512

template <typename TAct> struct Raw8;
template <> struct Raw8<__half> { uint4 v; };
template <> struct Raw8<float> { float4 a, b; };
__device__ __forceinline__ void load_raw8(const __half* p, Raw8<__half>& r) { r.v = __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void load_raw8(const float* p, Raw8<float>& r) {
  r.a = __ldg(reinterpret_cast<const float4*>(p)); r.b = __ldg(reinterpret_cast<const float4*>(p) + 1);
}
__device__ __forceinline__ void unpack8(const Raw8<__half>& r, float (&x)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&r.v);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); x[2 * i] = f.x; x[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void unpack8(const Raw8<float>& r, float (&x)[8]) {
  x[0] = r.a.x; x[1] = r.a.y; x[2] = r.a.z; x[3] = r.a.w; x[4] = r.b.x; x[5] = r.b.y; x[6] = r.b.z; x[7] = r.b.w;
}

template <typename TAct, int IPW>
__global__ void __launch_bounds__(HEAD_WARPS * 32, 1)
head_kernel(const TAct* __restrict__ act, int C, int n_views, int n_clips, const float* __restrict__ w_d1,
            const float* __restrict__ b_d1, const float* __restrict__ w_d2, int classes,
            int pool_max_avg, float* __restrict__ probs_mean, int32_t* __restrict__ argmax) {
  extern __shared__ float sm[];
  const int n = HEAD_T * C;
  const int nblk = (n + HEAD_BLK - 1) / HEAD_BLK;
  const int feat = pool_max_avg ? 2 * C : C;
  float* w1 = sm;                                       // [nblk][8][HEAD_WROW]
  float* w2t = w1 + nblk * 8 * HEAD_WROW;               // [classes][feat]
  float* pv = w2t + classes * feat;                     // [clips per iteration][n_views][32] per-view probabilities

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // eight loads in flight per thread: with one dependent L2 round trip per element this prologue was 13 % of the
  // launch (r02q profile)
  constexpr int NT = HEAD_WARPS * 32, PF = 8;
  for (int base = tid; base < n * HEAD_T; base += PF * NT) {               // dense_1 [n][9] -> [k][e][j groups][lane]
    float v[PF];
#pragma unroll
    for (int u = 0; u < PF; ++u) { const int idx = base + u * NT; v[u] = idx < n * HEAD_T ? __ldg(&w_d1[idx]) : 0.0f; }
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int idx = base + u * NT;
      if (idx < n * HEAD_T) {
        const int i = idx / HEAD_T, j = idx - i * HEAD_T;
        const int k = i / HEAD_BLK, r = i - k * HEAD_BLK, ln = r >> 3, e = r & 7;
        float* row = w1 + (k * 8 + e) * HEAD_WROW;
        row[j < 4 ? ln * 4 + j : (j < 8 ? 128 + ln * 4 + (j - 4) : 256 + ln)] = v[u];
      }
    }
  }
  for (int base = tid; base < feat * classes; base += PF * NT) {
    float v[PF];
#pragma unroll
    for (int u = 0; u < PF; ++u) { const int i = base + u * NT; v[u] = i < feat * classes ? __ldg(&w_d2[i]) : 0.0f; }
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int i = base + u * NT;
      if (i < feat * classes) { const int r = i / classes, j = i - r * classes; w2t[j * feat + r] = v[u]; }
    }
  }
  const float bias = lane < HEAD_T ? __ldg(&b_d1[lane]) : 0.0f;
  __syncthreads();

  const int wpc = n_views / IPW;                        // warps per clip (IPW divides n_views)
  const int cpi = HEAD_WARPS / wpc;                     // clips per CTA iteration
  const int ci = warp / wpc, v0 = (warp - ci * wpc) * IPW;
  for (int clip0 = blockIdx.x * cpi; clip0 < n_clips; clip0 += gridDim.x * cpi) {
    const int clip = clip0 + ci;
    if (ci < cpi && clip < n_clips) {
      const TAct* x0 = act + (static_cast<size_t>(clip) * n_views + v0) * n;
      // ---- dense_1 of IPW views against the same weights ----
      float p[IPW][HEAD_T];
#pragma unroll
      for (int u = 0; u < IPW; ++u)
#pragma unroll
        for (int j = 0; j < HEAD_T; ++j) p[u][j] = 0.0f;
      Raw8<TAct> nxt[IPW];
      if (8 * lane < n) {
#pragma unroll
        for (int u = 0; u < IPW; ++u) load_raw8(x0 + static_cast<size_t>(u) * n + 8 * lane, nxt[u]);
      }
      for (int k = 0; k < nblk; ++k) {
        const int i0 = HEAD_BLK * k + 8 * lane;
        float xa[IPW][8];
#pragma unroll
        for (int u = 0; u < IPW; ++u) unpack8(nxt[u], xa[u]);
        if (i0 + HEAD_BLK < n) {                          // the next block's loads fly under this block's FMAs
#pragma unroll
          for (int u = 0; u < IPW; ++u) load_raw8(x0 + static_cast<size_t>(u) * n + i0 + HEAD_BLK, nxt[u]);
        }
        if (i0 < n) {
          const float* wk = w1 + k * 8 * HEAD_WROW + lane * 4;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float4 wa = *reinterpret_cast<const float4*>(wk + e * HEAD_WROW);
            const float4 wb = *reinterpret_cast<const float4*>(wk + e * HEAD_WROW + 128);
            const float w8 = w1[(k * 8 + e) * HEAD_WROW + 256 + lane];
#pragma unroll
            for (int u = 0; u < IPW; ++u) {
              const float xv = xa[u][e];
              p[u][0] = fmaf(xv, wa.x, p[u][0]); p[u][1] = fmaf(xv, wa.y, p[u][1]);
              p[u][2] = fmaf(xv, wa.z, p[u][2]); p[u][3] = fmaf(xv, wa.w, p[u][3]);
              p[u][4] = fmaf(xv, wb.x, p[u][4]); p[u][5] = fmaf(xv, wb.y, p[u][5]);
              p[u][6] = fmaf(xv, wb.z, p[u][6]); p[u][7] = fmaf(xv, wb.w, p[u][7]);
              p[u][8] = fmaf(xv, w8, p[u][8]);
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < IPW; ++u) {
        const TAct* x = x0 + static_cast<size_t>(u) * n;
        // ---- softmax over T ----
        float l = -INFINITY;
#pragma unroll
        for (int j = 0; j < HEAD_T; ++j) {
          const float s = warp_sum(p[u][j]);
          if (lane == j) l = s + bias;
        }
        float mx = warp_max(l);
        float e = lane < HEAD_T ? expf(l - mx) : 0.0f;
        float s = warp_sum(e);
        const float a = __fdiv_rn(e, s);                   // lane t holds att[t]
        float att[HEAD_T];
#pragma unroll
        for (int t = 0; t < HEAD_T; ++t) att[t] = __shfl_sync(0xffffffffu, a, t);
        // ---- multiply_1 + pooling: lane owns channels 256 kb + 8 lane .. + 7 ----
        float z0[HEAD_MAX_KB * 8], z1[HEAD_MAX_KB * 8];
#pragma unroll
        for (int kb = 0; kb < HEAD_MAX_KB; ++kb) {
          const int c0 = HEAD_BLK * kb + 8 * lane;
          float m[8], sum_x[8], sum_w[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) { m[q] = -INFINITY; sum_x[q] = 0.0f; sum_w[q] = 0.0f; }
          if (c0 < C) {
#pragma unroll
            for (int t = 0; t < HEAD_T; ++t) {
              Raw8<TAct> r;
              load_raw8(x + t * C + c0, r);
              float xv[8];
              unpack8(r, xv);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float wv = __fmul_rn(xv[q], att[t]);       // multiply_1
                m[q] = fmaxf(m[q], wv);
                sum_x[q] += xv[q];
                sum_w[q] += wv;
              }
            }
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            z0[kb * 8 + q] = 0.0f; z1[kb * 8 + q] = 0.0f;
            if (c0 < C) {
              if (pool_max_avg) {
                z0[kb * 8 + q] = m[q];                                              // global_max_pooling1d_1(x * a)
                z1[kb * 8 + q] = __fdiv_rn(sum_x[q], static_cast<float>(HEAD_T));   // global_average_pooling1d_1(x)
              } else {
                z0[kb * 8 + q] = __fdiv_rn(sum_w[q], static_cast<float>(HEAD_T));   // exp 106: mean_t(x * a)
              }
            }
          }
        }
        // ---- dense_2 + softmax ----
        l = -INFINITY;
        for (int j = 0; j < classes; ++j) {
          const float* wc = w2t + j * feat;
          float d = 0.0f;
#pragma unroll
          for (int kb = 0; kb < HEAD_MAX_KB; ++kb) {
            const int c0 = HEAD_BLK * kb + 8 * lane;
            if (c0 < C) {
              const float4 wa = *reinterpret_cast<const float4*>(wc + c0), wb = *reinterpret_cast<const float4*>(wc + c0 + 4);
              const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
              for (int q = 0; q < 8; ++q) d = fmaf(z0[kb * 8 + q], w[q], d);
              if (pool_max_avg) {
                const float4 va = *reinterpret_cast<const float4*>(wc + C + c0), vb = *reinterpret_cast<const float4*>(wc + C + c0 + 4);
                const float v[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
#pragma unroll
                for (int q = 0; q < 8; ++q) d = fmaf(z1[kb * 8 + q], v[q], d);
              }
            }
          }
          d = warp_sum(d);
          if (lane == j) l = d;
        }
        mx = warp_max(l);
        e = lane < classes ? expf(l - mx) : 0.0f;
        s = warp_sum(e);
        pv[(ci * n_views + v0 + u) * 32 + lane] = __fdiv_rn(e, s);
      }
    }
    __syncthreads();
    if (warp < cpi && clip0 + warp < n_clips) {           // warp w averages the views of clip clip0 + w
      const int b = clip0 + warp;
      float acc_p = 0.0f;
      for (int u = 0; u < n_views; ++u)
        acc_p = __fadd_rn(acc_p, pv[(warp * n_views + u) * 32 + lane]);   // probs + loud_probs + left_probs ...
      const float pm = __fdiv_rn(acc_p, static_cast<float>(n_views));   // ... / 3
      if (probs_mean && lane < classes) probs_mean[static_cast<size_t>(b) * classes + lane] = pm;
      // probs.argmax(axis=-1): first index among equal maxima
      float best = lane < classes ? pm : -INFINITY;
      int idx = lane < classes ? lane : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
      }
      if (argmax && lane == 0) argmax[b] = idx;
    }
    __syncthreads();
  }
}

}  // namespace

namespace {
template <typename T>
__global__ void to_float_kernel(const T* __restrict__ src, float* __restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    dst[i] = to_float(src[i]);
}
}  // namespace

// ---- head of conv_1d_time_sliced_model (model.py:759-765) ----
// GlobalAveragePooling1D as the A-operand loader of the hidden Dense layer's GEMM: A[row, c] = mean_t x[row, t, c]
template <typename TAct>
struct LoadGap {
  const TAct* x; int T, C;
  __device__ __forceinline__ float operator()(int m, int c) const {
    const TAct* base = x + static_cast<size_t>(m) * T * C + c;
    float s = 0.0f;
    for (int t = 0; t < T; ++t) s += to_float(base[static_cast<size_t>(t) * C]);
    return __fdiv_rn(s, static_cast<float>(T));
  }
};
struct EpiRelu6 {            // Dense(256, use_bias=False) -> Activation(relu6)
  float* C;
  __device__ __forceinline__ void operator()(int m, int n, float (&acc)[G_TM][G_TN], int M, int N) const {
#pragma unroll
    for (int i = 0; i < G_TM; ++i)
#pragma unroll
      for (int j = 0; j < G_TN; ++j)
        if (m + i < M && n + j < N) C[static_cast<size_t>(m + i) * N + n + j] = fminf(fmaxf(acc[i][j], 0.0f), 6.0f);
  }
};

namespace {
// Dense(classes, softmax, no bias) per view, TTA mean in view order, first-index argmax: one warp per clip
__global__ void __launch_bounds__(256) dense_softmax_tta_kernel(const float* __restrict__ hid, int hidden, int n_views, int n_clips,
                                                                const float* __restrict__ w2, int classes, int w_in_smem,
                                                                float* __restrict__ probs_mean, int32_t* __restrict__ argmax) {
  extern __shared__ float s_w2_buf[];                     // [hidden][classes] when it fits (w_in_smem), else read through L1 / L2
  const float* s_w2 = w2;
  if (w_in_smem) {
    for (int i = threadIdx.x; i < hidden * classes; i += blockDim.x) s_w2_buf[i] = __ldg(&w2[i]);
    __syncthreads();
    s_w2 = s_w2_buf;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  for (int b = blockIdx.x * wpb + warp; b < n_clips; b += gridDim.x * wpb) {
    float acc_p = 0.0f;
    for (int v = 0; v < n_views; ++v) {
      const float* x = hid + (static_cast<size_t>(b) * n_views + v) * hidden;
      float l = -INFINITY;
      if (lane < classes) {
        float d = 0.0f;
        for (int k = 0; k < hidden; ++k) d = fmaf(__ldg(&x[k]), s_w2[k * classes + lane], d);
        l = d;
      }
      const float mx = warp_max(l);
      const float e = lane < classes ? expf(l - mx) : 0.0f;
      const float s = warp_sum(e);
      acc_p = __fadd_rn(acc_p, __fdiv_rn(e, s));
    }
    const float pm = __fdiv_rn(acc_p, static_cast<float>(n_views));
    if (probs_mean && lane < classes) probs_mean[static_cast<size_t>(b) * classes + lane] = pm;
    float best = lane < classes ? pm : -INFINITY;
    int idx = lane < classes ? lane : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
    }
    if (argmax && lane == 0) argmax[b] = idx;
  }
}
}  // namespace

int launch_dense_softmax_tta(kws_handle* h, const float* hid, int hidden, int n_views, int n_clips, const float* w2, int classes,
                             float* probs_mean, int32_t* argmax, cudaStream_t st) {
  if (classes > HEAD_MAX_CLASSES) return fail(h, KWS_EUNSUPPORTED, "too many classes");
  const size_t w_bytes = static_cast<size_t>(hidden) * classes * sizeof(float);
  const int in_smem = w_bytes <= 48 * 1024 ? 1 : 0;
  const int grid = std::max(1, std::min(4 * h->num_sms, (n_clips + 7) / 8));
  KWS_T0(h, KC_HEAD, st);
  dense_softmax_tta_kernel<<<grid, 256, in_smem ? w_bytes : 0, st>>>(hid, hidden, n_views, n_clips, w2, classes, in_smem, probs_mean, argmax);
  KWS_T1(h, st);
  KWS_LAUNCH_CHECK(h);
  return KWS_OK;
}

static int launch_head_gap_dense(kws_handle* h, Model& m, const void* act, bool act_half, int n_clips, int n_views,
                                 float* probs_mean, int32_t* argmax, cudaStream_t st) {
  const int rows = n_clips * n_views;
  if (m.hidden_ws_rows < static_cast<size_t>(rows)) {
    if (m.hidden_ws) { KWS_CUDA(h, cudaStreamSynchronize(st)); cudaFree(m.hidden_ws); m.hidden_ws = nullptr; }
    KWS_CUDA(h, cudaMalloc(&m.hidden_ws, static_cast<size_t>(rows) * m.hidden * sizeof(float)));
    m.hidden_ws_rows = rows;
  }
  EpiRelu6 e{m.hidden_ws};
  KWS_T0(h, KC_HEAD, st);
  if (act_half) launch_gemm_f32(LoadGap<__half>{static_cast<const __half*>(act), m.t_last, m.c_last}, m.w_d1, rows, m.hidden, m.c_last, e, st);
  else launch_gemm_f32(LoadGap<float>{static_cast<const float*>(act), m.t_last, m.c_last}, m.w_d1, rows, m.hidden, m.c_last, e, st);
  KWS_T1(h, st);
  KWS_LAUNCH_CHECK(h);
  return launch_dense_softmax_tta(h, m.hidden_ws, m.hidden, n_views, n_clips, m.w_d2, m.classes, probs_mean, argmax, st);
}

int launch_to_float(kws_handle* h, const void* src, bool src_half, float* dst, size_t n, cudaStream_t st) {
  if (n == 0) return KWS_OK;
  const int grid = static_cast<int>(std::min<size_t>((n + 255) / 256, 4096));
  if (src_half) to_float_kernel<__half><<<grid, 256, 0, st>>>(static_cast<const __half*>(src), dst, n);
  else to_float_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(src), dst, n);
  KWS_LAUNCH_CHECK(h);
  return KWS_OK;
}

int launch_head(kws_handle* h, Model& m, const void* act, bool act_half, int n_clips, int n_views,
                float* probs_mean, int32_t* argmax, cudaStream_t st) {
  if (m.classes > HEAD_MAX_CLASSES) return fail(h, KWS_EUNSUPPORTED, "too many classes");
  if (n_views < 1 || n_views > HEAD_WARPS) return fail(h, KWS_EINVAL, "n_views must be in 1..16");
  if (m.head_kind == 1) return launch_head_gap_dense(h, m, act, act_half, n_clips, n_views, probs_mean, argmax, st);
  if (m.t_last != HEAD_T) return fail(h, KWS_EUNSUPPORTED, "head expects 9 time steps");
  if (m.classes > HEAD_MAX_CLASSES) return fail(h, KWS_EUNSUPPORTED, "too many classes");
  if (n_views < 1 || n_views > HEAD_WARPS) return fail(h, KWS_EINVAL, "n_views must be in 1..16");
  const int C = m.c_last;
  if (C > HEAD_MAX_KB * HEAD_BLK || C % 8) return fail(h, KWS_EUNSUPPORTED, "head expects at most 512 channels, a multiple of 8");
  const int feat = m.pool_max_avg ? 2 * C : C;
  const int ipw = n_views % 4 == 0 ? 4 : (n_views % 2 == 0 ? 2 : 1);   // views per warp (weights reused from registers)
  const int cpi = HEAD_WARPS / (n_views / ipw);
  const int nblk = (HEAD_T * C + HEAD_BLK - 1) / HEAD_BLK;
  const size_t smem = (static_cast<size_t>(nblk) * 8 * HEAD_WROW + static_cast<size_t>(m.classes) * feat +
                       static_cast<size_t>(cpi) * n_views * 32) * sizeof(float);
  if (smem > 227 * 1024) return fail(h, KWS_EUNSUPPORTED, "head weights do not fit in shared memory");
  const int grid = std::max(1, std::min(h->num_sms, (n_clips + cpi - 1) / cpi));
  const int pm = m.pool_max_avg ? 1 : 0;
  const uint32_t attr_bit = 32u << ((act_half ? 0 : 3) + (ipw == 4 ? 2 : ipw == 2 ? 1 : 0));
  auto launch = [&](auto kern, auto* a) -> int {
    if (!(h->smem_attr_done & attr_bit)) {
      KWS_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      h->smem_attr_done |= attr_bit;
    }
    KWS_T0(h, KC_HEAD, st);
    kern<<<grid, HEAD_WARPS * 32, smem, st>>>(a, C, n_views, n_clips, m.w_d1, m.b_d1, m.w_d2, m.classes, pm, probs_mean, argmax);
    KWS_T1(h, st);
    return KWS_OK;
  };
  int rc;
  if (act_half) {
    const __half* a = static_cast<const __half*>(act);
    rc = ipw == 4 ? launch(head_kernel<__half, 4>, a) : ipw == 2 ? launch(head_kernel<__half, 2>, a) : launch(head_kernel<__half, 1>, a);
  } else {
    const float* a = static_cast<const float*>(act);
    rc = ipw == 4 ? launch(head_kernel<float, 4>, a) : ipw == 2 ? launch(head_kernel<float, 2>, a) : launch(head_kernel<float, 1>, a);
  }
  if (rc) return rc;
  KWS_LAUNCH_CHECK(h);
  return KWS_OK;
}

}  // namespace kws
