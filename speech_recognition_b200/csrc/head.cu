// K7 head -- attention pooling + classifier + TTA mean + argmax, one CTA per clip.
// Replaces, per view (reference model.py:819-830):
//   Flatten -> Dense(T, softmax) -> x * a[:, None] -> GlobalMaxPool1D || GlobalAveragePooling1D
//   -> Dense(classes, softmax)            (exp 195/206)
//   Flatten -> Dense(T, softmax, no bias) -> mean_t(x * a) -> Dense(32, softmax)   (exp 106)
// and across views (make_submission.py:137-146): probs = (p_0 + p_1 + ...) / n_views in view
// order, argmax with first-index tie rule.  Memory/latency-bound CUDA-core work with
// warp-shuffle reductions; the weights (166 KB + 48 KB) stay L2/L1-resident.
#include <algorithm>

#include "common.cuh"

namespace kws {

namespace {

constexpr int HEAD_THREADS = 256;
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

template <typename TAct>
__global__ void __launch_bounds__(HEAD_THREADS)
head_kernel(const TAct* __restrict__ act, int C, int n_views, const float* __restrict__ w_d1,
            const float* __restrict__ b_d1, const float* __restrict__ w_d2, int classes,
            int pool_max_avg, float* __restrict__ probs_mean, int32_t* __restrict__ argmax) {
  extern __shared__ float sm[];
  float* xs = sm;                               // [HEAD_T * C]
  float* z = xs + HEAD_T * C;                   // [2*C]
  float* red = z + 2 * C;                       // [8 warps][HEAD_T]
  float* att = red + (HEAD_THREADS / 32) * HEAD_T;   // [HEAD_T]
  float* logits = att + 16;                     // [HEAD_MAX_CLASSES]

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = HEAD_T * C;
  const int feat = pool_max_avg ? 2 * C : C;
  float acc_p = 0.0f;                           // warp 0, lane j: running sum of class j

  for (int v = 0; v < n_views; ++v) {
    const TAct* x = act + (static_cast<size_t>(b) * n_views + v) * n;
    float p[HEAD_T];
#pragma unroll
    for (int j = 0; j < HEAD_T; ++j) p[j] = 0.0f;
    for (int i = tid; i < n; i += HEAD_THREADS) {
      const float xv = to_float(x[i]);
      xs[i] = xv;
      const float* wr = w_d1 + static_cast<size_t>(i) * HEAD_T;
#pragma unroll
      for (int j = 0; j < HEAD_T; ++j) p[j] = fmaf(xv, __ldg(&wr[j]), p[j]);
    }
#pragma unroll
    for (int j = 0; j < HEAD_T; ++j) {
      const float s = warp_sum(p[j]);
      if (lane == 0) red[warp * HEAD_T + j] = s;
    }
    __syncthreads();
    if (warp == 0) {                            // attention softmax over T (dense_1)
      float l = -INFINITY;
      if (lane < HEAD_T) {
        l = __ldg(&b_d1[lane]);
#pragma unroll
        for (int w = 0; w < HEAD_THREADS / 32; ++w) l += red[w * HEAD_T + lane];
      }
      const float mx = warp_max(l);
      const float e = lane < HEAD_T ? expf(l - mx) : 0.0f;
      const float s = warp_sum(e);
      if (lane < HEAD_T) att[lane] = __fdiv_rn(e, s);
    }
    __syncthreads();
    for (int c = tid; c < C; c += HEAD_THREADS) {
      float mx = -INFINITY, sum_x = 0.0f, sum_w = 0.0f;
#pragma unroll
      for (int t = 0; t < HEAD_T; ++t) {
        const float xv = xs[t * C + c];
        const float wv = __fmul_rn(xv, att[t]);          // multiply_1
        mx = fmaxf(mx, wv);
        sum_x += xv;
        sum_w += wv;
      }
      if (pool_max_avg) {
        z[c] = mx;                                       // global_max_pooling1d_1(x * a)
        z[C + c] = __fdiv_rn(sum_x, static_cast<float>(HEAD_T));   // global_average_pooling1d_1(x)
      } else {
        z[c] = __fdiv_rn(sum_w, static_cast<float>(HEAD_T));       // exp 106: mean_t(x * a)
      }
    }
    __syncthreads();
    for (int j = warp; j < classes; j += HEAD_THREADS / 32) {     // dense_2
      float s = 0.0f;
      for (int i = lane; i < feat; i += 32) s = fmaf(z[i], __ldg(&w_d2[static_cast<size_t>(i) * classes + j]), s);
      s = warp_sum(s);
      if (lane == 0) logits[j] = s;
    }
    __syncthreads();
    if (warp == 0) {
      const float l = lane < classes ? logits[lane] : -INFINITY;
      const float mx = warp_max(l);
      const float e = lane < classes ? expf(l - mx) : 0.0f;
      const float s = warp_sum(e);
      acc_p = __fadd_rn(acc_p, __fdiv_rn(e, s));         // probs + loud_probs + left_probs ...
    }
    __syncthreads();
  }
  if (warp == 0) {
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
  if (m.t_last != HEAD_T) return fail(h, KWS_EUNSUPPORTED, "head expects 9 time steps");
  if (m.classes > HEAD_MAX_CLASSES) return fail(h, KWS_EUNSUPPORTED, "too many classes");
  const int C = m.c_last;
  const size_t smem = (static_cast<size_t>(HEAD_T) * C + 2 * C + (HEAD_THREADS / 32) * HEAD_T + 16 +
                       HEAD_MAX_CLASSES) * sizeof(float);
  KWS_T0(h, KC_HEAD, st);
  if (act_half) {
    head_kernel<__half><<<n_clips, HEAD_THREADS, smem, st>>>(
        static_cast<const __half*>(act), C, n_views, m.w_d1, m.b_d1, m.w_d2, m.classes,
        m.pool_max_avg ? 1 : 0, probs_mean, argmax);
  } else {
    head_kernel<float><<<n_clips, HEAD_THREADS, smem, st>>>(
        static_cast<const float*>(act), C, n_views, m.w_d1, m.b_d1, m.w_d2, m.classes,
        m.pool_max_avg ? 1 : 0, probs_mean, argmax);
  }
  KWS_T1(h, st);
  KWS_LAUNCH_CHECK(h);
  return KWS_OK;
}

}  // namespace kws
