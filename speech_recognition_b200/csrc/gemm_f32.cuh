// fp32 CUDA-core GEMM with a fused A-operand loader and a fused epilogue.
//   C[m, n] = epilogue( sum_k A(m, k) * W[k, n] )
// A(m,k) is produced on the fly by a loader functor (implicit STFT framing,
// implicit im2col of the raw waveform with the TTA view applied, depthwise k=3
// FIR of the previous activation), so none of these intermediates touches HBM.
// This is the KWS_PREC_FP32 tier (1e-4 parity against the oracle); the
// KWS_PREC_TC tier runs the same contractions on tcgen05 (tc_*.cu).
#pragma once
#include "common.cuh"

namespace kws {

constexpr int G_BM = 128, G_BN = 128, G_BK = 16, G_TM = 8, G_TN = 8;
constexpr int G_THREADS = (G_BM / G_TM) * (G_BN / G_TN);   // 256

template <class ALoad, class Epi>
__global__ void __launch_bounds__(G_THREADS)
gemm_f32_kernel(ALoad aload, const float* __restrict__ W, int M, int N, int K, Epi epi) {
  __shared__ __align__(16) float As[G_BK][G_BM + 4];
  __shared__ __align__(16) float Bs[G_BK][G_BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % (G_BN / G_TN), ty = tid / (G_BN / G_TN);
  const int m0 = blockIdx.x * G_BM, n0 = blockIdx.y * G_BN;

  float acc[G_TM][G_TN];
#pragma unroll
  for (int i = 0; i < G_TM; ++i)
#pragma unroll
    for (int j = 0; j < G_TN; ++j) acc[i][j] = 0.0f;

  // A tile: thread -> (k = tid % 16, rows tid/16 + 16*i); B tile: (n = tid % 128, k = tid/128 + 2*i)
  const int ak = tid % G_BK, am = tid / G_BK;
  const int bn = tid % G_BN, bk = tid / G_BN;

  for (int k0 = 0; k0 < K; k0 += G_BK) {
    float areg[G_BM * G_BK / G_THREADS];
    float breg[G_BK * G_BN / G_THREADS];
#pragma unroll
    for (int i = 0; i < G_BM * G_BK / G_THREADS; ++i) {
      const int m = m0 + am + (G_THREADS / G_BK) * i;
      const int k = k0 + ak;
      areg[i] = (m < M && k < K) ? aload(m, k) : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < G_BK * G_BN / G_THREADS; ++i) {
      const int k = k0 + bk + (G_THREADS / G_BN) * i;
      const int n = n0 + bn;
      breg[i] = (k < K && n < N) ? __ldg(&W[static_cast<size_t>(k) * N + n]) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < G_BM * G_BK / G_THREADS; ++i) As[ak][am + (G_THREADS / G_BK) * i] = areg[i];
#pragma unroll
    for (int i = 0; i < G_BK * G_BN / G_THREADS; ++i) Bs[bk + (G_THREADS / G_BN) * i][bn] = breg[i];
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < G_BK; ++kk) {
      float a[G_TM], b[G_TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * G_TM]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * G_TM + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * G_TN]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][tx * G_TN + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < G_TM; ++i)
#pragma unroll
        for (int j = 0; j < G_TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  epi(m0 + ty * G_TM, n0 + tx * G_TN, acc, M, N);
}

// ---------------- A loaders ----------------

// plain row-major A[M,K]
struct LoadPlain {
  const float* A; int lda;
  __device__ __forceinline__ float operator()(int m, int k) const {
    return __ldg(&A[static_cast<size_t>(m) * lda + k]);
  }
};

// implicit STFT framing (input_data.py:361-365): A[b*frames + f, k] = x[b, hop*f + k]
struct LoadFrames {
  const float* x; int frames, hop;
  __device__ __forceinline__ float operator()(int m, int k) const {
    const int b = m / frames, f = m - b * frames;
    return __ldg(&x[static_cast<size_t>(b) * L + hop * f + k]);
  }
};

// implicit im2col of overlapping_time_slice_stack (k40 s20 SAME, pad 10/10) followed by the
// k3 s2 VALID Conv1D (model.py:805-807): row (r, j) covers samples 40j-10 .. 40j+69 through
// 3 patches of 40 at hop 20 => k = f*40 + i  ->  sample 40j - 10 + 20f + i (zero outside).
// The TTA view (np.roll shift + gain, make_submission.py:126-130) is applied in the load.
struct LoadSliceConv1 {
  const float* wav; int t_out; int n_views; ViewTable vt;
  __device__ __forceinline__ float operator()(int m, int k) const {
    const int r = m / t_out, j = m - r * t_out;
    const int b = r / n_views, v = r - b * n_views;
    const int f = k / 40, i = k - f * 40;
    const int p = 40 * j - 10 + 20 * f + i;
    if (p < 0 || p >= L) return 0.0f;
    int src = p - vt.shift[v];
    src %= L; if (src < 0) src += L;
    const float x = __ldg(&wav[static_cast<size_t>(b) * L + src]);
    const float g = vt.gain[v];
    return g == 1.0f ? x : __fmul_rn(g, x);
  }
};

// depthwise k=3 FIR of the previous channels-last activation (model.py:40-43):
// A[(r,t), c] = sum_j wd[j,c] * x[r, t*stride + j - pad_left, c]  (zero outside: TF 'SAME')
struct LoadDepthwise {
  const float* x; const float* wd; int t_in, t_out, cin, stride, pad_left;
  __device__ __forceinline__ float operator()(int m, int c) const {
    const int r = m / t_out, t = m - r * t_out;
    const float* base = x + static_cast<size_t>(r) * t_in * cin + c;
    float acc = 0.0f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int ti = t * stride + j - pad_left;
      if (ti >= 0 && ti < t_in) acc = fmaf(__ldg(&wd[j * cin + c]), __ldg(&base[static_cast<size_t>(ti) * cin]), acc);
    }
    return acc;
  }
};

// ---------------- epilogues ----------------

struct EpiStore {            // plain store
  float* C; int ldc;
  __device__ __forceinline__ void operator()(int m, int n, float (&acc)[G_TM][G_TN], int M, int N) const {
#pragma unroll
    for (int i = 0; i < G_TM; ++i)
#pragma unroll
      for (int j = 0; j < G_TN; ++j)
        if (m + i < M && n + j < N) C[static_cast<size_t>(m + i) * ldc + n + j] = acc[i][j];
  }
};

struct EpiBnRelu6 {          // BatchNorm (inference) + ReLU6, model.py:50-51 / :30-31
  float* C; const float* scale; const float* shift;
  __device__ __forceinline__ void operator()(int m, int n, float (&acc)[G_TM][G_TN], int M, int N) const {
#pragma unroll
    for (int j = 0; j < G_TN; ++j) {
      if (n + j >= N) continue;
      const float s = __ldg(&scale[n + j]), sh = __ldg(&shift[n + j]);
#pragma unroll
      for (int i = 0; i < G_TM; ++i)
        if (m + i < M) C[static_cast<size_t>(m + i) * N + n + j] = fminf(fmaxf(fmaf(acc[i][j], s, sh), 0.0f), 6.0f);
    }
  }
};

struct EpiMagnitude {        // columns (2j, 2j+1) = (re, im) of bin j -> |X| (input_data.py:366), or |X|^2 (audio.py:15-19)
  float* S; int n_bins; bool power;
  __device__ __forceinline__ void operator()(int m, int n, float (&acc)[G_TM][G_TN], int M, int N) const {
#pragma unroll
    for (int i = 0; i < G_TM; ++i)
#pragma unroll
      for (int j = 0; j < G_TN; j += 2) {
        const int bin = (n + j) >> 1;
        if (m + i < M && bin < n_bins) {
          const float re = acc[i][j], im = acc[i][j + 1];
          const float pw = fmaf(re, re, im * im);
          S[static_cast<size_t>(m + i) * n_bins + bin] = power ? pw : sqrtf(pw);
        }
      }
  }
};

struct EpiLog {              // log(mel + 1e-6)  (input_data.py:378), or log(max(mel, 1e-12)) (contrib_audio Mfcc)
  float* C; int ldc; bool floor_mode;
  __device__ __forceinline__ void operator()(int m, int n, float (&acc)[G_TM][G_TN], int M, int N) const {
#pragma unroll
    for (int i = 0; i < G_TM; ++i)
#pragma unroll
      for (int j = 0; j < G_TN; ++j)
        if (m + i < M && n + j < N)
          C[static_cast<size_t>(m + i) * ldc + n + j] = floor_mode ? logf(fmaxf(acc[i][j], 1e-12f)) : logf(acc[i][j] + 1e-6f);
  }
};

template <class ALoad, class Epi>
inline void launch_gemm_f32(const ALoad& a, const float* W, int M, int N, int K, const Epi& e, cudaStream_t st) {
  dim3 grid((M + G_BM - 1) / G_BM, (N + G_BN - 1) / G_BN);
  gemm_f32_kernel<ALoad, Epi><<<grid, G_THREADS, 0, st>>>(a, W, M, N, K, e);
}

}  // namespace kws
