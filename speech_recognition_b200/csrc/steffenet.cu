// steffeNet (reference model.py:1663-1726) on the fp32 CUDA-core GEMM with fused loaders / epilogues:
//
//   Reshape([-1, 1]) -> Conv1D(256, 75, strides=50, SAME, no bias) -> BN -> ReLU6
//   -> _context_conv(256, 3, SAME)                                   (depthwise k3 -> 1x1 conv -> BN -> ReLU6)
//   -> for nh in 320, 384, 512, 768, 1024, 1536:
//        _residual_block(nh, 3, strides=2): shortcut = BN(Conv1D(nh, 1, strides=2)); two SAME depthwise-separable
//                                            blocks (the first with stride 2); Add
//        _residual_block(nh, 3):            shortcut = x; two SAME depthwise-separable blocks; Add
//   -> GlobalMaxPooling1D || GlobalAveragePooling1D -> Dense(num_classes, softmax, no bias)
//
// The reference ships no checkpoint of it and never calls the builder; it is here because it re-uses the path's
// building blocks (depthwise FIR as the A-operand loader of the pointwise GEMM, BN + ReLU6 epilogue, TTA views applied
// while the waveform is read).  Channel counts go to 1536, beyond the tensor-core kernels' tiling (cout <= 512), so both
// precision tiers run these fp32 kernels.  Keras numbers layers in creation order: the shortcut Conv1D / BN of a
// stride-2 residual block come BEFORE the block's own layers (see oracle/network.py: steffenet_plan).
#include <cmath>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "gemm_f32.cuh"

namespace kws {

namespace {

constexpr int ST_K = 75, ST_STRIDE = 50, ST_C0 = 256;
const int kWidths[6] = {320, 384, 512, 768, 1024, 1536};

// implicit im2col of Conv1D(256, 75, strides=50, padding='same') on the (rolled, scaled) waveform:
// out length ceil(16000 / 50) = 320, pad_total = 319 * 50 + 75 - 16000 = 25 -> (12, 13)
struct LoadConv75 {
  const float* wav; int t_out; int n_views; int pad_left; ViewTable vt;
  __device__ __forceinline__ float operator()(int m, int k) const {
    const int r = m / t_out, j = m - r * t_out;
    const int b = r / n_views, v = r - b * n_views;
    const int p = ST_STRIDE * j - pad_left + k;
    if (p < 0 || p >= L) return 0.0f;
    int src = (p - vt.shift[v]) % L; if (src < 0) src += L;
    const float x = __ldg(&wav[static_cast<size_t>(b) * L + src]);
    const float g = vt.gain[v];
    return g == 1.0f ? x : __fmul_rn(g, x);
  }
};
// Conv1D(nh, 1, strides=2, padding='same'): A[(r, t), c] = x[r, 2 t, c]
struct LoadStride2 {
  const float* x; int t_in, t_out, cin;
  __device__ __forceinline__ float operator()(int m, int c) const {
    const int r = m / t_out, t = m - r * t_out;
    return __ldg(&x[(static_cast<size_t>(r) * t_in + 2 * t) * cin + c]);
  }
};
struct EpiBn {                  // BatchNormalization without activation (the shortcut branch)
  float* C; const float* scale; const float* shift;
  __device__ __forceinline__ void operator()(int m, int n, float (&acc)[G_TM][G_TN], int M, int N) const {
#pragma unroll
    for (int j = 0; j < G_TN; ++j) {
      if (n + j >= N) continue;
      const float s = __ldg(&scale[n + j]), sh = __ldg(&shift[n + j]);
#pragma unroll
      for (int i = 0; i < G_TM; ++i)
        if (m + i < M) C[static_cast<size_t>(m + i) * N + n + j] = fmaf(acc[i][j], s, sh);
    }
  }
};
struct EpiBnRelu6Add {          // relu6(BN(.)) + residual  (Add()([x, residual]), model.py:1696)
  float* C; const float* scale; const float* shift; const float* res;
  __device__ __forceinline__ void operator()(int m, int n, float (&acc)[G_TM][G_TN], int M, int N) const {
#pragma unroll
    for (int j = 0; j < G_TN; ++j) {
      if (n + j >= N) continue;
      const float s = __ldg(&scale[n + j]), sh = __ldg(&shift[n + j]);
#pragma unroll
      for (int i = 0; i < G_TM; ++i)
        if (m + i < M) {
          const size_t o = static_cast<size_t>(m + i) * N + n + j;
          C[o] = __fadd_rn(fminf(fmaxf(fmaf(acc[i][j], s, sh), 0.0f), 6.0f), __ldg(&res[o]));
        }
    }
  }
};

// z[row, 0:C] = max_t x[row, t, :], z[row, C:2C] = mean_t x[row, t, :]
__global__ void max_avg_pool_kernel(const float* __restrict__ x, int rows, int T, int C, float* __restrict__ z) {
  const size_t n = static_cast<size_t>(rows) * C;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t r = i / C; const int c = static_cast<int>(i - r * C);
    float mx = -INFINITY, sum = 0.0f;
    for (int t = 0; t < T; ++t) { const float v = __ldg(&x[(r * T + t) * C + c]); mx = fmaxf(mx, v); sum += v; }
    z[r * 2 * C + c] = mx;
    z[r * 2 * C + C + c] = __fdiv_rn(sum, static_cast<float>(T));
  }
}

enum { SK_DWPW = 0, SK_SHORTCUT = 1, SK_IDENTITY = 2, SK_DWPW_ADD = 3 };
struct Step { int kind, cin, cout, stride, t_in, t_out, pad_left; size_t o_dw, o_pw, o_scale, o_shift; };

}  // namespace

struct SteffeNet {
  std::vector<Step> steps;
  float* blob = nullptr;
  size_t o_conv = 0, o_scale0 = 0, o_shift0 = 0, o_d1 = 0;
  int t0 = 0, pad0 = 0, c_last = 0, t_last = 0;
  size_t max_elems = 0;                  // per clip-view, floats
  float* third = nullptr; float* pooled = nullptr; size_t rows_alloc = 0;
};

void steffe_free(SteffeNet* s) {
  if (!s) return;
  if (s->blob) cudaFree(s->blob);
  if (s->third) cudaFree(s->third);
  if (s->pooled) cudaFree(s->pooled);
  delete s;
}

int steffe_build(kws_handle* h, Model& m, const kws_tensor_h* t, int n) {
  std::map<std::string, const kws_tensor_h*> by_name;
  for (int i = 0; i < n; ++i) {
    if (!t[i].name || !t[i].data) return fail(h, KWS_EINVAL, "null tensor entry");
    by_name[t[i].name] = &t[i];
  }
  std::vector<float> host;
  auto push = [&](const float* p, size_t cnt) { size_t o = host.size(); host.insert(host.end(), p, p + cnt); while (host.size() % 4) host.push_back(0.f); return o; };
  auto get = [&](const std::string& name, int64_t numel, const float** out) -> int {
    auto it = by_name.find(name);
    if (it == by_name.end()) return fail(h, KWS_EINVAL, "missing tensor '" + name + "'");
    if (it->second->numel != numel)
      return fail(h, KWS_EINVAL, "tensor '" + name + "' has " + std::to_string(it->second->numel) + " elements, expected " + std::to_string(numel));
    *out = it->second->data;
    return KWS_OK;
  };
  auto fold_bn = [&](int idx, int ch, size_t* o_scale, size_t* o_shift) -> int {
    const float *g, *b, *mu, *var;
    const std::string base = "batch_normalization_" + std::to_string(idx) + "/";
    int r;
    if ((r = get(base + "gamma", ch, &g)) || (r = get(base + "beta", ch, &b)) || (r = get(base + "moving_mean", ch, &mu)) ||
        (r = get(base + "moving_variance", ch, &var))) return r;
    std::vector<float> sc(ch), sh(ch);
    for (int c = 0; c < ch; ++c) {
      const double s = static_cast<double>(g[c]) / std::sqrt(static_cast<double>(var[c]) + 1e-3);
      sc[c] = static_cast<float>(s);
      sh[c] = static_cast<float>(static_cast<double>(b[c]) - static_cast<double>(mu[c]) * s);
    }
    *o_scale = push(sc.data(), ch); *o_shift = push(sh.data(), ch);
    return KWS_OK;
  };
  SteffeNet* s = new SteffeNet();
  auto bail = [&](int rc) { steffe_free(s); return rc; };
  const float* p = nullptr;
  int rc;
  int conv = 1, bn = 1, dw = 0;
  // stem
  s->t0 = (L + ST_STRIDE - 1) / ST_STRIDE;                                    // 320
  s->pad0 = std::max((s->t0 - 1) * ST_STRIDE + ST_K - L, 0) / 2;              // 12
  if ((rc = get("conv1d_1/kernel", 1LL * ST_K * ST_C0, &p))) return bail(rc);
  s->o_conv = push(p, static_cast<size_t>(ST_K) * ST_C0);
  if ((rc = fold_bn(bn, ST_C0, &s->o_scale0, &s->o_shift0))) return bail(rc);
  ++conv; ++bn;
  int T = s->t0, C = ST_C0;
  s->max_elems = static_cast<size_t>(T) * C;
  auto add_dwpw = [&](int cout, int stride, bool with_add) -> int {
    Step st{};
    st.kind = with_add ? SK_DWPW_ADD : SK_DWPW; st.cin = C; st.cout = cout; st.stride = stride; st.t_in = T;
    st.t_out = (T + stride - 1) / stride;
    st.pad_left = std::max((st.t_out - 1) * stride + 3 - T, 0) / 2;           // SAME: (1, 1) at stride 1, (0, 1) at stride 2 (T even)
    ++dw;
    int r;
    if ((r = get("depthwise_conv2d_" + std::to_string(dw) + "/depthwise_kernel", 3LL * C, &p))) return r;
    st.o_dw = push(p, 3 * static_cast<size_t>(C));
    if ((r = get("conv1d_" + std::to_string(conv) + "/kernel", 1LL * C * cout, &p))) return r;
    st.o_pw = push(p, static_cast<size_t>(C) * cout);
    if ((r = fold_bn(bn, cout, &st.o_scale, &st.o_shift))) return r;
    ++conv; ++bn;
    T = st.t_out; C = cout;
    s->max_elems = std::max(s->max_elems, static_cast<size_t>(T) * C);
    s->steps.push_back(st);
    return KWS_OK;
  };
  if ((rc = add_dwpw(ST_C0, 1, false))) return bail(rc);                      // _context_conv(256, 3, SAME)
  for (int wi = 0; wi < 6; ++wi) {
    const int nh = kWidths[wi];
    for (int stride = 2; stride >= 1; --stride) {
      Step sc{};
      if (stride == 2) {
        sc.kind = SK_SHORTCUT; sc.cin = C; sc.cout = nh; sc.stride = 2; sc.t_in = T; sc.t_out = (T + 1) / 2;
        if ((rc = get("conv1d_" + std::to_string(conv) + "/kernel", 1LL * C * nh, &p))) return bail(rc);
        sc.o_pw = push(p, static_cast<size_t>(C) * nh);
        if ((rc = fold_bn(bn, nh, &sc.o_scale, &sc.o_shift))) return bail(rc);
        ++conv; ++bn;
      } else {
        sc.kind = SK_IDENTITY;
      }
      s->steps.push_back(sc);
      if ((rc = add_dwpw(nh, stride, false))) return bail(rc);
      if ((rc = add_dwpw(nh, 1, true))) return bail(rc);
    }
  }
  s->c_last = C; s->t_last = T;
  auto it = by_name.find("dense_1/kernel");
  if (it == by_name.end() || it->second->numel % (2 * C) || it->second->numel / (2 * C) > 32)
    return bail(fail(h, KWS_EINVAL, "dense_1/kernel missing or not [2 * 1536, classes <= 32]"));
  m.classes = static_cast<int>(it->second->numel / (2 * C));
  s->o_d1 = push(it->second->data, static_cast<size_t>(it->second->numel));
  if (cudaMalloc(&s->blob, host.size() * sizeof(float)) != cudaSuccess) return bail(fail(h, KWS_ENOMEM, "steffeNet weights do not fit"));
  if (cudaMemcpy(s->blob, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
    return bail(fail(h, KWS_ECUDA, "steffeNet weight upload failed"));
  m.steffe = s;
  m.c0 = ST_C0; m.t0 = s->t0; m.t_last = T; m.c_last = C;
  return KWS_OK;
}

int launch_forward_steffe(kws_handle* h, Model& m, const float* wav, int B, const ViewTable& vt, float* probs_mean, int32_t* argmax,
                          cudaStream_t st) {
  SteffeNet* s = m.steffe;
  const int V = vt.n;
  const int clips_per_chunk = std::max(1, h->max_rows / V);
  const size_t rows_max = static_cast<size_t>(std::min(clips_per_chunk, B)) * V;
  const size_t need = rows_max * s->max_elems * sizeof(float);
  if (h->act_bytes < need) {
    KWS_CUDA(h, cudaStreamSynchronize(st));
    h->tmap_cache.clear();
    for (int i = 0; i < 2; ++i) {
      if (h->act[i]) cudaFree(h->act[i]);
      h->act[i] = nullptr;
      KWS_CUDA(h, cudaMalloc(&h->act[i], need));
    }
    h->act_bytes = need;
  }
  if (s->rows_alloc < rows_max) {
    KWS_CUDA(h, cudaStreamSynchronize(st));
    if (s->third) cudaFree(s->third);
    if (s->pooled) cudaFree(s->pooled);
    s->third = nullptr; s->pooled = nullptr;
    KWS_CUDA(h, cudaMalloc(&s->third, need));
    KWS_CUDA(h, cudaMalloc(&s->pooled, rows_max * 2 * s->c_last * sizeof(float)));
    s->rows_alloc = rows_max;
  }
  const float* W = s->blob;
  for (int b0 = 0; b0 < B; b0 += clips_per_chunk) {
    const int nb = std::min(clips_per_chunk, B - b0);
    const int rows = nb * V;
    float* X = static_cast<float*>(h->act[0]);      // current activation (and identity shortcut)
    float* Y = static_cast<float*>(h->act[1]);
    float* Z = s->third;
    {
      LoadConv75 a{wav + static_cast<size_t>(b0) * L, s->t0, V, s->pad0, vt};
      EpiBnRelu6 e{X, W + s->o_scale0, W + s->o_shift0};
      KWS_T0(h, KC_CONV1, st);
      launch_gemm_f32(a, W + s->o_conv, rows * s->t0, ST_C0, ST_K, e, st);
      KWS_T1(h, st);
      KWS_LAUNCH_CHECK(h);
    }
    const float* res = nullptr;                     // shortcut of the residual block in flight
    bool res_is_x = false;
    for (const Step& sp : s->steps) {
      if (sp.kind == SK_SHORTCUT) {                 // X -> Z (BN, no activation); the block below reads X too
        LoadStride2 a{X, sp.t_in, sp.t_out, sp.cin};
        EpiBn e{Z, W + sp.o_scale, W + sp.o_shift};
        KWS_T0(h, KC_BLOCKS, st);
        launch_gemm_f32(a, W + sp.o_pw, rows * sp.t_out, sp.cout, sp.cin, e, st);
        KWS_T1(h, st);
        KWS_LAUNCH_CHECK(h);
        res = Z; res_is_x = false;
      } else if (sp.kind == SK_IDENTITY) {
        res = X; res_is_x = true;
      } else if (sp.kind == SK_DWPW) {              // X -> Y
        LoadDepthwise a{X, W + sp.o_dw, sp.t_in, sp.t_out, sp.cin, sp.stride, sp.pad_left};
        EpiBnRelu6 e{Y, W + sp.o_scale, W + sp.o_shift};
        KWS_T0(h, KC_BLOCKS, st);
        launch_gemm_f32(a, W + sp.o_pw, rows * sp.t_out, sp.cout, sp.cin, e, st);
        KWS_T1(h, st);
        KWS_LAUNCH_CHECK(h);
        if (res == nullptr) std::swap(X, Y);        // the stem's _context_conv: no residual block around it
      } else {                                      // second block of a residual block: Y -> (free buffer) + res
        float* out = res_is_x ? Z : X;              // identity: X is the shortcut, Z is free; strided: Z is the shortcut, X is free
        LoadDepthwise a{Y, W + sp.o_dw, sp.t_in, sp.t_out, sp.cin, sp.stride, sp.pad_left};
        EpiBnRelu6Add e{out, W + sp.o_scale, W + sp.o_shift, res};
        KWS_T0(h, KC_BLOCKS, st);
        launch_gemm_f32(a, W + sp.o_pw, rows * sp.t_out, sp.cout, sp.cin, e, st);
        KWS_T1(h, st);
        KWS_LAUNCH_CHECK(h);
        if (res_is_x) std::swap(X, Z);              // the sum becomes the current activation
        res = nullptr;
        s->third = Z;                               // (pointer roles rotate; ownership stays with act[0], act[1], third)
      }
    }
    h->act[0] = X; h->act[1] = Y; s->third = Z;
    const size_t n = static_cast<size_t>(rows) * s->c_last;
    KWS_T0(h, KC_HEAD, st);
    max_avg_pool_kernel<<<static_cast<int>(std::min<size_t>((n + 255) / 256, 8192)), 256, 0, st>>>(X, rows, s->t_last, s->c_last, s->pooled);
    KWS_T1(h, st);
    KWS_LAUNCH_CHECK(h);
    const int rc = launch_dense_softmax_tta(h, s->pooled, 2 * s->c_last, V, nb, W + s->o_d1, m.classes,
                                            probs_mean ? probs_mean + static_cast<size_t>(b0) * m.classes : nullptr,
                                            argmax ? argmax + b0 : nullptr, st);
    if (rc) return rc;
  }
  return KWS_OK;
}

}  // namespace kws
