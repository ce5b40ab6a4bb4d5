// Weight ingestion for the raw-waveform Depthwise1D networks (exp 195/206:
// reference model.py:775-838; exp 106: logs_106 graph) and the fp32 forward.
// Tensors arrive as host fp32 arrays keyed by their Keras variable names, exactly
// what keras.models.load_model (make_submission.py:64-71) would have produced.
#include <cmath>
#include <cstring>
#include <map>

#include "common.cuh"
#include "gemm_f32.cuh"

namespace kws {

int model_build_tc(kws_handle* h, Model& m, const std::vector<std::vector<float>>& pw_host,
                   const std::vector<float>& conv1_host, const std::vector<std::vector<float>>& dw_host,
                   const std::vector<std::vector<float>>& scales);   // tc_net.cu

namespace {

// classes == 0: taken from the shape of dense_2/kernel (conv_1d_time_sliced_model has a num_classes argument)
struct ArchSpec { int conv1; int n_blocks; int blocks[MAX_BLOCKS][2]; bool bias; bool max_avg; int classes; int head_kind; int hidden; };

const ArchSpec kArch195 = {128, 11, {{128, 1}, {192, 2}, {192, 1}, {256, 2}, {256, 1}, {320, 2}, {320, 1},
                                     {384, 2}, {384, 1}, {512, 2}, {512, 1}}, true, true, 12, 0, 0};
const ArchSpec kArch106 = {64, 11, {{128, 1}, {192, 2}, {192, 1}, {256, 2}, {256, 1}, {320, 2}, {320, 1},
                                    {384, 2}, {384, 1}, {448, 2}, {448, 1}}, false, false, 32, 0, 0};
// conv_1d_time_sliced_model(filter_mult=1), model.py:716-772: conv1d_1 has 32 filters, 13 blocks, and the head is
// GlobalAveragePooling1D -> Dense(256, no bias) -> ReLU6 -> Dense(num_classes, softmax, no bias)
const ArchSpec kArchTimeSliced = {32, 13, {{64, 1}, {128, 2}, {128, 1}, {192, 2}, {192, 1}, {256, 2}, {256, 1}, {320, 2}, {320, 1},
                                           {384, 2}, {384, 1}, {512, 2}, {512, 1}}, false, false, 0, 1, 256};

void same_pad(int T, int k, int s, int* out, int* pad_left) {   // TF 'SAME'
  *out = (T + s - 1) / s;
  int total = (*out - 1) * s + k - T;
  if (total < 0) total = 0;
  *pad_left = total / 2;
}

}  // namespace

int model_build(kws_handle* h, int slot, int arch, const kws_tensor_h* t, int n) {
  if (slot < 0 || slot >= KWS_MAX_MODELS) return fail(h, KWS_EINVAL, "model slot out of range");
  const ArchSpec* spec = nullptr;
  if (arch == KWS_ARCH_STEFFENET) {
    Model& sm = h->models[slot];
    if (sm.blob) { cudaFree(sm.blob); sm.blob = nullptr; }
    if (sm.tc_blob) { cudaFree(sm.tc_blob); sm.tc_blob = nullptr; }
    if (sm.hidden_ws) { cudaFree(sm.hidden_ws); sm.hidden_ws = nullptr; }
    if (sm.steffe) { steffe_free(sm.steffe); sm.steffe = nullptr; }
    sm = Model();
    sm.arch = arch;
    const int rc = steffe_build(h, sm, t, n);
    if (rc == KWS_OK) sm.loaded = true;
    return rc;
  }
  if (arch == KWS_ARCH_195 || arch == 206) spec = &kArch195;
  else if (arch == KWS_ARCH_106) spec = &kArch106;
  else if (arch == KWS_ARCH_TIME_SLICED) spec = &kArchTimeSliced;
  else return fail(h, KWS_EUNSUPPORTED, "unknown architecture " + std::to_string(arch) +
                                            " (this path ships 195 / 206, 106 and conv_1d_time_sliced)");
  std::map<std::string, const kws_tensor_h*> by_name;
  for (int i = 0; i < n; ++i) {
    if (!t[i].name || !t[i].data) return fail(h, KWS_EINVAL, "null tensor entry");
    by_name[t[i].name] = &t[i];
  }
  auto get = [&](const std::string& name, int64_t numel, const float** out) -> int {
    auto it = by_name.find(name);
    if (it == by_name.end()) return fail(h, KWS_EINVAL, "missing tensor '" + name + "'");
    if (it->second->numel != numel)
      return fail(h, KWS_EINVAL, "tensor '" + name + "' has " + std::to_string(it->second->numel) +
                                     " elements, expected " + std::to_string(numel));
    *out = it->second->data;
    return KWS_OK;
  };

  Model& m = h->models[slot];
  if (m.blob) { cudaFree(m.blob); m.blob = nullptr; }
  if (m.tc_blob) { cudaFree(m.tc_blob); m.tc_blob = nullptr; }
  if (m.hidden_ws) { cudaFree(m.hidden_ws); m.hidden_ws = nullptr; }
  if (m.steffe) { steffe_free(m.steffe); m.steffe = nullptr; }
  m = Model();
  m.arch = arch; m.classes = spec->classes; m.c0 = spec->conv1;
  m.dense1_bias = spec->bias; m.pool_max_avg = spec->max_avg;
  m.n_blocks = spec->n_blocks; m.head_kind = spec->head_kind; m.hidden = spec->hidden;
  m.c0_tc = (m.c0 + 63) / 64 * 64;
  if (m.classes == 0) {
    auto it = by_name.find("dense_2/kernel");
    if (it == by_name.end() || m.hidden == 0 || it->second->numel % m.hidden || it->second->numel / m.hidden > 32)
      return fail(h, KWS_EINVAL, "dense_2/kernel missing or not [hidden, classes <= 32]");
    m.classes = static_cast<int>(it->second->numel / m.hidden);
  }
  int n_patch, pl;
  same_pad(L, 40, 20, &n_patch, &pl);                      // 800 patches, pad (10,10)
  m.t0 = (n_patch - 3) / 2 + 1;                            // 399
  int T = m.t0, C = m.c0;
  m.max_act_elems = static_cast<size_t>(T) * m.c0_tc;
  for (int i = 0; i < m.n_blocks; ++i) {
    LayerDesc& d = m.layers[i];
    d.cin = C; d.cout = spec->blocks[i][0]; d.stride = spec->blocks[i][1]; d.t_in = T;
    if (d.stride == 1) { d.t_out = T - 2; d.pad_left = 0; }
    else same_pad(T, 3, 2, &d.t_out, &d.pad_left);
    T = d.t_out; C = d.cout;
    m.max_act_elems = std::max(m.max_act_elems, static_cast<size_t>(T) * C);
  }
  m.t_last = T; m.c_last = C;
  const int feat = m.head_kind == 1 ? m.hidden : (m.pool_max_avg ? 2 * C : C);

  // ---- gather + fold on the host ----
  std::vector<float> host;
  auto push = [&](const float* p, size_t cnt) { size_t o = host.size(); host.insert(host.end(), p, p + cnt); return o; };
  auto push_pad = [&]() { while (host.size() % 4) host.push_back(0.f); };
  const float* p = nullptr;
  int rc;
  if ((rc = get("conv1d_1/kernel", 120LL * m.c0, &p))) return rc;
  std::vector<float> conv1_host(p, p + 120 * m.c0);
  size_t o_conv1 = push(p, 120 * m.c0); push_pad();
  size_t o_scale[MAX_BLOCKS + 1], o_shift[MAX_BLOCKS + 1], o_dw[MAX_BLOCKS], o_pw[MAX_BLOCKS];
  std::vector<std::vector<float>> scales(MAX_BLOCKS + 1), shifts(MAX_BLOCKS + 1);
  auto fold_bn = [&](int idx, int ch) -> int {
    const float *g, *b, *mu, *var;
    const std::string base = "batch_normalization_" + std::to_string(idx + 1) + "/";
    int r;
    if ((r = get(base + "gamma", ch, &g))) return r;
    if ((r = get(base + "beta", ch, &b))) return r;
    if ((r = get(base + "moving_mean", ch, &mu))) return r;
    if ((r = get(base + "moving_variance", ch, &var))) return r;
    scales[idx].resize(ch); shifts[idx].resize(ch);
    for (int c = 0; c < ch; ++c) {
      // batchnorm/Rsqrt, mul, mul_1, mul_2, sub, add_1 of the reference graph, eps 1e-3
      const double s = static_cast<double>(g[c]) / std::sqrt(static_cast<double>(var[c]) + 1e-3);
      scales[idx][c] = static_cast<float>(s);
      shifts[idx][c] = static_cast<float>(static_cast<double>(b[c]) - static_cast<double>(mu[c]) * s);
    }
    o_scale[idx] = push(scales[idx].data(), ch); push_pad();
    o_shift[idx] = push(shifts[idx].data(), ch); push_pad();
    return KWS_OK;
  };
  if ((rc = fold_bn(0, m.c0))) return rc;
  std::vector<std::vector<float>> pw_host(m.n_blocks), dw_host(m.n_blocks);
  for (int i = 0; i < m.n_blocks; ++i) {
    const LayerDesc& d = m.layers[i];
    if ((rc = get("depthwise_conv2d_" + std::to_string(i + 1) + "/depthwise_kernel", 3LL * d.cin, &p))) return rc;
    dw_host[i].assign(p, p + 3 * d.cin);
    o_dw[i] = push(p, 3 * d.cin); push_pad();                  // [1,3,C,1] -> [3,C]
    if ((rc = get("conv1d_" + std::to_string(i + 2) + "/kernel", 1LL * d.cin * d.cout, &p))) return rc;
    pw_host[i].assign(p, p + static_cast<size_t>(d.cin) * d.cout);
    o_pw[i] = push(p, static_cast<size_t>(d.cin) * d.cout); push_pad();
    if ((rc = fold_bn(i + 1, d.cout))) return rc;
  }
  const size_t d1_numel = m.head_kind == 1 ? static_cast<size_t>(m.c_last) * m.hidden : static_cast<size_t>(m.t_last) * m.c_last * m.t_last;
  if ((rc = get("dense_1/kernel", static_cast<int64_t>(d1_numel), &p))) return rc;
  size_t o_d1 = push(p, d1_numel); push_pad();
  std::vector<float> bias(m.t_last, 0.f);
  if (m.dense1_bias) {
    if ((rc = get("dense_1/bias", m.t_last, &p))) return rc;
    bias.assign(p, p + m.t_last);
  }
  size_t o_b1 = push(bias.data(), m.t_last); push_pad();
  if ((rc = get("dense_2/kernel", 1LL * feat * m.classes, &p))) return rc;
  size_t o_d2 = push(p, static_cast<size_t>(feat) * m.classes); push_pad();
  std::vector<float> shift0_tc(m.c0_tc, 0.0f);                  // padded filters: zero weights, zero shift -> ReLU6(0) = 0
  std::copy(shifts[0].begin(), shifts[0].end(), shift0_tc.begin());
  size_t o_shift0_tc = push(shift0_tc.data(), m.c0_tc); push_pad();

  KWS_CUDA(h, cudaMalloc(&m.blob, host.size() * sizeof(float)));
  KWS_CUDA(h, cudaMemcpy(m.blob, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
  m.w_conv1 = m.blob + o_conv1;
  for (int i = 0; i <= m.n_blocks; ++i) { m.bn_scale[i] = m.blob + o_scale[i]; m.bn_shift[i] = m.blob + o_shift[i]; }
  for (int i = 0; i < m.n_blocks; ++i) { m.w_dw[i] = m.blob + o_dw[i]; m.w_pw[i] = m.blob + o_pw[i]; }
  m.tc_shift0 = m.blob + o_shift0_tc;
  m.w_d1 = m.blob + o_d1; m.b_d1 = m.blob + o_b1; m.w_d2 = m.blob + o_d2;

  // tensor-core operand images (fp16, pre-swizzled, BN scale folded in); the BN shift stays fp32
  if ((rc = model_build_tc(h, m, pw_host, conv1_host, dw_host, scales))) return rc;
  m.loaded = true;
  return KWS_OK;
}

// ---------------------------------------------------------------------------------------------
// fp32 forward: conv1d_1 (implicit im2col GEMM) -> 11 / 13 x [depthwise prologue + pointwise GEMM +
// BN + ReLU6] -> head.  Activations are channels-last fp32 [rows, T, C], ping-ponged in the
// handle's workspace; rows = clips_in_chunk * n_views, views of a clip adjacent.
// ---------------------------------------------------------------------------------------------
int launch_forward_f32(kws_handle* h, Model& m, const float* wav, int B, const ViewTable& vt,
                       float* probs_mean, int32_t* argmax, cudaStream_t st, int dbg_layer, float* dbg_out) {
  const int V = vt.n;
  const int clips_per_chunk = std::max(1, h->max_rows / V);
  const size_t need = static_cast<size_t>(clips_per_chunk) * V * m.max_act_elems * sizeof(float);
  for (int i = 0; i < 2; ++i) {
    size_t cur = h->act_bytes;
    if (cur < need) {
      if (h->act[i]) cudaFree(h->act[i]);
      h->act[i] = nullptr;
      KWS_CUDA(h, cudaMalloc(&h->act[i], need));
    }
  }
  if (h->act_bytes < need) h->act_bytes = need;

  for (int b0 = 0; b0 < B; b0 += clips_per_chunk) {
    const int nb = std::min(clips_per_chunk, B - b0);
    const int rows = nb * V;
    float* cur = static_cast<float*>(h->act[0]);
    float* nxt = static_cast<float*>(h->act[1]);
    {
      LoadSliceConv1 a{wav + static_cast<size_t>(b0) * L, m.t0, V, vt};
      EpiBnRelu6 e{cur, m.bn_scale[0], m.bn_shift[0]};
      KWS_T0(h, KC_CONV1, st);
      launch_gemm_f32(a, m.w_conv1, rows * m.t0, m.c0, 120, e, st);
      KWS_T1(h, st);
      KWS_LAUNCH_CHECK(h);
    }
    if (dbg_layer == 0) return launch_to_float(h, cur, false, dbg_out, static_cast<size_t>(rows) * m.t0 * m.c0, st);
    for (int i = 0; i < m.n_blocks; ++i) {
      const LayerDesc& d = m.layers[i];
      LoadDepthwise a{cur, m.w_dw[i], d.t_in, d.t_out, d.cin, d.stride, d.pad_left};
      EpiBnRelu6 e{nxt, m.bn_scale[i + 1], m.bn_shift[i + 1]};
      KWS_T0(h, KC_BLOCKS, st);
      launch_gemm_f32(a, m.w_pw[i], rows * d.t_out, d.cout, d.cin, e, st);
      KWS_T1(h, st);
      KWS_LAUNCH_CHECK(h);
      std::swap(cur, nxt);
      if (dbg_layer == i + 1)
        return launch_to_float(h, cur, false, dbg_out, static_cast<size_t>(rows) * d.t_out * d.cout, st);
    }
    int rc = launch_head(h, m, cur, /*act_half=*/false, nb, V,
                         probs_mean ? probs_mean + static_cast<size_t>(b0) * m.classes : nullptr,
                         argmax ? argmax + b0 : nullptr, st);
    if (rc) return rc;
  }
  return KWS_OK;
}

}  // namespace kws
