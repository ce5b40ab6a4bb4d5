// K8 -- per-clip driver arithmetic of the inference scripts, one thread per clip
// (rows are tiny: 12..32 values; HBM-bound, trivially).
//   convert : 32 -> 12 class map with max over the unknown group, re-softmax without max
//             subtraction, uint8 = trunc(p*255)      (convert_from_see_v3_bugfix.py:61-110,
//             freeze_graph_32_classes.py:55-69)
//   select  : label = first argmax of the uint8 row, keep = !(float32(max)/255 < thresh)
//             (create_pseudo_with_thresh.py:17-18,40-43)
//   vote    : N-way majority with first-seen tie rule and fallback to submission 0
//             (majority_vote.py:26-56); unanimity (REPR_106_pseudo.py:12) = clear with M=3,min=3
#include "common.cuh"

namespace kws {

namespace {

constexpr int SEL_THREADS = 256;
constexpr int MAX_CIN = 64, MAX_COUT = 32, MAX_SUBS = 16;

struct ClassMap { int n_in; int n_out; int8_t map[MAX_CIN]; };

__global__ void __launch_bounds__(SEL_THREADS)
convert_kernel(const float* __restrict__ probs, int B, ClassMap cm, float* __restrict__ probs_out,
               uint8_t* __restrict__ probs_u8) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float o[MAX_COUT];
#pragma unroll
  for (int j = 0; j < MAX_COUT; ++j) o[j] = -INFINITY;
  const float* p = probs + static_cast<size_t>(b) * cm.n_in;
  for (int c = 0; c < cm.n_in; ++c) {
    const int j = cm.map[c];
    const float v = p[c];
#pragma unroll
    for (int q = 0; q < MAX_COUT; ++q)
      if (q == j) o[q] = fmaxf(o[q], v);                 // np.float32(unknown_probs).max(axis=0)
  }
  float s = 0.0f;
#pragma unroll
  for (int j = 0; j < MAX_COUT; ++j)
    if (j < cm.n_out) { o[j] = expf(o[j]); s = __fadd_rn(s, o[j]); }   // softmax(): exp(x)/sum, no max shift
#pragma unroll
  for (int j = 0; j < MAX_COUT; ++j)
    if (j < cm.n_out) {
      const float q = __fdiv_rn(o[j], s);
      if (probs_out) probs_out[static_cast<size_t>(b) * cm.n_out + j] = q;
      if (probs_u8) probs_u8[static_cast<size_t>(b) * cm.n_out + j] =
          static_cast<uint8_t>(static_cast<int>(__fmul_rn(q, 255.0f)));   // norm_probs[...] = see_probs*255 (trunc)
    }
}

__global__ void __launch_bounds__(SEL_THREADS)
select_kernel(const uint8_t* __restrict__ probs_u8, int B, int C, int min_keep_u8,
              int32_t* __restrict__ label, uint8_t* __restrict__ keep) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const uint8_t* p = probs_u8 + static_cast<size_t>(b) * C;
  int best = -1, idx = 0;
  for (int c = 0; c < C; ++c) {
    const int v = p[c];
    if (v > best) { best = v; idx = c; }                 // argmax: first index wins ties
  }
  if (label) label[b] = idx;
  if (keep) keep[b] = best >= min_keep_u8 ? 1 : 0;
}

__global__ void __launch_bounds__(SEL_THREADS)
vote_kernel(const int32_t* __restrict__ labels, int M, int B, int min_count,
            int32_t* __restrict__ voted, uint8_t* __restrict__ clear) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int l[MAX_SUBS];
#pragma unroll
  for (int m = 0; m < MAX_SUBS; ++m) l[m] = m < M ? labels[static_cast<size_t>(m) * B + b] : -1;
  // dict insertion order + max(): among the labels with the highest count, the one whose first
  // occurrence is earliest wins -> scan first occurrences in order, strict '>' keeps the earliest.
  int best_label = l[0], best_count = 0;
#pragma unroll
  for (int m = 0; m < MAX_SUBS; ++m) {
    if (m >= M) continue;
    bool first = true;
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < MAX_SUBS; ++q) {
      if (q >= M) continue;
      if (l[q] == l[m]) { if (q < m) first = false; ++cnt; }
    }
    if (first && cnt > best_count) { best_count = cnt; best_label = l[m]; }
  }
  const bool ok = best_count >= min_count;
  if (voted) voted[b] = ok ? best_label : l[0];          // fallback: subs[0] (majority_vote.py:48)
  if (clear) clear[b] = ok ? 1 : 0;
}

}  // namespace

int launch_convert(kws_handle* h, const float* probs, int B, int C_in, const int32_t* class_map_h,
                   int C_out, float* probs_out, uint8_t* probs_u8, cudaStream_t st) {
  if (C_in <= 0 || C_in > MAX_CIN || C_out <= 0 || C_out > MAX_COUT)
    return fail(h, KWS_EINVAL, "class counts out of range");
  ClassMap cm;
  cm.n_in = C_in; cm.n_out = C_out;
  std::vector<int> hits(C_out, 0);
  for (int c = 0; c < C_in; ++c) {
    if (class_map_h[c] < 0 || class_map_h[c] >= C_out) return fail(h, KWS_EINVAL, "class_map entry out of range");
    cm.map[c] = static_cast<int8_t>(class_map_h[c]);
    hits[class_map_h[c]]++;
  }
  for (int j = 0; j < C_out; ++j)
    if (!hits[j]) return fail(h, KWS_EINVAL, "class_map leaves an output class without a source");
  if (B == 0) return KWS_OK;
  convert_kernel<<<(B + SEL_THREADS - 1) / SEL_THREADS, SEL_THREADS, 0, st>>>(probs, B, cm, probs_out, probs_u8);
  KWS_LAUNCH_CHECK(h);
  return KWS_OK;
}

int launch_select(kws_handle* h, const uint8_t* probs_u8, int B, int C, double thresh,
                  int32_t* label, uint8_t* keep, cudaStream_t st) {
  if (C <= 0) return fail(h, KWS_EINVAL, "C must be positive");
  // keep iff !(float32(v)/255 < thresh): monotone in v, so resolve the comparison once on the
  // host with the reference's own arithmetic and compare integers on the device (bit-exact).
  int min_keep = 256;
  for (int v = 0; v <= 255; ++v) {
    const float q = static_cast<float>(v) / 255.0f;
    if (!(static_cast<double>(q) < thresh)) { min_keep = v; break; }
  }
  if (B == 0) return KWS_OK;
  select_kernel<<<(B + SEL_THREADS - 1) / SEL_THREADS, SEL_THREADS, 0, st>>>(probs_u8, B, C, min_keep, label, keep);
  KWS_LAUNCH_CHECK(h);
  return KWS_OK;
}

int launch_vote(kws_handle* h, const int32_t* labels, int M, int B, int min_count, int32_t* voted,
                uint8_t* clear, cudaStream_t st) {
  if (M <= 0 || M > MAX_SUBS) return fail(h, KWS_EINVAL, "vote supports 1..16 submissions");
  if (B == 0) return KWS_OK;
  vote_kernel<<<(B + SEL_THREADS - 1) / SEL_THREADS, SEL_THREADS, 0, st>>>(labels, M, B, min_count, voted, clear);
  KWS_LAUNCH_CHECK(h);
  return KWS_OK;
}

}  // namespace kws
