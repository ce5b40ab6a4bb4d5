// Tensor-core tier of the Depthwise1D network forward (reference model.py:34-52,67-76,805-817).
//
// One warp-specialised, persistent kernel template serves both GEMM-shaped layer types:
//   * K5 slice_conv1 : overlapping_time_slice_stack (k40 s20 SAME) + Conv1D(C0,k3,s2) + BN + ReLU6.
//                      The 3 overlapping patches of a row cover 80 consecutive samples, so the
//                      layer is a K=80 contraction against taps pre-summed on the host
//                      (W80[u] = sum_f W[f, u-20f]); the TTA view (np.roll + gain,
//                      make_submission.py:126-130) is applied while the waveform window is staged.
//   * K6 dw_pw_block : depthwise k3 FIR (CUDA cores) as the A-operand producer of the pointwise
//                      GEMM, BN + ReLU6 in the epilogue.
// Roles of a dw_pw block (608 threads, 1 CTA / SM, static round-robin over 128-row tiles):
//   warps 0-7   epilogue   : tcgen05.ld accumulator -> +shift -> ReLU6 -> fp16 -> swizzled smem box -> TMA store
//                            (two warps per TMEM lane quarter on alternate 64-column chunks)
//   warp  8     MMA        : all lanes walk the loop, one elected lane issues; one predicated PTX sequence
//                            per K slab (4 x tcgen05.mma + the commits), A,B from swizzled smem, D in TMEM
//   warp  9     B loader   : cp.async.bulk of pre-swizzled fp16 weight blocks (resident when they fit)
//   warp  10    raw loader : cp.async.bulk.tensor (TMA) of the previous activation's rows, one
//                            [rows x 64 channel] box per K slab, several slabs ahead of the producers
//   warps 11-18 A producers: 256 threads fill the A ring in order; each thread slides a 3-row window over
//                            4 consecutive output rows of one 8-channel chunk (packed half2 FMAs: the FIR
//                            outputs are complete before it waits for the A slot) and stores the
//                            UMMA-swizzled A slab; one mbarrier arrival per warp
// Pipelines: raw ring (TMA <-> producers), A ring (producers <-> MMA), B ring (loader <-> MMA),
// accumulator stages in TMEM (MMA <-> epilogue), all on mbarriers; tcgen05.commit releases smem
// slots / publishes accumulators.  Layers with more than 256 columns run with one accumulator stage,
// or -- when every K slab of a tile fits the A ring -- with two-pass accumulation (GemmParams::two_pass).
// Operands are fp16 (same tensor rate as bf16, 3 more mantissa bits; ReLU6 bounds activations to
// [0,6] so the range is safe), accumulation is fp32 in TMEM.  The BatchNorm scale is folded into
// the fp16 weights at load time, the shift stays fp32 in the epilogue.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda.h>

#include "common.cuh"
#include "tc_common.cuh"

// Profiling build (KWS_PROFILE_BUILD=1 python -m speech_recognition_b200.build --force): compiles the
// KWS_KNOCKOUT switches and the KWS_TRACE event log into the block kernel.  The production build
// carries neither (the trace call sites alone cost ~20 instructions per K slab on the producers' chain).
#ifndef KWS_PROFILE_BUILD
#define KWS_PROFILE_BUILD 0
#endif

namespace kws {

using namespace tc;

namespace {

constexpr bool kProfile = KWS_PROFILE_BUILD != 0;

// Warp roles depend on the layer type: the dw_pw blocks are fed by 8 depthwise producer warps and
// drained by 4 epilogue warps; conv1d_1 is bound by its epilogue (one pass per TTA view of a
// group over a K = 80 GEMM), so it gets 8 epilogue warps (two per TMEM lane quarter, alternating
// 64-column chunks), 6 producer warps and no raw-loader warp.
template <int MODE> struct Roles {
  static constexpr int EPI_WARPS = 8;
  static constexpr int MMA_WARP = EPI_WARPS;
  static constexpr int LOAD_WARP = EPI_WARPS + 1;
  static constexpr int RAW_WARP = MODE == 0 ? -1 : EPI_WARPS + 2;
  static constexpr int PROD_WARP0 = MODE == 0 ? EPI_WARPS + 2 : EPI_WARPS + 3;
  static constexpr int PROD_WARPS = MODE == 0 ? 6 : 8;
  [[maybe_unused]] static constexpr int PROD_THREADS = PROD_WARPS * 32;           // 192 / 256
  static constexpr int THREADS = 32 * (PROD_WARP0 + PROD_WARPS); // 512 / 608
};
constexpr int NUM_PROD_THREADS = Roles<1>::PROD_THREADS;         // dw_pw: 256
constexpr int CONV1_PROD_THREADS = Roles<0>::PROD_THREADS;       // conv1: 192
constexpr int RUN_ROWS = TILE_M * 8 / NUM_PROD_THREADS;           // dw_pw: consecutive output rows per producer thread (4)
constexpr int MAX_STAGES = 8;                                    // A / raw rings
constexpr int MAX_B_BLOCKS = 16;                                 // B ring (K slabs x N halves)
constexpr int TMEM_COLS = 512;
constexpr int SMEM_LIMIT = 227 * 1024;

constexpr int CONV1_K = 80;                                      // samples per output row
constexpr int CONV1_ROW_HOP = 40;
constexpr int CONV1_WIN = CONV1_ROW_HOP * (TILE_M - 1) + CONV1_K;   // 5160 staged samples per tile
constexpr uint32_t CONV1_WIN_BYTES = ((CONV1_WIN + 8) * 2 + 15) & ~15;   // one fp16 window buffer
constexpr int conv1_quads(int threads) { return ((CONV1_WIN + 6) / 4 + threads - 1) / threads; }   // float4 loads per producer thread and tile


// TTA views that share a roll shift differ only by the gain, and conv1d_1 is linear and bias-free
// (model.py:68-74), so they share the A tile and the accumulator: gain * conv(x) == conv(gain * x).
struct ViewGroups {
  int n_groups;
  int shift[KWS_MAX_VIEWS];       // roll shift of the group, already reduced to [0, L)
  int start[KWS_MAX_VIEWS + 1];   // members of group g = [start[g], start[g+1])
  int view[KWS_MAX_VIEWS];        // member -> view index (output row block)
  float gain[KWS_MAX_VIEWS];      // member -> gain
};

struct alignas(64) GemmParams {
  CUtensorMap tmap_in;      // dw_pw: previous activation [rows_in, cin] fp16, box [box_rows, 64]
  CUtensorMap tmap_out;     // output activation, box [32 rows, 64 ch], 128-byte swizzle; dw_pw: 2-D
                            // [rows_out, cout]; conv1: 3-D [clip-views, t_out, cout] (rows clipped per view)
  CUtensorMap tmap_tail;    // dw_pw with ncta % 64 == 32: box [32 rows, 32 ch], no swizzle, for the last 32 columns
  // conv1 A side
  const float* wav;         // waveforms [B, 16000] fp32
  ViewGroups vg;            // TTA views grouped by shift: one staged window + one MMA per (clip, group),
  int n_views;              // one epilogue pass per member view (the gain is applied to the accumulator)
  // dw_pw A side
  const __half* dw_h;       // depthwise taps [3][cin] fp16
  const __half* in_act;     // the previous activation itself (L2 prefetch of the rows of tiles ahead)
  long long rows_in;
  int prefetch_tiles;       // how many of this CTA's tiles ahead the raw loader prefetches into L2 (0 = off)
  unsigned long long* trace;   // KWS_TRACE=<block index> (profiling only): per-role event log of CTA 0, see trace_ev
  int knockout;             // KWS_KNOCKOUT (profiling only, results become wrong): 1 no TMA stores, 2 no FIR,
                            // 4 no MMAs, 8 no epilogue math, 16 no raw loads, 32 no weight loads
  // B side: pre-swizzled fp16 blocks of n_inst rows x 128 B, block j = (kb, nh) = (j / n_halves, j % n_halves)
  const uint8_t* w_img;
  // epilogue
  const float* shift;       // beta - mean * scale (the scale lives in the weights)
  int block_index;          // host-side only: which block this launch is (timing class)
  const __half* out_act;    // host-side only: the output activation (for the tail tensor map)
  long long out_rows;
  // shapes
  int cin, cout, stride, pad_left, t_in, t_out;
  int rows_out;             // valid output rows
  int num_tiles;
  int tiles_per_group;      // conv1: tiles per clip-view (4); dw_pw: unused
  int num_kb;               // K slabs
  int last_ksteps;          // K=16 steps in the last slab (4, or 1 for conv1)
  int a_stages, b_stages, acc_stages, b_resident, raw_stages;
  int n_split, ncta;        // the cout columns are split over n_split adjacent CTAs, ncta = cout / n_split each:
                            // CTA c works on columns [(c % n_split) * ncta, +ncta) of tiles c / n_split, + grid / n_split, ...
                            // so that its slice of the weights stays resident and two accumulators fit in TMEM
  int n_inst, n_halves;     // ncta = n_inst * n_halves, n_inst <= 256
  int two_pass;             // dw_pw, n_halves == 2: the A ring holds every K slab of a tile and the MMAs run half by half
                            // (all slabs for columns [0, n_inst), then all slabs for [n_inst, 2 n_inst)), each half into
                            // its own TMEM region with its own full / empty barrier: the epilogue of one half overlaps the
                            // MMAs of the other half / the next tile although two whole accumulators do not fit in TMEM
  int out_bufs;             // store boxes per epilogue warp (2, or 1 when shared memory is short)
  int a_stage_bytes;        // dw_pw: 16 KB; conv1: 32 KB (both slabs of a tile)
  int raw_stage_bytes, box_rows, n_boxes;
  unsigned long long t_out_magic;   // ceil(2^40 / t_out)
};

constexpr int OUT_STAGE_BYTES = 32 * ROW_BYTES;                 // one warp's [32 rows x 64 ch] store box

struct SmemLayout {
  uint32_t a_off, b_off, out_off, raw_off, aux_off, bar_off, total;
};

__host__ __device__ inline SmemLayout smem_layout(const GemmParams& p, bool conv1) {
  SmemLayout s;
  uint32_t o = 0;
  s.a_off = o; o += static_cast<uint32_t>(p.a_stages) * p.a_stage_bytes;
  s.b_off = o; o += static_cast<uint32_t>(p.b_stages) * p.n_inst * ROW_BYTES;
  s.out_off = o; o += static_cast<uint32_t>((conv1 ? Roles<0>::EPI_WARPS : Roles<1>::EPI_WARPS) * p.out_bufs * OUT_STAGE_BYTES);
  s.raw_off = o; o += static_cast<uint32_t>(p.raw_stages) * p.raw_stage_bytes;
  s.aux_off = o;
  // aux: shift[cout] fp32, then (dw_pw) taps [3*cin] fp16 + row metadata [2 parities][128] u32
  //      or (conv1) the staged fp16 waveform window
  o += static_cast<uint32_t>(p.ncta) * 4u;
  o += conv1 ? 2u * CONV1_WIN_BYTES
             : (static_cast<uint32_t>(3 * p.cin * 2 + 15) & ~15u) + 2 * TILE_M * 4u;
  o = (o + 15u) & ~15u;
  s.bar_off = o; o += (4 * MAX_STAGES + 2 * MAX_B_BLOCKS + 4) * 8 + 16;
  s.total = o + 1024;        // slack for the manual 1024-byte alignment of the base
  return s;
}

// ------------------------------------------------------------------------------------------------
// A producers
// ------------------------------------------------------------------------------------------------

// Row metadata of a 128-row output tile (dw_pw), one word per output row:
//   bits  0-15 : row of tap 0 inside the raw box (relative to the box's first row)
//   bits 16-18 : validity of taps 0..2 (TF 'SAME' zero padding)
//   bit  19    : the previous output row belongs to the same clip-view (the window can slide)
__device__ __forceinline__ int div_t_out(const GemmParams& p, int m) {   // m / t_out, exact for m * t_out < 2^40
  return static_cast<int>((static_cast<unsigned long long>(m) * p.t_out_magic) >> 40);
}
__device__ __forceinline__ uint32_t row_meta(const GemmParams& p, int stride, int tile, int r) {
  const int m0 = tile * TILE_M;
  const int v0 = div_t_out(p, m0), t0 = m0 - v0 * p.t_out;
  const int lo = v0 * p.t_in + t0 * stride - p.pad_left;      // first row of the raw box
  int m = m0 + r;
  const bool valid = m < p.rows_out;
  if (!valid) m = p.rows_out - 1;                              // finite data; the row is never stored
  const int v = div_t_out(p, m), t = m - v * p.t_out;
  const int ti0 = t * stride - p.pad_left;
  const int rel = v * p.t_in + ti0 - lo;
  uint32_t mask = 0;
  if (ti0 >= 0) mask |= 1u;
  if (ti0 + 1 < p.t_in) mask |= 2u;
  if (ti0 + 2 < p.t_in) mask |= 4u;
  const uint32_t cont = (valid && t > 0) ? 1u : 0u;
  return static_cast<uint32_t>(rel) | (mask << 16) | (cont << 19);
}

__device__ __forceinline__ uint4 lds128(const uint8_t* p) { return *reinterpret_cast<const uint4*>(p); }

// Depthwise k = 3 FIR of one 16-byte chunk (8 channels): fp16 inputs and taps, fp32 accumulation (the mixed-precision
// FMA of sm_100: fma.rn.f32.f16 = FHFMA, operands taken straight from the packed halves), ONE rounding to the fp16 A
// operand of the pointwise GEMM.  (Until r02 this was a packed-half chain hmul2 / hfma2 / hfma2 = three fp16
// roundings; KWS_FIR_FP16=1 at build time restores it for A/B timing.)
__device__ __forceinline__ float fhfma(uint32_t a, uint32_t b, float c, bool hi) {
  float d;
  if (hi) asm("{\n\t.reg .b16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\tfma.rn.f32.f16 %0, ah, bh, %3;\n\t}"
              : "=f"(d) : "r"(a), "r"(b), "f"(c));
  else asm("{\n\t.reg .b16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\tfma.rn.f32.f16 %0, al, bl, %3;\n\t}"
           : "=f"(d) : "r"(a), "r"(b), "f"(c));
  return d;
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint4 fir3(const uint4& x0, const uint4& x1, const uint4& x2, const uint4& k0,
                                      const uint4& k1, const uint4& k2) {
  uint4 o;
#ifdef KWS_FIR_FP16
  const __half2* a = reinterpret_cast<const __half2*>(&x0);
  const __half2* b = reinterpret_cast<const __half2*>(&x1);
  const __half2* c = reinterpret_cast<const __half2*>(&x2);
  const __half2* ka = reinterpret_cast<const __half2*>(&k0);
  const __half2* kb = reinterpret_cast<const __half2*>(&k1);
  const __half2* kc = reinterpret_cast<const __half2*>(&k2);
  __half2* r = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int q = 0; q < 4; ++q) r[q] = __hfma2(c[q], kc[q], __hfma2(b[q], kb[q], __hmul2(a[q], ka[q])));
#else
  const uint32_t* a = reinterpret_cast<const uint32_t*>(&x0);
  const uint32_t* b = reinterpret_cast<const uint32_t*>(&x1);
  const uint32_t* c = reinterpret_cast<const uint32_t*>(&x2);
  const uint32_t* ka = reinterpret_cast<const uint32_t*>(&k0);
  const uint32_t* kb = reinterpret_cast<const uint32_t*>(&k1);
  const uint32_t* kc = reinterpret_cast<const uint32_t*>(&k2);
  uint32_t* r = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float lo = fhfma(c[q], kc[q], fhfma(b[q], kb[q], fhfma(a[q], ka[q], 0.0f, false), false), false);
    const float hi = fhfma(c[q], kc[q], fhfma(b[q], kb[q], fhfma(a[q], ka[q], 0.0f, true), true), true);
    r[q] = pack_f16x2(lo, hi);
  }
#endif
  return o;
}

// One K slab (64 channels) of a 128-row tile by the 256 producer threads: thread = (16-byte channel
// chunk c, run of 4 consecutive output rows).  raw = [box rows][64 ch] fp16 (dense 128-byte rows).
// The run is decoded once per tile (RunDesc); per slab the thread computes its 4 output chunks in
// registers (fir_run) and only then waits for the A slot and stores them (store_run), so the raw
// loads and the FIRs of slab k overlap the MMA that still reads the slot.
// Fast path (the 4 rows continue one clip-view): every raw row of the run is loaded up front
// (6 / 9 independent LDS.128), then the FIRs run from registers.  A run that crosses into the next
// clip-view reloads its window at the crossing.
struct RunDesc {
  uint32_t mt[RUN_ROWS];    // row metadata words of the run
  uint32_t base_off;        // byte offset of (first raw row, this thread's chunk) inside a raw stage
  uint32_t dst_off[RUN_ROWS];   // swizzled byte offsets of the 4 output chunks inside an A slab
  bool all_cont;
};
__device__ __forceinline__ RunDesc decode_run(const uint32_t* meta, int tg) {
  RunDesc d;
  const int c = tg & 7, g = tg >> 3;
  const uint4 m4 = *reinterpret_cast<const uint4*>(meta + RUN_ROWS * g);
  d.mt[0] = m4.x; d.mt[1] = m4.y; d.mt[2] = m4.z; d.mt[3] = m4.w;
  d.all_cont = (m4.y & m4.z & m4.w & (1u << 19)) != 0;
  d.base_off = (m4.x & 0xffffu) * ROW_BYTES + c * 16;
#pragma unroll
  for (int i = 0; i < RUN_ROWS; ++i) d.dst_off[i] = swz_off(RUN_ROWS * g + i, c);
  return d;
}
template <int STRIDE>
__device__ __forceinline__ void fir_run(const uint8_t* raw, const RunDesc& d, const uint4& k0, const uint4& k1,
                                        const uint4& k2, int c, uint4 (&o)[RUN_ROWS]) {
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  if (d.all_cont) {
    const uint8_t* base = raw + d.base_off;
    if (STRIDE == 1) {
      uint4 x[RUN_ROWS + 2];
#pragma unroll
      for (int i = 0; i < RUN_ROWS + 2; ++i) x[i] = lds128(base + i * ROW_BYTES);
#pragma unroll
      for (int i = 0; i < RUN_ROWS; ++i) o[i] = fir3(x[i], x[i + 1], x[i + 2], k0, k1, k2);
    } else {
      uint4 x[2 * RUN_ROWS + 1];
#pragma unroll
      for (int i = 0; i < 2 * RUN_ROWS + 1; ++i) x[i] = lds128(base + i * ROW_BYTES);
      if (!((d.mt[0] >> 16) & 1u)) x[0] = zero;                  // left 'SAME' pad: only a run's first row can be t = 0
      if (!((d.mt[RUN_ROWS - 1] >> 18) & 1u)) x[2 * RUN_ROWS] = zero;   // right pad: only its last row can be t = T_out - 1
#pragma unroll
      for (int i = 0; i < RUN_ROWS; ++i) o[i] = fir3(x[2 * i], x[2 * i + 1], x[2 * i + 2], k0, k1, k2);
    }
    return;
  }
  uint4 w0 = zero, w1 = zero, w2 = zero;
#pragma unroll
  for (int i = 0; i < RUN_ROWS; ++i) {
    const uint8_t* src = raw + (d.mt[i] & 0xffffu) * ROW_BYTES + c * 16;
    const bool cont = (i > 0) && ((d.mt[i] >> 19) & 1u);
    if (STRIDE == 1) {
      // VALID convolution: every tap of a valid row is inside its clip
      if (cont) { w0 = w1; w1 = w2; }
      else { w0 = lds128(src); w1 = lds128(src + ROW_BYTES); }
      w2 = lds128(src + 2 * ROW_BYTES);
    } else {
      if (cont) w0 = w2;
      else w0 = ((d.mt[i] >> 16) & 1u) ? lds128(src) : zero;
      w1 = lds128(src + ROW_BYTES);
      w2 = ((d.mt[i] >> 18) & 1u) ? lds128(src + 2 * ROW_BYTES) : zero;
    }
    o[i] = fir3(w0, w1, w2, k0, k1, k2);
  }
}

// conv1 producer.  A tile is 128 output rows of one (clip, view group): row r needs samples
// [40 r - 10, 40 r + 70) of the rolled clip, so the tile needs one contiguous window of
// 40 (rows - 1) + 80 samples.  The window is fetched with source-aligned float4 loads (L % 4 == 0,
// so a quad never straddles the np.roll wrap-around), one tile AHEAD of its use (registers), then
// written to shared memory as fp16 at its rolled position, and finally copied as 16-byte chunks
// into the swizzled slabs: slab 0 = first 64 samples of a row, slab 1 = last 16.
struct Conv1Tile {
  int ps_lo;        // first valid sample position of the window in rolled coordinates (>= 0)
  int n;            // valid samples [ps_lo, ps_lo + n)
  int win_off;      // ps_lo - p_start: where sample ps_lo lands in the window (10 for the first tile)
  int src_al;       // 4-aligned source index of the first quad
  int mis;          // source misalignment: quad element e of quad i is sample ps_lo + 4 i + e - mis
  int rows;         // valid rows of the tile
};

template <int NT>                                                // NT = producer threads
struct Conv1ProducerT {
  static constexpr int QUADS = conv1_quads(NT);
  // window of `rows` output rows starting at conv1d_1 row t_start of a clip rolled by sm samples
  __device__ __forceinline__ static Conv1Tile make(int sm, int t_start, int rows) {
    Conv1Tile t;
    t.rows = rows;
    const int p_start = CONV1_ROW_HOP * t_start - 10;         // patch stack pads 10 samples on the left
    const int p_end = min(L, p_start + CONV1_ROW_HOP * (rows - 1) + CONV1_K);
    t.ps_lo = max(p_start, 0);
    t.n = p_end - t.ps_lo;
    t.win_off = t.ps_lo - p_start;
    int src = t.ps_lo - sm; if (src < 0) src += L;              // np.roll(x, shift)[ps] = x[(ps - shift) mod L]
    t.src_al = src & ~3;
    t.mis = src & 3;
    return t;
  }
  __device__ __forceinline__ static Conv1Tile describe(const GemmParams& p, int tile, const float** x) {
    const int unit = tile / p.tiles_per_group, jb = tile - unit * p.tiles_per_group;
    const int b = unit / p.vg.n_groups, g = unit - b * p.vg.n_groups;
    *x = p.wav + static_cast<size_t>(b) * L;
    return make(p.vg.shift[g], jb * TILE_M, min(TILE_M, p.t_out - jb * TILE_M));
  }
  __device__ __forceinline__ static void load(const Conv1Tile& t, const float* x, int ptid, float4 (&q)[QUADS]) {
    const int n_quads = (t.n + t.mis + 3) >> 2;
#pragma unroll
    for (int k = 0; k < QUADS; ++k) {
      const int i = ptid + k * NT;
      if (i < n_quads) {
        int s4 = t.src_al + 4 * i; if (s4 >= L) s4 -= L;
        q[k] = __ldg(reinterpret_cast<const float4*>(x + s4));
      }
    }
  }
  __device__ __forceinline__ static void store(const Conv1Tile& t, int ptid, const float4 (&q)[QUADS], __half* s_win) {
    if (ptid < t.win_off) s_win[ptid] = __float2half_rn(0.0f);       // 'SAME' left pad of the patch stack
    const int n_quads = (t.n + t.mis + 3) >> 2;
#pragma unroll
    for (int k = 0; k < QUADS; ++k) {
      const int i = ptid + k * NT;
      if (i < n_quads) {
        const int e0 = 4 * i - t.mis;                               // sample offset of element 0 from ps_lo
        __half* dst = s_win + t.win_off + e0;
        const float v[4] = {q[k].x, q[k].y, q[k].z, q[k].w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (e0 + e >= 0 && e0 + e < t.n) dst[e] = __float2half_rn(v[e]);
      }
    }
  }
  __device__ __forceinline__ static void fill_slabs(uint8_t* stage, int ptid, const __half* s_win, int rows) {
    // rows x 10 chunks (8 in slab 0, 2 in slab 1); rows beyond `rows` keep stale data and are never stored
    for (int task = ptid; task < rows * 10; task += NT) {
      const int r = task / 10, ch = task - r * 10;
      const uint4 v = *reinterpret_cast<const uint4*>(s_win + CONV1_ROW_HOP * r + 8 * ch);
      uint8_t* slab = stage + (ch < 8 ? 0 : A_SLAB_BYTES);
      *reinterpret_cast<uint4*>(slab + swz_off(r, ch < 8 ? ch : ch - 8)) = v;
    }
  }
};

// ------------------------------------------------------------------------------------------------
// epilogue: 32 accumulator columns of one row -> +shift -> ReLU6 -> 32 fp16 = chunks ch0..ch0+3 of the
// row's 128-byte line in the swizzled store box
// ------------------------------------------------------------------------------------------------
template <bool kGain>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&v)[32], const float* s_shift32, uint8_t* row_base,
                                               int ch0, int sw, float gain) {
  // Every shift is loaded BEFORE the first store: row_base is a byte pointer and may alias anything, so a load placed
  // after a store to it cannot be hoisted by the compiler, and the warp would wait out one shared-memory round trip per
  // 16-byte chunk (r02c profile of the fused kernel's output warps: 0.12 instructions per cycle, short-scoreboard bound).
  float4 sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sh[i] = *reinterpret_cast<const float4*>(s_shift32 + 4 * i);
  const __half2 six = __float2half2_rn(6.0f);
  uint32_t o[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float f0 = __uint_as_float(v[4 * i]), f1 = __uint_as_float(v[4 * i + 1]);
    const float f2 = __uint_as_float(v[4 * i + 2]), f3 = __uint_as_float(v[4 * i + 3]);
    const uint32_t a = kGain ? pack_relu_f16x2(fmaf(f0, gain, sh[i].x), fmaf(f1, gain, sh[i].y)) : pack_relu_f16x2(f0 + sh[i].x, f1 + sh[i].y);
    const uint32_t b = kGain ? pack_relu_f16x2(fmaf(f2, gain, sh[i].z), fmaf(f3, gain, sh[i].w)) : pack_relu_f16x2(f2 + sh[i].z, f3 + sh[i].w);
    const __half2 ha = __hmin2(*reinterpret_cast<const __half2*>(&a), six);
    const __half2 hb = __hmin2(*reinterpret_cast<const __half2*>(&b), six);
    o[2 * i] = *reinterpret_cast<const uint32_t*>(&ha);
    o[2 * i + 1] = *reinterpret_cast<const uint32_t*>(&hb);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *reinterpret_cast<uint4*>(row_base + (((ch0 + q) ^ sw) << 4)) = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
}

// Event log for pipeline analysis (KWS_TRACE=<block index>, tools/gpu_trace.sh, tools/trace_view.py): role r
// owns records [r * TRACE_EVENTS, (r + 1) * TRACE_EVENTS), record = (event << 56 | index << 40 | clock64 & (2^40-1));
// only CTA 0 writes, one thread per role.  p.trace == nullptr (always, unless KWS_TRACE is set) disables it.
constexpr int TRACE_EVENTS = 4096, TRACE_ROLES = 4;
__device__ __forceinline__ void trace_ev(unsigned long long* trace, int role, int& cnt, int ev, int idx) {
  if constexpr (!kProfile) return;
  if (trace != nullptr && blockIdx.x == 0 && cnt < TRACE_EVENTS) {
    const unsigned long long c = static_cast<unsigned long long>(clock64());
    trace[role * TRACE_EVENTS + cnt++] = (static_cast<unsigned long long>(ev) << 56) |
                                         (static_cast<unsigned long long>(idx & 0xffff) << 40) | (c & ((1ull << 40) - 1));
  }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int MODE>   // 0 = conv1, 1 = dw_pw stride 1, 2 = dw_pw stride 2
__global__ void __launch_bounds__(Roles<MODE>::THREADS, 1) tc_gemm_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr bool kConv1 = (MODE == 0);
  constexpr int kStride = (MODE == 2) ? 2 : 1;
  using R = Roles<MODE>;
  constexpr int TC_THREADS = R::THREADS, NUM_EPI_WARPS = R::EPI_WARPS, MMA_WARP = R::MMA_WARP, LOAD_WARP = R::LOAD_WARP,
                RAW_WARP = R::RAW_WARP, PROD_WARP0 = R::PROD_WARP0;
  constexpr int kColGroups = NUM_EPI_WARPS / 4;                  // epilogue warps per TMEM lane quarter
  const SmemLayout lay = smem_layout(p, kConv1);
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_base = smem + lay.a_off;
  uint8_t* b_base = smem + lay.b_off;
  uint8_t* raw_base = smem + lay.raw_off;
  uint8_t* out_base = smem + lay.out_off;
  float* s_shift = reinterpret_cast<float*>(smem + lay.aux_off);
  __half* s_dwh = reinterpret_cast<__half*>(s_shift + p.ncta);                       // dw_pw
  uint32_t* s_meta = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(s_dwh) + ((3 * p.cin * 2 + 15) & ~15));
  __half* s_win = reinterpret_cast<__half*>(s_shift + p.ncta);                       // conv1
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bar_off);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + MAX_STAGES;
  uint64_t* raw_full = bars + 2 * MAX_STAGES;
  uint64_t* raw_empty = bars + 3 * MAX_STAGES;
  uint64_t* b_full = bars + 4 * MAX_STAGES;
  uint64_t* b_empty = b_full + MAX_B_BLOCKS;
  uint64_t* acc_full = b_empty + MAX_B_BLOCKS;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int col0 = (static_cast<int>(blockIdx.x) % p.n_split) * p.ncta;      // this CTA's output columns
  const int tile0 = static_cast<int>(blockIdx.x) / p.n_split;
  const int tstride = static_cast<int>(gridDim.x) / p.n_split;

  // ---- one-time setup ----
  for (int i = tid; i < p.ncta; i += TC_THREADS) s_shift[i] = p.shift[col0 + i];
  if (!kConv1)
    for (int i = tid; i < 3 * p.cin; i += TC_THREADS) s_dwh[i] = p.dw_h[i];
  if (warp == MMA_WARP) {
    if (lane == 0) {
      for (int i = 0; i < MAX_STAGES; ++i) {
        mbar_init(&a_full[i], (kConv1 ? CONV1_PROD_THREADS : NUM_PROD_THREADS) / 32);   // one arrival per warp
        mbar_init(&a_empty[i], 1);
        mbar_init(&raw_full[i], 1);
        mbar_init(&raw_empty[i], NUM_PROD_THREADS / 32);
      }
      for (int i = 0; i < MAX_B_BLOCKS; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], NUM_EPI_WARPS); }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  if (tid == 0) tmem_slot[1] = 0u;
  if (!kConv1 && warp == RAW_WARP && lane == 0) tma_prefetch_desc(&p.tmap_in);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t z = reinterpret_cast<volatile uint32_t*>(tmem_slot)[1];
  const int zi = static_cast<int>(z);
  unsigned long long* const trace = p.trace + z;
  const uint32_t b_block_bytes = static_cast<uint32_t>(p.n_inst) * ROW_BYTES;
  const int blocks_per_tile = p.num_kb * p.n_halves;

  if (warp < NUM_EPI_WARPS) {
    // =========================== epilogue ===========================
    // Each warp owns 32 accumulator rows (TMEM lane quarter warp % 4) and every kColGroups-th 64-column
    // chunk (column group warp / 4).  Per 64 output channels it converts its rows into a
    // private, 128-byte-swizzled [32 x 64] fp16 box in shared memory (conflict-free 16-byte stores)
    // and one lane hands the box to the TMA store engine: full 128-byte lines leave the SM, rows
    // beyond the tensor (or beyond the clip-view for conv1) are clipped by the tensor map.
    int acc = 0; uint32_t acc_phase = 0;
    uint8_t* my_out = out_base + warp * p.out_bufs * OUT_STAGE_BYTES;
    int obuf = 0;
    const int quarter = warp & 3;
    constexpr int c_step = 64 * kColGroups;
    // with an odd number of 64-column chunks (ncta = 160, 192, 320) the two warps of a lane quarter swap the longer
    // share every tile, so both drain 1.5 / 2.5 chunks per tile on average
    const int n_pass = (!kConv1 && p.two_pass) ? 2 : 1;          // two_pass: one drain per N half
    const int ncols = n_pass == 2 ? p.n_inst : p.ncta;           // columns per drain
    const int rotate = (!kConv1 && kColGroups == 2 && (((ncols + 63) / 64) & 1)) ? 1 : 0;
    int c_group = warp >> 2;
    if (lane == 0) tma_prefetch_desc(&p.tmap_out);
    int tcnt = 0, tidx = 0;
    for (int tile = tile0; tile < p.num_tiles; tile += tstride) {
      int row0, rv0 = 0, m0 = 0, m1 = 1;                         // first output row of this warp's box; member views
      if (kConv1) {
        const int unit = tile / p.tiles_per_group;
        const int b = unit / p.vg.n_groups, g = unit - b * p.vg.n_groups;
        rv0 = b * p.n_views;
        m0 = p.vg.start[g]; m1 = p.vg.start[g + 1];
        row0 = (tile - unit * p.tiles_per_group) * TILE_M + quarter * 32;
        if (row0 >= p.t_out) m1 = m0;                            // this warp's rows are all past the clip-view's end
      } else {
        row0 = tile * TILE_M + quarter * 32;
      }
      if (kProfile && (p.knockout & 8)) m1 = m0;
      for (int h = 0; h < n_pass; ++h, c_group ^= rotate) {
        const int c_first = c_group * 64;
        const int ai = n_pass == 2 ? h : acc;                    // accumulator barrier / region of this drain
        const int cbase = n_pass == 2 ? h * ncols : 0;           // first column of this drain inside the CTA's columns
        if (warp == 0 && lane == 0) trace_ev(trace, 0, tcnt, 1, tidx);
        mbar_wait(&acc_full[ai], acc_phase);
        if (warp == 0 && lane == 0) trace_ev(trace, 0, tcnt, 2, tidx);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                               static_cast<uint32_t>(n_pass == 2 ? cbase : acc * p.ncta);
        for (int mem = m0; mem < m1; ++mem) {
          const float gain = kConv1 ? p.vg.gain[mem] : 1.0f;
          const int rv = kConv1 ? rv0 + p.vg.view[mem] : 0;
          uint32_t va[32], vb[32];
          if (c_first < ncols) tmem_ld32(taddr + c_first, va);
          for (int c0 = c_first; c0 < ncols; c0 += c_step) {
            const bool tail = ncols - c0 < 64;                   // last 32 columns of a 160- or 96-column slice
            uint8_t* box = my_out + obuf * OUT_STAGE_BYTES;
            if (lane == 0) {                                     // the store that last used this buffer has read it
              if (p.out_bufs == 2) bulk_wait_group_read<1>(); else bulk_wait_group_read<0>();
            }
            __syncwarp();
            tmem_ld_wait();
            if (!tail) {
              uint8_t* row_base = box + lane * ROW_BYTES;
              tmem_ld32(taddr + c0 + 32, vb);
              epilogue_chunk<kConv1>(va, s_shift + cbase + c0, row_base, 0, lane & 7, gain);
              tmem_ld_wait();
              if (c0 + c_step < ncols) tmem_ld32(taddr + c0 + c_step, va);
              epilogue_chunk<kConv1>(vb, s_shift + cbase + c0 + 32, row_base, 4, lane & 7, gain);
            } else {
              epilogue_chunk<kConv1>(va, s_shift + cbase + c0, box + lane * (ROW_BYTES / 2), 0, 0, gain);   // dense 64-byte rows
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && !(kProfile && (p.knockout & 1))) {
              if (kConv1) tma_store_3d(&p.tmap_out, c0, row0, rv, box);
              else if (tail) tma_store_2d(&p.tmap_tail, col0 + cbase + c0, row0, box);
              else tma_store_2d(&p.tmap_out, col0 + cbase + c0, row0, box);
              bulk_commit_group();
            }
            obuf = (obuf + 1) & (p.out_bufs - 1);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[ai]);              // one arrival per warp
        if (warp == 0 && lane == 0) trace_ev(trace, 0, tcnt, 3, tidx++);
      }
      if (n_pass == 2) acc_phase ^= 1;                           // each half's barrier is used once per tile
      else if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) bulk_wait_group_all();
  } else if (warp == MMA_WARP) {
    // =========================== MMA issuer ===========================
    // all 32 lanes walk the loop (uniform control flow and registers); one elected lane issues.
    // A operand ring: conv1: one stage = both slabs of a tile; dw_pw: one stage = one K slab; freed by this warp's commit.
    // This warp is the one strictly serial role of the pipeline (every K slab of every tile passes through
    // it), so its loop keeps every parameter in a register (the `+ zi` idiom of tc_common.cuh) and does nothing
    // but wait / fence / issue.
    {
      const uint32_t idesc = ((umma_idesc_f16(TILE_M, p.n_inst, /*fp16*/ 0)) + zi);
      const uint32_t a_lo0 = ((umma_desc_lo(smem_u32(a_base))) + zi), b_lo0 = ((umma_desc_lo(smem_u32(b_base))) + zi);
      const uint32_t a_stage_lo = ((static_cast<uint32_t>(p.a_stage_bytes) >> 4) + zi);
      const uint32_t b_block_lo = ((b_block_bytes >> 4) + zi);
      const int a_depth = ((p.a_stages) + zi);
      uint64_t* const a_free = a_empty;
      const int num_kb = ((p.num_kb) + zi), n_halves = ((p.n_halves) + zi), last_ksteps = ((p.last_ksteps) + zi);
      const int n_inst = ((p.n_inst) + zi), ncta = ((p.ncta) + zi), b_stages = ((p.b_stages) + zi), acc_stages = ((p.acc_stages) + zi);
      const int num_tiles = ((p.num_tiles) + zi), tstep = ((tstride) + zi);
      const bool resident = ((p.b_resident) + zi) != 0, ko_mma = kProfile && ((p.knockout & 4) + zi) != 0;
      const uint32_t tmem0 = ((tmem_base) + zi);
      int sa = 0; uint32_t pa = 0; int sb = 0; uint32_t pb = 0; int acc = 0; uint32_t acc_phase = 0;
      if (resident && tile0 < num_tiles) {                       // the weights arrive once per launch
        for (int j = 0; j < blocks_per_tile; ++j) mbar_wait(&b_full[j], 0);
      }
      if constexpr (!kConv1) {
        // dw_pw: one predicated PTX sequence per (K slab, N half): 4 MMAs + the commits that free the weight
        // slot (streamed weights) and, after the last N half, the A stage.
        const uint32_t a_full0 = smem_u32(a_full) + z, a_free0 = smem_u32(a_free) + z, b_full0 = smem_u32(b_full) + z;
        const uint32_t b_empty0 = smem_u32(b_empty) + z, acc_full0 = smem_u32(acc_full) + z;
        int tcnt = 0, tn = 0;
        if (p.two_pass) {
          // two passes over the K slabs of a tile, one per N half; a_stages == num_kb, so slab kb of every tile sits in
          // A slot kb and both passes find it there (the slot is released by the second pass's commit only)
          uint32_t tile_ph = 0;                                  // every a_full / acc barrier completes once per tile
          for (int tile = tile0; tile < num_tiles; tile += tstep) {
            for (int h = 0; h < 2; ++h) {
              if (lane == 0) trace_ev(trace, 1, tcnt, 1, tn);
              mbar_wait(&acc_empty[h], tile_ph ^ 1);
              if (lane == 0) trace_ev(trace, 1, tcnt, 2, tn);
              const uint32_t d = tmem0 + static_cast<uint32_t>(h * n_inst);
              for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait_addr(a_full0 + 8u * kb, tile_ph);
                if (lane == 0) trace_ev(trace, 1, tcnt, 3, tn);
                const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(kb) * a_stage_lo;
                uint32_t b_lo, b_bar = 0u;
                if (!resident) {
                  mbar_wait_addr(b_full0 + 8u * sb, pb);
                  b_lo = b_lo0 + static_cast<uint32_t>(sb) * b_block_lo;
                  b_bar = b_empty0 + 8u * sb;
                } else {
                  b_lo = b_lo0 + static_cast<uint32_t>(kb * 2 + h) * b_block_lo;
                }
                tc_fence_after();
                const uint32_t a_bar = h == 1 ? a_free0 + 8u * kb : 0u;
                if (!ko_mma) umma_slab4_commit(d, a_lo, b_lo, idesc, kb != 0 ? 1u : 0u, b_bar, a_bar);
                else { if (b_bar) umma_commit_elect(b_bar); if (a_bar) umma_commit_elect(a_bar); }
                if (!resident) { if (++sb == b_stages) { sb = 0; pb ^= 1; } }
                if (lane == 0) trace_ev(trace, 1, tcnt, 4, tn++);
              }
              umma_commit_elect(acc_full0 + 8u * h);
            }
            tile_ph ^= 1;
          }
        } else
        for (int tile = tile0; tile < num_tiles; tile += tstep) {
          if (lane == 0) trace_ev(trace, 1, tcnt, 1, tn);
          mbar_wait(&acc_empty[acc], acc_phase ^ 1);
          if (lane == 0) trace_ev(trace, 1, tcnt, 2, tn);
          const uint32_t d0 = tmem0 + static_cast<uint32_t>(acc * ncta);
          uint32_t b_lo = b_lo0;                                 // resident: block j = (kb, nh) in order
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait_addr(a_full0 + 8u * sa, pa);
            if (lane == 0) trace_ev(trace, 1, tcnt, 3, tn);
            const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(sa) * a_stage_lo;
            const uint32_t a_bar = a_free0 + 8u * sa;
            if (resident && n_halves == 1) {
              tc_fence_after();
              if (!ko_mma) umma_slab4_commit(d0, a_lo, b_lo, idesc, kb != 0 ? 1u : 0u, a_bar, 0u);
              else umma_commit_elect(a_bar);
              b_lo += b_block_lo;
            } else {
              for (int nh = 0; nh < n_halves; ++nh) {
                uint32_t b_bar = 0u;
                if (!resident) {
                  mbar_wait_addr(b_full0 + 8u * sb, pb);
                  b_lo = b_lo0 + static_cast<uint32_t>(sb) * b_block_lo;
                  b_bar = b_empty0 + 8u * sb;
                }
                tc_fence_after();
                umma_slab4_commit(d0 + static_cast<uint32_t>(nh * n_inst), a_lo, b_lo, idesc, kb != 0 ? 1u : 0u, b_bar,
                                  nh == n_halves - 1 ? a_bar : 0u);
                if (!resident) { if (++sb == b_stages) { sb = 0; pb ^= 1; } }
                else b_lo += b_block_lo;
              }
            }
            if (lane == 0) trace_ev(trace, 1, tcnt, 4, tn++);
            if (++sa == a_depth) { sa = 0; pa ^= 1; }
          }
          umma_commit_elect(acc_full0 + 8u * acc);
          if (++acc == acc_stages) { acc = 0; acc_phase ^= 1; }
        }
      } else {
      for (int tile = tile0; tile < num_tiles; tile += tstep) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        const uint32_t d0 = tmem0 + static_cast<uint32_t>(acc * ncta);
        mbar_wait(&a_full[sa], pa);                              // one stage = both slabs of the tile
        uint32_t b_lo = b_lo0;                                   // resident: block j = (kb, nh) in order
        for (int kb = 0; kb < num_kb; ++kb) {
          const uint32_t a_lo = a_lo0 + static_cast<uint32_t>(sa) * a_stage_lo + static_cast<uint32_t>(kb) * (A_SLAB_BYTES >> 4);
          const int ksteps = (kb == num_kb - 1) ? last_ksteps : 4;
          for (int nh = 0; nh < n_halves; ++nh) {
            if (!resident) {
              mbar_wait(&b_full[sb], pb);
              b_lo = b_lo0 + static_cast<uint32_t>(sb) * b_block_lo;
            }
            tc_fence_after();
            if (elect_one()) {
              const uint32_t d = d0 + static_cast<uint32_t>(nh * n_inst);
              umma_f16_lo(d, a_lo, b_lo, idesc, kb != 0 ? 1u : 0u);
              for (int ks = 1; ks < ksteps; ++ks) umma_f16_lo(d, a_lo + 2 * ks, b_lo + 2 * ks, idesc, 1u);
              if (!resident) umma_commit(&b_empty[sb]);
            }
            __syncwarp();
            if (!resident) { if (++sb == b_stages) { sb = 0; pb ^= 1; } }
            else b_lo += b_block_lo;
          }
        }
        if (elect_one()) {
          umma_commit(&a_free[sa]);
          umma_commit(&acc_full[acc]);
        }
        __syncwarp();
        if (++sa == a_depth) { sa = 0; pa ^= 1; }
        if (++acc == acc_stages) { acc = 0; acc_phase ^= 1; }
      }
      }
    }
  } else if (warp == LOAD_WARP) {
    // =========================== weight loader ===========================
    if (lane == 0) {
      auto load_block = [&](int j, int slot, uint64_t* bar) {
        if (kProfile && (p.knockout & 32)) { mbar_arrive(bar); return; }   // no weight traffic (stale smem as weights)
        mbar_arrive_expect_tx(bar, b_block_bytes);
        // image = [K slab][cout rows x 128 B]; block j = (K slab j / n_halves, rows col0 + (j % n_halves) * n_inst ...)
        const uint8_t* src = p.w_img + (static_cast<size_t>(j / p.n_halves) * p.cout + col0 + (j % p.n_halves) * p.n_inst) * ROW_BYTES;
        uint8_t* dst = b_base + slot * b_block_bytes;
        for (uint32_t o = 0; o < b_block_bytes; o += 16384)
          bulk_g2s(dst + o, src + o, min(16384u, b_block_bytes - o), bar);
      };
      if (p.b_resident) {
        for (int j = 0; j < blocks_per_tile; ++j) load_block(j, j, &b_full[j]);
      } else {
        int sb = 0; uint32_t pb = 0;
        for (int tile = tile0; tile < p.num_tiles; tile += tstride) {
          for (int jj = 0; jj < blocks_per_tile; ++jj) {
            // block index = kb * n_halves + nh; two_pass consumes them half by half: (nh, kb) order
            const int j = p.two_pass ? (jj % p.num_kb) * 2 + jj / p.num_kb : jj;
            mbar_wait(&b_empty[sb], pb ^ 1);
            load_block(j, sb, &b_full[sb]);
            if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else if (warp == RAW_WARP) {
    // =========================== raw activation loader (TMA) ===========================
    if (!kConv1 && lane == 0) {
      int rs = 0; uint32_t pr = 0;
      int tcnt = 0, tn = 0;
      // The raw ring holds only 2-6 slabs (39-100 KB) per SM, too few bytes in flight to cover the DRAM
      // latency of a whole tile; the rows of the tiles this CTA will work on next are therefore pulled
      // into L2 ahead of time.  A tile's rows x all channels are ONE contiguous range of the activation.
      const bool do_prefetch = p.prefetch_tiles > 0 && col0 == 0;    // one CTA of a column-split group is enough
      auto prefetch_tile = [&](int tile) {
        if (tile >= p.num_tiles) return;
        const int m0 = tile * TILE_M, m1 = min(m0 + TILE_M, p.rows_out) - 1;
        const int v0 = div_t_out(p, m0), t0 = m0 - v0 * p.t_out;
        const int v1 = div_t_out(p, m1), t1 = m1 - v1 * p.t_out;
        long long r0 = static_cast<long long>(v0) * p.t_in + t0 * kStride - p.pad_left;
        long long r1 = static_cast<long long>(v1) * p.t_in + t1 * kStride - p.pad_left + 3;
        r0 = r0 < 0 ? 0 : r0;
        r1 = r1 > p.rows_in ? p.rows_in : r1;
        const uint32_t row_bytes = static_cast<uint32_t>(p.cin) * 2u;
        const uint8_t* src = reinterpret_cast<const uint8_t*>(p.in_act) + r0 * row_bytes;
        const long long bytes = (r1 - r0) * row_bytes;
        for (long long o = 0; o < bytes; o += 32768)
          bulk_prefetch_l2(src + o, static_cast<uint32_t>(bytes - o < 32768 ? bytes - o : 32768));
      };
      if (do_prefetch)
        for (int d = 1; d <= p.prefetch_tiles; ++d) prefetch_tile(tile0 + d * tstride);
      for (int tile = tile0; tile < p.num_tiles; tile += tstride) {
        const int m0 = tile * TILE_M;
        const int v0 = div_t_out(p, m0), t0 = m0 - v0 * p.t_out;
        const int lo = v0 * p.t_in + t0 * kStride - p.pad_left;
        if (do_prefetch && tile != tile0) prefetch_tile(tile + p.prefetch_tiles * tstride);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          trace_ev(trace, 2, tcnt, 1, tn);
          mbar_wait(&raw_empty[rs], pr ^ 1);
          trace_ev(trace, 2, tcnt, 2, tn++);
          if (kProfile && (p.knockout & 16)) { mbar_arrive(&raw_full[rs]); if (++rs == p.raw_stages) { rs = 0; pr ^= 1; } continue; }
          mbar_arrive_expect_tx(&raw_full[rs], static_cast<uint32_t>(p.raw_stage_bytes));
          uint8_t* dst = raw_base + rs * p.raw_stage_bytes;
          for (int bx = 0; bx < p.n_boxes; ++bx)
            tma_load_2d(dst + bx * p.box_rows * ROW_BYTES, &p.tmap_in, kb * SLAB_K, lo + bx * p.box_rows,
                        &raw_full[rs]);
          if (++rs == p.raw_stages) { rs = 0; pr ^= 1; }
        }
      }
    }
  } else {
    // =========================== A producers ===========================
    const int ptid = tid - PROD_WARP0 * 32;
    if constexpr (kConv1) {
      int sa = 0; uint32_t pa = 0, buf = 0;
      using Conv1Producer = Conv1ProducerT<CONV1_PROD_THREADS>;
      float4 q[Conv1Producer::QUADS];
      const float* x;
      Conv1Tile cur{}, nxt{};
      int tile = tile0;
      if (tile < p.num_tiles) { cur = Conv1Producer::describe(p, tile, &x); Conv1Producer::load(cur, x, ptid, q); }
      for (; tile < p.num_tiles; tile += tstride) {
        __half* win = reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(s_win) + buf * CONV1_WIN_BYTES);
        Conv1Producer::store(cur, ptid, q, win);
        const int next = tile + tstride;
        if (next < p.num_tiles) {                                // next tile's loads fly while this one is filled
          nxt = Conv1Producer::describe(p, next, &x);
          Conv1Producer::load(nxt, x, ptid, q);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(CONV1_PROD_THREADS) : "memory");   // window complete
        mbar_wait(&a_empty[sa], pa ^ 1);
        Conv1Producer::fill_slabs(a_base + sa * p.a_stage_bytes, ptid, win, cur.rows);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[sa]);
        if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
        // no second barrier: the next tile writes the other window buffer, and this buffer is only
        // rewritten after the next tile's bar.sync, which every thread reaches after this fill
        buf ^= 1;
        cur = nxt;
      }
    } else {
      // all 256 producer threads work on every slab, in ring order (any ring depth >= 2 works)
      const int raw_stages = p.raw_stages + zi, a_stages = p.a_stages + zi, num_kb = p.num_kb + zi, cin = p.cin + zi;
      const int num_tiles = p.num_tiles + zi, tstep = tstride + zi;
      const uint32_t raw_stage_bytes = static_cast<uint32_t>(p.raw_stage_bytes) + z, a_stage_bytes = static_cast<uint32_t>(p.a_stage_bytes) + z;
      const bool ko_fir = kProfile && ((p.knockout & 2) + zi) != 0;
      const int c = ptid & 7;
      int rs = 0, sa = 0, par = 0; uint32_t pr = 0, pa = 0;
      int tcnt = 0, tn = 0;
      for (int tile = tile0; tile < num_tiles; tile += tstep) {
        uint32_t* meta = s_meta + par * TILE_M;
        if (ptid < TILE_M) meta[ptid] = row_meta(p, kStride, tile, ptid);
        asm volatile("bar.sync 1, %0;" ::"n"(NUM_PROD_THREADS) : "memory");
        const RunDesc rd = decode_run(meta, ptid);
        for (int kb = 0; kb < num_kb; ++kb) {
          const __half* tp = s_dwh + kb * SLAB_K + c * 8;          // taps of this thread's 8 channels
          const uint4 k0 = *reinterpret_cast<const uint4*>(tp);
          const uint4 k1 = *reinterpret_cast<const uint4*>(tp + cin);
          const uint4 k2 = *reinterpret_cast<const uint4*>(tp + 2 * cin);
          uint4 o[RUN_ROWS];
          if (ptid == 0) trace_ev(trace, 3, tcnt, 1, tn);
          mbar_wait(&raw_full[rs], pr);
          if (ptid == 0) trace_ev(trace, 3, tcnt, 2, tn);
          if (ko_fir) { o[0] = o[1] = o[2] = o[3] = k0; }
          else fir_run<kStride>(raw_base + rs * raw_stage_bytes, rd, k0, k1, k2, c, o);
          // every raw row of this slab has been consumed (the FIR outputs exist): the slot may be refilled
          asm volatile("" ::"r"(o[0].x), "r"(o[1].x), "r"(o[2].x), "r"(o[3].x) : "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(&raw_empty[rs]);            // one arrival per warp
          if (ptid == 0) trace_ev(trace, 3, tcnt, 3, tn);
          mbar_wait(&a_empty[sa], pa ^ 1u);
          if (ptid == 0) trace_ev(trace, 3, tcnt, 4, tn);
          uint8_t* slab = a_base + sa * a_stage_bytes;
#pragma unroll
          for (int i = 0; i < RUN_ROWS; ++i) *reinterpret_cast<uint4*>(slab + rd.dst_off[i]) = o[i];
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_full[sa]);
          if (ptid == 0) trace_ev(trace, 3, tcnt, 5, tn++);
          if (++rs == raw_stages) { rs = 0; pr ^= 1u; }
          if (++sa == a_stages) { sa = 0; pa ^= 1u; }
        }
        par ^= 1;
      }
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// Fused conv1d_1 + block 1 (model.py:805-812).  The conv1d_1 activation [399, C0] -- the largest
// tensor of the network -- never leaves the SM: its accumulator is read from TMEM by four
// "middle" warps that apply gain / BN shift / ReLU6, round to fp16 (the same rounding the
// unfused path stores) into a swizzled shared-memory row buffer, then run the k=3 depthwise FIR of
// block 1 over it (each thread slides over 8 consecutive rows of one 8-channel chunk, like
// produce_slab) and write the UMMA-swizzled A operand of the pointwise GEMM, whose accumulator is
// drained by four output warps exactly like tc_gemm_kernel.  (A first version did the FIR across
// lanes with warp shuffles straight from the TMEM rows; it needed ~4x the instructions and the
// stage is issue-bound.)
// A tile = 128 conv1d_1 rows -> 126 block-1 rows (stride-1 VALID), 4 tiles per clip-view.
// TMEM: 2 x C0 columns (conv1d_1 accumulators) + 2 x C1 columns (pointwise accumulators) <= 512.
//   warps 0-7  middle  : acc1 -> gain, shift, ReLU6, fp16 -> row buffer | FIR -> A2 slabs; two warps per TMEM
//                        lane quarter, each on half of the channels (this stage is issue/latency-bound: it
//                        needs two warps per scheduler to hide its own instruction latencies)
//   warps 8-11 output  : acc2 -> shift, ReLU6, fp16 -> swizzled box -> TMA store
//   warp  12   MMA     : one bulk copy of both weight images (resident); conv1d_1 GEMM (K = 80) once per
//                        (clip, view group, tile), pointwise GEMM per view
//   warps 13-15 window : Conv1Producer (waveform window one tile ahead, swizzled A1 fill)
// ------------------------------------------------------------------------------------------------
constexpr int FUSE_ROWS = TILE_M - 2;                            // block-1 rows per tile
// 16 warps = 512 threads: the register file then allows 128 registers per thread (the middle warps are the
// critical role and spill at 96)
constexpr int FUSE_MID_WARPS = 8, FUSE_OUT_WARP0 = 8, FUSE_PROD_THREADS = 96;
// 16 warps = 512 threads, so that every role may use up to 128 registers (the middle warps need them).  (r02d tried
// 20 warps -- 8 output warps -- with setmaxnreg re-balancing: an inc can only take registers that a dec of the SAME CTA
// released, the launch-time slack of the register file is not in that pool, and the kernel dead-locked.)
template <bool kT> struct FuseRoles {
  static constexpr int OUT_WARPS = 4;
  static constexpr int MMA_WARP = FUSE_OUT_WARP0 + OUT_WARPS;
  static constexpr int PROD_WARP0 = MMA_WARP + 1;
  static constexpr int THREADS = 32 * PROD_WARP0 + FUSE_PROD_THREADS;   // 512
};
constexpr uint32_t FUSE_A2_LBO = 1024, FUSE_A2_SBO = 2048;        // transposed form: MN-major A2 (see the middle warps)
constexpr int FUSE_OUT_BOX_BYTES = 8 * OUT_STAGE_BYTES;          // 4 warps x 2 boxes, or 8 warps x 1 box

struct alignas(64) FusedParams {
  CUtensorMap tmap_out;      // block-1 output, 3-D [clip-views, t2, c1], box [64 ch, 32 rows, 1], 128-byte swizzle
  CUtensorMap tmap_out30;    // same with 30-row boxes: the last lane quarter owns rows 96..125 of a tile only
  const float* wav;
  ViewGroups vg;
  int n_views;
  const uint8_t* w1_img;     // conv1d_1: 2 slabs of c0 rows (80 folded taps)
  const uint8_t* w2_img;     // conv1d_2: c0/64 slabs of c1 rows
  const __half* dw_h;        // depthwise_conv2d_1 taps [3][c0]
  const float* shift1;       // [c0]
  const float* shift2;       // [c1]
  int c0, c1, t1, t2;        // channels; rows per clip-view after conv1d_1 (399) / after block 1 (397)
  int blocks_per_view;       // ceil(t2 / 126)
  int num_units;             // clips * view groups * blocks_per_view
  int knockout;              // KWS_FKNOCK (profiling build only, results become wrong): 1 output warps skip conversion + store,
                             // 2 middle warps skip the FIR, 4 no TMA stores, 8 no pointwise MMAs
};

struct FusedSmem { uint32_t a1, w1, w2, a2, raw, out, win, sh1, sh2, taps, bars, total; };

__host__ __device__ inline FusedSmem fused_smem(int c0, int c1) {
  FusedSmem s; uint32_t o = 0;
  s.a1 = o; o += 2 * A_SLAB_BYTES;
  s.w1 = o; o += 2u * c0 * ROW_BYTES;
  s.w2 = o; o += static_cast<uint32_t>(c0 / SLAB_K) * c1 * ROW_BYTES;
  // two row buffers: relu6(bn(conv1d_1)) rows of one view, swizzled like A slabs, FIR-filtered IN PLACE into the A
  // operand of the pointwise GEMM; view v + 1 is packed and filtered in the other buffer while the MMA reads this one
  s.a2 = o; o += 2u * static_cast<uint32_t>(c0 / SLAB_K) * A_SLAB_BYTES;
  s.raw = s.a2;
  s.out = o; o += FUSE_OUT_BOX_BYTES;                                    // (the r01 FIR of the last rows reads 2 rows past a raw slab: into this region)
  s.win = o; o += 2 * CONV1_WIN_BYTES;
  s.sh1 = o; o += c0 * 4u;
  s.sh2 = o; o += c1 * 4u;
  s.taps = o; o += (3u * c0 * 2 + 15) & ~15u;
  s.bars = o; o += 16 * 8 + 16;
  s.total = o + 1024;
  return s;
}

// 8 accumulator values -> gain, shift, ReLU6 -> 8 fp16 (one 16-byte chunk)
__device__ __forceinline__ uint4 pack8_relu6(const uint32_t* v, const float* sh, float gain) {
  const __half2 six = __float2half2_rn(6.0f);
  uint32_t o[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const uint32_t a = pack_relu_f16x2(fmaf(__uint_as_float(v[2 * e]), gain, sh[2 * e]),
                                       fmaf(__uint_as_float(v[2 * e + 1]), gain, sh[2 * e + 1]));
    const __half2 ha = __hmin2(*reinterpret_cast<const __half2*>(&a), six);
    o[e] = *reinterpret_cast<const uint32_t*>(&ha);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

// fp16 x fp16 + fp32 -> fp32 with the halves picked from packed registers (FHFMA with .H0 / .H1 operand selectors)
#define KWS_FHFMA(AH, BH)                                                                                   \
  asm("{\n\t.reg .b16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\t"                     \
      "fma.rn.f32.f16 %0, " AH ", " BH ", %3;\n\t}"                                                          \
      : "=f"(d) : "r"(a), "r"(b), "f"(c))
template <bool A_HI, bool B_HI>
__device__ __forceinline__ float fhfma_sel(uint32_t a, uint32_t b, float c) {
  float d;
  if constexpr (A_HI && B_HI) KWS_FHFMA("ah", "bh");
  else if constexpr (A_HI) KWS_FHFMA("ah", "bl");
  else if constexpr (B_HI) KWS_FHFMA("al", "bh");
  else KWS_FHFMA("al", "bl");
  return d;
}
#undef KWS_FHFMA

// kT (c0 == 128): conv1d_1 is computed TRANSPOSED -- D1^T[channel, time] = W80^T x window^T, i.e. the weight image is
// the A operand (M = 128 channels) and the staged window rows are the B operand (N = 128 time steps); both are K-major
// SW128 slabs, so only the two descriptors swap.  A TMEM lane then holds ONE channel over time, and the middle
// warps (lane = channel) do gain / BN shift / ReLU6 / fp16 rounding and the k = 3 depthwise FIR along the registers of
// a thread: no shared-memory round trip, no barrier between the middle warps, per-thread scalar shift and taps.  The
// accumulator of a (clip, view group, tile) unit is read from TMEM ONCE into registers and reused for every member
// view of the group (TMEM reads are 64 B / cycle / SM; re-reading it per view was a third of the kernel's floor).
template <bool kT>
__global__ void __launch_bounds__(FuseRoles<kT>::THREADS, 1) conv1_block1_kernel(const __grid_constant__ FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int FUSE_OUT_WARPS = FuseRoles<kT>::OUT_WARPS, FUSE_MMA_WARP = FuseRoles<kT>::MMA_WARP,
                FUSE_PROD_WARP0 = FuseRoles<kT>::PROD_WARP0, FUSE_THREADS = FuseRoles<kT>::THREADS;
  const FusedSmem lay = fused_smem(p.c0, p.c1);
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a1_base = smem + lay.a1;
  uint8_t* w1_base = smem + lay.w1;
  uint8_t* w2_base = smem + lay.w2;
  uint8_t* a2_base = smem + lay.a2;
  uint8_t* out_base = smem + lay.out;
  uint8_t* win_base = smem + lay.win;
  [[maybe_unused]] uint8_t* raw_base = smem + lay.raw;   // (row-major form only)
  float* s_sh1 = reinterpret_cast<float*>(smem + lay.sh1);
  float* s_sh2 = reinterpret_cast<float*>(smem + lay.sh2);
  __half* s_taps = reinterpret_cast<__half*>(smem + lay.taps);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bars);
  uint64_t* a1_full = bars;            // [1]
  uint64_t* a1_empty = bars + 1;       // [1]
  uint64_t* acc1_full = bars + 2;      // [2]
  uint64_t* acc1_empty = bars + 4;     // [2]
  uint64_t* a2_full = bars + 6;        // [2]
  uint64_t* a2_empty = bars + 8;       // [2]
  uint64_t* acc2_full = bars + 10;     // [2]
  uint64_t* acc2_empty = bars + 12;    // [2]
  uint64_t* w_full = bars + 14;        // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nkb2 = p.c0 / SLAB_K;                                // K slabs of the pointwise GEMM
  const int nch = p.c0 / 8;                                      // 16-byte channel chunks of a conv1d_1 row

  for (int i = tid; i < p.c0; i += FUSE_THREADS) s_sh1[i] = p.shift1[i];
  for (int i = tid; i < p.c1; i += FUSE_THREADS) s_sh2[i] = p.shift2[i];
  for (int i = tid; i < 3 * p.c0; i += FUSE_THREADS) s_taps[i] = p.dw_h[i];
  if (warp == FUSE_MMA_WARP) {
    if (lane == 0) {
      mbar_init(a1_full, FUSE_PROD_THREADS);
      mbar_init(a1_empty, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&acc1_full[i], 1); mbar_init(&acc1_empty[i], FUSE_MID_WARPS);
        mbar_init(&a2_full[i], FUSE_MID_WARPS); mbar_init(&a2_empty[i], 1);
        mbar_init(&acc2_full[i], 1); mbar_init(&acc2_empty[i], FUSE_OUT_WARPS * 32);
      }
      mbar_init(w_full, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc1_cols = kT ? static_cast<uint32_t>(TILE_M) : static_cast<uint32_t>(p.c0);   // columns of one conv1d_1 accumulator stage
  const uint32_t acc2_col0 = 2u * acc1_cols;

  // Each CTA works on a contiguous range of units (tile j of view group g of clip b, j fastest): the (b, g, j) of
  // the next unit is an increment (the divisions of decode() were ~700 cycles of the middle warps' chain per unit),
  // and the twelve units of a clip read its waveform from the same SM.
  const int u0 = static_cast<int>(static_cast<long long>(p.num_units) * blockIdx.x / gridDim.x);
  const int u1 = static_cast<int>(static_cast<long long>(p.num_units) * (blockIdx.x + 1) / gridDim.x);
  auto decode = [&](int unit, int& b, int& g, int& j) {
    j = unit % p.blocks_per_view;
    const int ug = unit / p.blocks_per_view;
    g = ug % p.vg.n_groups;
    b = ug / p.vg.n_groups;
  };
  auto advance = [&](int& b, int& g, int& j) {                   // (b, g, j) of unit + 1
    if (++j == p.blocks_per_view) { j = 0; if (++g == p.vg.n_groups) { g = 0; ++b; } }
  };

  if (kT && warp < FUSE_MID_WARPS) {
    // =========================== middle (transposed): lane = channel, registers = time ===========================
    // warp = (lane quarter q, time half hf): channel c = 32 q + lane, block-1 rows [64 hf, 64 hf + 64) of the tile
    // (rows 126, 127 of the second half are computed from zeros and never stored to global memory).
    const int q = warp & 3, hf = warp >> 2;
    const int c = q * 32 + lane;
    const float shift = s_sh1[c];
    const uint32_t k01 = static_cast<uint32_t>(__half_as_ushort(s_taps[c])) |
                         (static_cast<uint32_t>(__half_as_ushort(s_taps[p.c0 + c])) << 16);     // taps 0, 1 of this channel
    const uint32_t k2 = static_cast<uint32_t>(__half_as_ushort(s_taps[2 * p.c0 + c]));           // tap 2
    // A2 is written in the MN-major SW128 operand layout: a 128-byte row holds 64 consecutive time steps of ONE channel
    // (8 channels per 1024-byte atom, chunk j of a row at position j ^ (c % 8)), atom (c / 8, time half) at
    // (c / 8) * 2048 + half * 1024.  This thread's 64 outputs are exactly one row: 8 conflict-free 16-byte stores
    // instead of 64 scattered 2-byte ones (K-major A2 in r02b/c: 1.4 shared-memory wavefronts per 2-byte store).
    const uint32_t row_addr = smem_u32(a2_base) + static_cast<uint32_t>(c >> 3) * FUSE_A2_SBO + static_cast<uint32_t>(hf) * FUSE_A2_LBO +
                              (static_cast<uint32_t>(c) & 7u) * ROW_BYTES;
    const uint32_t csw = static_cast<uint32_t>(c) & 7u;
    const uint32_t buf_bytes = static_cast<uint32_t>(nkb2) * A_SLAB_BYTES;
    const __half2 six = __float2half2_rn(6.0f);
    int n2 = 0, i = 0;
    int b = 0, g = 0, j = 0;
    if (u0 < u1) decode(u0, b, g, j);
    for (int unit = u0; unit < u1; ++unit, ++i, advance(b, g, j)) {
      const int s1 = i & 1;
      mbar_wait(&acc1_full[s1], static_cast<uint32_t>(i >> 1) & 1u);
      tc_fence_after();
      // conv1d_1 rows [64 hf, 64 hf + 66) of this channel: read ONCE, reused for every member view of the group
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(s1) * acc1_cols + 64u * hf;
      uint32_t a0[32], a1[32], a2[2];
      tmem_ld32(taddr, a0);
      tmem_ld32(taddr + 32, a1);
      if (hf == 0) tmem_ld2(taddr + 64, a2); else { a2[0] = 0u; a2[1] = 0u; }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc1_empty[s1]);               // the accumulator stage is free for the unit after next
      const bool on = 64 * hf < p.t2 - j * FUSE_ROWS;            // the last tile of a clip-view holds 19 block-1 rows
      for (int mem = p.vg.start[g]; mem < p.vg.start[g + 1]; ++mem, ++n2) {
        const float gain = p.vg.gain[mem];
        const int bsel = n2 & 1;
        mbar_wait(&a2_empty[bsel], (static_cast<uint32_t>(n2 >> 1) & 1u) ^ 1u);   // the MMA of view n2 - 2 has read this buffer
        if (on && !(kProfile && (p.knockout & 2))) {
          const uint32_t bo = static_cast<uint32_t>(bsel) * buf_bytes;
          // y[t] = fp16(relu6(gain * acc + shift)) -- the rounding the unfused path stores -- as packed pairs (y[2i], y[2i+1])
          auto ypair = [&](int ip) -> uint32_t {
            const uint32_t lo = ip < 16 ? a0[2 * (ip & 15)] : (ip < 32 ? a1[2 * (ip & 15)] : a2[0]);
            const uint32_t hi = ip < 16 ? a0[2 * (ip & 15) + 1] : (ip < 32 ? a1[2 * (ip & 15) + 1] : a2[1]);
            const uint32_t pk = pack_relu_f16x2(fmaf(__uint_as_float(lo), gain, shift), fmaf(__uint_as_float(hi), gain, shift));
            const __half2 h = __hmin2(*reinterpret_cast<const __half2*>(&pk), six);
            return *reinterpret_cast<const uint32_t*>(&h);
          };
          uint32_t cur = ypair(0);
#pragma unroll
          for (int jc = 0; jc < 8; ++jc) {                       // 16-byte chunk jc = block-1 rows 8 jc .. 8 jc + 7 of this half
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const uint32_t nxt = ypair(4 * jc + e + 1);
              // same order as fir3(): tap 0 first, fp32 accumulation, one rounding to fp16
              const float o0 = fhfma_sel<false, false>(nxt, k2, fhfma_sel<true, true>(cur, k01, fhfma_sel<false, false>(cur, k01, 0.0f)));
              const float o1 = fhfma_sel<true, false>(nxt, k2, fhfma_sel<false, true>(nxt, k01, fhfma_sel<true, false>(cur, k01, 0.0f)));
              o[e] = pack_f16x2(o0, o1);
              cur = nxt;
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_addr + bo + ((static_cast<uint32_t>(jc) ^ csw) << 4)),
                         "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a2_full[bsel]);              // one arrival per warp
      }
    }
  } else if (!kT && warp < FUSE_MID_WARPS) {
    // =========================== middle: acc1 -> depthwise FIR -> A2 ===========================
    const int q = warp & 3, hf = warp >> 2;                      // TMEM lane quarter, channel half
    const int row = q * 32 + lane;
    const int nchw = nch / 2, ch0 = hf * nchw;                   // pack phase: this warp's 16-byte channel chunks [ch0, ch0 + nchw)
    const int fkb = tid >> 7, ftg = tid & 127;                   // FIR phase: K slab and (chunk, 8-row run) of this thread
    const int fc = ftg & 7, fg = ftg >> 3;
    const uint32_t buf_bytes = static_cast<uint32_t>(nkb2) * A_SLAB_BYTES;
    // this thread's depthwise taps never change (fkb < nkb2 for every FIR thread that is ever on)
    const int fcg = (fkb < nkb2 ? fkb : 0) * 8 + fc;
    const uint4 k0 = *reinterpret_cast<const uint4*>(s_taps + fcg * 8);
    const uint4 k1 = *reinterpret_cast<const uint4*>(s_taps + p.c0 + fcg * 8);
    const uint4 k2 = *reinterpret_cast<const uint4*>(s_taps + 2 * p.c0 + fcg * 8);
    int n2 = 0, i = 0;
    int b = 0, g = 0, j = 0;
    if (u0 < u1) decode(u0, b, g, j);
    for (int unit = u0; unit < u1; ++unit, ++i, advance(b, g, j)) {
      const int s1 = i & 1;
      mbar_wait(&acc1_full[s1], static_cast<uint32_t>(i >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(s1 * p.c0 + ch0 * 8);
      // the last tile of a clip-view holds 21 conv1d_1 rows / 19 block-1 rows: quarters and runs past them are skipped
      const bool pack_on = q * 32 < p.t1 - j * FUSE_ROWS;
      const bool fir_on = fkb < nkb2 && 8 * fg < p.t2 - j * FUSE_ROWS;
      for (int mem = p.vg.start[g]; mem < p.vg.start[g + 1]; ++mem, ++n2) {
        const float gain = p.vg.gain[mem];
        const int bsel = n2 & 1;
        uint8_t* buf = a2_base + bsel * buf_bytes;
        mbar_wait(&a2_empty[bsel], (static_cast<uint32_t>(n2 >> 1) & 1u) ^ 1u);   // the MMA of view n2 - 2 has read this buffer
        // ---- pack: this row's half of relu6(bn(gain * conv1d_1)) as fp16 into the row buffer ----
        if (pack_on) {
          uint32_t va[32];
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            if (cc * 4 < nchw) {
              tmem_ld32(taddr + cc * 32, va);
              tmem_ld_wait();
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int cg = ch0 + cc * 4 + k;
                *reinterpret_cast<uint4*>(buf + (cg >> 3) * A_SLAB_BYTES + swz_off(row, cg & 7)) =
                    pack8_relu6(va + 8 * k, s_sh1 + cg * 8, gain);
              }
            }
          }
        }
        asm volatile("bar.sync 2, %0;" ::"n"(FUSE_MID_WARPS * 32) : "memory");   // all rows of this view are in the buffer
        // ---- depthwise FIR, in place: rows 8 fg .. 8 fg + 7 of chunk fc of K slab fkb (rows 126, 127 are never stored
        //      to global memory).  Output row r overwrites buffer row r, which the thread of the previous run still
        //      needs as its taps: every thread loads its 10 rows first, a barrier, then the stores. ----
        uint4 o[8];
        if (fir_on) {
          const uint8_t* rsl = buf + fkb * A_SLAB_BYTES + (8 * fg) * ROW_BYTES;
          uint4 x[10];
#pragma unroll
          for (int r = 0; r < 10; ++r) x[r] = lds128(rsl + r * ROW_BYTES + ((fc ^ (r & 7)) << 4));
#pragma unroll
          for (int r = 0; r < 8; ++r) o[r] = fir3(x[r], x[r + 1], x[r + 2], k0, k1, k2);
          asm volatile("" ::"r"(o[0].x), "r"(o[1].x), "r"(o[2].x), "r"(o[3].x), "r"(o[4].x), "r"(o[5].x), "r"(o[6].x),
                       "r"(o[7].x) : "memory");
        }
        asm volatile("bar.sync 2, %0;" ::"n"(FUSE_MID_WARPS * 32) : "memory");   // every tap has been read
        if (fir_on) {
          uint8_t* dst = buf + fkb * A_SLAB_BYTES + (8 * fg) * ROW_BYTES;
#pragma unroll
          for (int r = 0; r < 8; ++r) *reinterpret_cast<uint4*>(dst + r * ROW_BYTES + ((fc ^ r) << 4)) = o[r];
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a2_full[bsel]);              // one arrival per warp
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc1_empty[s1]);
    }
  } else if (warp < FUSE_MMA_WARP) {
    // =========================== output: acc2 -> BN shift, ReLU6 -> TMA store ===========================
    // This is the role every other one ends up waiting for (r02c profile: 83 % busy at 0.12 instructions per cycle, the
    // middle warps 39 %): one warp cannot hide its own TMEM / shared-memory / ALU latencies.  So: kColGroups warps per
    // TMEM lane quarter, each on every kColGroups-th 64-column chunk; both 32-column halves of a chunk are loaded
    // before the first wait; the accumulator stage is handed back as soon as its columns are in registers.
    constexpr int kColGroups = FUSE_OUT_WARPS / 4, kBoxes = 8 / FUSE_OUT_WARPS;
    const int ow = warp - FUSE_OUT_WARP0;
    const int q = ow & 3, c_first = (ow >> 2) * 64;              // TMEM lane quarter; first 64-column chunk
    constexpr int c_step = 64 * kColGroups;
    uint8_t* box0 = out_base + ow * kBoxes * OUT_STAGE_BYTES;
    int ob = 0, n2 = 0;
    if (lane == 0) { tma_prefetch_desc(&p.tmap_out); tma_prefetch_desc(&p.tmap_out30); }
    int b = 0, g = 0, j = 0;
    if (u0 < u1) decode(u0, b, g, j);
    for (int unit = u0; unit < u1; ++unit, advance(b, g, j)) {
      const int row0 = j * FUSE_ROWS + q * 32;                   // first block-1 row of this warp's box
      for (int mem = p.vg.start[g]; mem < p.vg.start[g + 1]; ++mem, ++n2) {
        const int s2 = n2 & 1;
        mbar_wait(&acc2_full[s2], static_cast<uint32_t>(n2 >> 1) & 1u);
        tc_fence_after();
        if (row0 < p.t2 && c_first < p.c1) {
          const int rv = b * p.n_views + p.vg.view[mem];
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc2_col0 + static_cast<uint32_t>(s2 * p.c1);
          // 32 columns per step, two register buffers: the load of step s + 1 is issued right after the wait of step s
          // and flies under the conversion of step s (tcgen05.wait::ld waits for ALL loads of the thread, so a load can
          // only overlap math that does not need it).  The four output warps run in lock step: without this the TMEM read
          // port (64 B / cycle / SM) idles while they convert and they idle while it reads (r02 knockouts: the two add up).
          uint32_t va[32], vb[32];
          const int n_steps = (p.c1 - c_first + c_step - 1) / c_step * 2;       // 32-column steps of this warp
          auto col_of = [&](int st) { return c_first + (st >> 1) * c_step + (st & 1) * 32; };
          tmem_ld32(taddr + col_of(0), va);
          for (int st = 0; st < n_steps; st += 2) {
            uint8_t* box = box0 + ob * OUT_STAGE_BYTES;
            uint8_t* row_base = box + lane * ROW_BYTES;
            const int c0 = col_of(st);
            if (lane == 0) bulk_wait_group_read<kBoxes - 1>();   // the store that last used this box has read it
            __syncwarp();
            tmem_ld_wait();
            tmem_ld32(taddr + col_of(st + 1), vb);
            if (!(kProfile && (p.knockout & 1))) epilogue_chunk<false>(va, s_sh2 + c0, row_base, 0, lane & 7, 1.0f);
            tmem_ld_wait();
            const bool more = st + 2 < n_steps;
            if (more) tmem_ld32(taddr + col_of(st + 2), va);
            else { tc_fence_before(); mbar_arrive(&acc2_empty[s2]); }   // every column of this view is in registers
            if (!(kProfile && (p.knockout & 1))) epilogue_chunk<false>(vb, s_sh2 + c0 + 32, row_base, 4, lane & 7, 1.0f);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && !(kProfile && (p.knockout & 5))) {
              tma_store_3d(q == 3 ? &p.tmap_out30 : &p.tmap_out, c0, row0, rv, box);
              bulk_commit_group();
            }
            ob = (ob + 1) & (kBoxes - 1);
          }
        } else {
          tc_fence_before();
          mbar_arrive(&acc2_empty[s2]);
        }
      }
    }
    if (lane == 0) bulk_wait_group_all();
  } else if (warp == FUSE_MMA_WARP) {
    // =========================== MMA issuer ===========================
    // all 32 lanes walk the loop (uniform control flow and registers); one elected lane issues
    {
      const uint32_t idesc1 = kT ? umma_idesc_f16(p.c0, TILE_M, /*fp16*/ 0) : umma_idesc_f16(TILE_M, p.c0, /*fp16*/ 0);
      const uint32_t idesc2 = umma_idesc_f16(TILE_M, p.c1, /*fp16*/ 0);
      const uint32_t a1 = umma_desc_lo(smem_u32(a1_base)), w1 = umma_desc_lo(smem_u32(w1_base));
      const uint32_t w2 = umma_desc_lo(smem_u32(w2_base)), a2 = umma_desc_lo(smem_u32(a2_base));
      const uint32_t w1_slab = static_cast<uint32_t>(p.c0) * (ROW_BYTES >> 4), w2_slab = static_cast<uint32_t>(p.c1) * (ROW_BYTES >> 4);
      const uint32_t a2_buf_lo = static_cast<uint32_t>(nkb2) * (A_SLAB_BYTES >> 4);
      const uint32_t a2_empty0 = smem_u32(a2_empty), acc2_full0 = smem_u32(acc2_full);
      if (lane == 0) {                                           // both weight images, once (resident)
        const uint32_t b1 = 2u * p.c0 * ROW_BYTES, b2 = static_cast<uint32_t>(nkb2) * p.c1 * ROW_BYTES;
        mbar_arrive_expect_tx(w_full, b1 + b2);
        for (uint32_t o = 0; o < b1; o += 16384) bulk_g2s(w1_base + o, p.w1_img + o, min(16384u, b1 - o), w_full);
        for (uint32_t o = 0; o < b2; o += 16384) bulk_g2s(w2_base + o, p.w2_img + o, min(16384u, b2 - o), w_full);
      }
      __syncwarp();
      mbar_wait(w_full, 0);
      uint32_t ph_a1 = 0;
      auto issue_conv1 = [&](int i) {                            // conv1d_1 GEMM of this CTA's i-th unit
        const int s1 = i & 1;
        mbar_wait(&acc1_empty[s1], (static_cast<uint32_t>(i >> 1) & 1u) ^ 1u);
        mbar_wait(a1_full, ph_a1); ph_a1 ^= 1u;
        tc_fence_after();
        const uint32_t d = tmem_base + static_cast<uint32_t>(s1) * acc1_cols;
        // kT: the weight image is the A operand (M = c0 channels), the window rows the B operand (N = 128 time steps)
        const uint32_t xa = kT ? w1 : a1, xb = kT ? a1 : w1;
        const uint32_t xa_s1 = kT ? w1_slab : (A_SLAB_BYTES >> 4), xb_s1 = kT ? (A_SLAB_BYTES >> 4) : w1_slab;
        if (elect_one()) {
          umma_f16_lo(d, xa, xb, idesc1, 0u);
          umma_f16_lo(d, xa + 2, xb + 2, idesc1, 1u);
          umma_f16_lo(d, xa + 4, xb + 4, idesc1, 1u);
          umma_f16_lo(d, xa + 6, xb + 6, idesc1, 1u);
          umma_f16_lo(d, xa + xa_s1, xb + xb_s1, idesc1, 1u);      // samples 64..79
          umma_commit(a1_empty);
          umma_commit(&acc1_full[s1]);
        }
        __syncwarp();
      };
      int i = 0, n2 = 0;
      if (u0 < u1) issue_conv1(0);
      int b = 0, g = 0, j = 0;
      if (u0 < u1) decode(u0, b, g, j);
      for (int unit = u0; unit < u1; ++unit, ++i, advance(b, g, j)) {
        if (unit + 1 < u1) issue_conv1(i + 1);                   // next tile's conv1d_1 runs under this tile's views
        for (int mem = p.vg.start[g]; mem < p.vg.start[g + 1]; ++mem, ++n2) {
          const int s2 = n2 & 1;
          const uint32_t ph2 = static_cast<uint32_t>(n2 >> 1) & 1u;
          mbar_wait(&acc2_empty[s2], ph2 ^ 1u);
          mbar_wait(&a2_full[s2], ph2);                          // A buffer n2 % 2 (filled in place by the middle warps)
          tc_fence_after();
          const uint32_t d = tmem_base + acc2_col0 + static_cast<uint32_t>(s2 * p.c1);
          uint32_t a = a2 + static_cast<uint32_t>(s2) * a2_buf_lo, w = w2;
          if (kProfile && (p.knockout & 8)) {
            umma_commit_elect(a2_empty0 + 8u * s2); umma_commit_elect(acc2_full0 + 8u * s2);
          } else if constexpr (kT) {
            // A2 is MN-major (see the middle warps): a K = 16 step spans two 8-channel groups = 2 SBO, a 64-channel slab 8 SBO
            a = umma_desc_lo_mn(smem_u32(a2_base), FUSE_A2_LBO) + static_cast<uint32_t>(s2) * a2_buf_lo;
            for (int kb = 0; kb < nkb2; ++kb, a += (8 * FUSE_A2_SBO) >> 4, w += w2_slab)
              umma_slab4_commit_mn(d, a, (2 * FUSE_A2_SBO) >> 4, umma_desc_hi_mn(FUSE_A2_SBO), w, idesc2 | (1u << 15), kb != 0 ? 1u : 0u,
                                   kb == nkb2 - 1 ? a2_empty0 + 8u * s2 : 0u, kb == nkb2 - 1 ? acc2_full0 + 8u * s2 : 0u);
          } else
          for (int kb = 0; kb < nkb2; ++kb, a += A_SLAB_BYTES >> 4, w += w2_slab)
            umma_slab4_commit(d, a, w, idesc2, kb != 0 ? 1u : 0u, kb == nkb2 - 1 ? a2_empty0 + 8u * s2 : 0u,
                              kb == nkb2 - 1 ? acc2_full0 + 8u * s2 : 0u);
        }
      }
    }
  } else {
    // =========================== waveform window producers ===========================
    using Conv1Producer = Conv1ProducerT<FUSE_PROD_THREADS>;
    const int ptid = tid - FUSE_PROD_WARP0 * 32;
    uint32_t pe = 0, buf = 0;
    float4 qd[Conv1Producer::QUADS];
    const float* x = nullptr;
    Conv1Tile cur{}, nxt{};
    auto describe = [&](int unit) {
      int b, g, j; decode(unit, b, g, j);
      x = p.wav + static_cast<size_t>(b) * L;
      return Conv1Producer::make(p.vg.shift[g], j * FUSE_ROWS, min(TILE_M, p.t1 - j * FUSE_ROWS));
    };
    int unit = u0;
    if (unit < u1) { cur = describe(unit); Conv1Producer::load(cur, x, ptid, qd); }
    for (; unit < u1; ++unit) {
      __half* win = reinterpret_cast<__half*>(win_base + buf * CONV1_WIN_BYTES);
      Conv1Producer::store(cur, ptid, qd, win);
      const int next = unit + 1;
      if (next < u1) { nxt = describe(next); Conv1Producer::load(nxt, x, ptid, qd); }
      asm volatile("bar.sync 1, %0;" ::"n"(FUSE_PROD_THREADS) : "memory");    // window complete
      mbar_wait(a1_empty, pe ^ 1u);
      Conv1Producer::fill_slabs(a1_base, ptid, win, cur.rows);
      fence_proxy_async_smem();
      mbar_arrive(a1_full);
      pe ^= 1u; buf ^= 1u;
      cur = nxt;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == FUSE_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------

// Keras kernel [K, N] (row-major, K = input channel / tap), column n scaled by col_scale[n]
// (the folded BatchNorm scale) -> pre-swizzled fp16 slabs [num_kb][N rows x 128 B]; element
// (k, n) lands in slab k/64, row n, chunk (k%64)/8.
void build_weight_image(const float* w, const float* col_scale, int K, int N, std::vector<__half>& img) {
  const int num_kb = (K + SLAB_K - 1) / SLAB_K;
  img.assign(static_cast<size_t>(num_kb) * N * SLAB_K, __float2half_rn(0.0f));
  for (int k = 0; k < K; ++k) {
    const int kb = k / SLAB_K, kk = k % SLAB_K;
    for (int n = 0; n < N; ++n) {
      const size_t byte = static_cast<size_t>(kb) * N * ROW_BYTES + swz_off(n, kk / 8) + (kk % 8) * 2;
      img[byte / 2] = __float2half_rn(w[static_cast<size_t>(k) * N + n] * col_scale[n]);
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// fp16 channels-last activation [d2][d1][c] (d2 = 1 for a plain [rows, c] matrix); box = [1][box_rows][64 ch]
int make_tensor_map(kws_handle* h, CUtensorMap* tm, const __half* act, int c, long long d1, long long d2,
                    int box_rows, bool swizzle128, int box_ch = SLAB_K) {
  static_assert(sizeof(CUtensorMap) == 128, "tensor map cache entry size");
  for (const auto& e : h->tmap_cache)
    if (e.act == act && e.c == c && e.d1 == d1 && e.d2 == d2 && e.box_rows == box_rows && e.swizzle == (swizzle128 ? 1 : 0) &&
        e.box_ch == box_ch) {
      memcpy(tm, e.map, sizeof(CUtensorMap));
      return KWS_OK;
    }
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(h, KWS_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint32_t rank = d2 > 1 ? 3 : 2;
  const cuuint64_t gdim[3] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(d1), static_cast<cuuint64_t>(d2)};
  const cuuint64_t gstride[2] = {static_cast<cuuint64_t>(c) * sizeof(__half),
                                 static_cast<cuuint64_t>(c) * sizeof(__half) * static_cast<cuuint64_t>(d1)};
  const cuuint32_t box[3] = {static_cast<cuuint32_t>(box_ch), static_cast<cuuint32_t>(box_rows), 1};
  const cuuint32_t estride[3] = {1, 1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<__half*>(act), gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, KWS_ECUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string(r));
  if (h->tmap_cache.size() >= 256) h->tmap_cache.clear();          // ragged batch sizes: start over rather than grow
  kws_handle::TmapEntry e{act, c, d1, d2, box_rows, swizzle128 ? 1 : 0, box_ch, {}};
  memcpy(e.map, tm, sizeof(CUtensorMap));
  h->tmap_cache.push_back(e);
  return KWS_OK;
}

template <int MODE>
int launch_tc_gemm(kws_handle* h, GemmParams& p, cudaStream_t st) {
  const bool conv1 = MODE == 0;
  if (p.cout % 32 || p.cout > 4 * 256) return fail(h, KWS_EUNSUPPORTED, "unsupported channel count for the tensor-core path");
  auto set_split = [&](int ns) {
    p.n_split = ns; p.ncta = p.cout / ns;
    p.n_halves = p.ncta > 256 ? 2 : 1;                          // split ncta into <= 256-wide instructions
    p.n_inst = p.ncta / p.n_halves;
    p.acc_stages = std::min(2, TMEM_COLS / p.ncta);
    return p.cout % ns == 0 && p.ncta % 32 == 0 && p.n_inst % 16 == 0 && p.n_inst <= 256 && p.ncta <= TMEM_COLS;
  };
  p.a_stage_bytes = conv1 ? 2 * A_SLAB_BYTES : A_SLAB_BYTES;
  p.out_bufs = 2;
  auto fits = [&](int a, int b, int r) {
    p.a_stages = a; p.b_stages = b; p.raw_stages = r;
    return static_cast<int>(smem_layout(p, conv1).total) <= SMEM_LIMIT;
  };
  bool chosen = false;
  if (conv1) {
    if (!set_split(1)) return fail(h, KWS_EUNSUPPORTED, "unsupported conv1d_1 width");
    const int blocks = p.num_kb * p.n_halves;
    p.b_resident = 1;
    chosen = fits(3, blocks, 0) || fits(2, blocks, 0);
  } else {
    // Wide layers (cout > 256) cannot double-buffer their accumulator in TMEM nor keep their weights
    // in shared memory, which serialises MMA and epilogue and re-streams up to 512 KB of weights per
    // tile from L2.  They are split column-wise over 2-4 adjacent CTAs instead: each CTA keeps its
    // slice of the weights resident for the whole launch and owns two accumulator stages; the A
    // operand of a row tile is then produced once per slice (its raw rows come from L2 after the
    // first CTA touched them).  Preference: smallest split with resident weights, deep raw ring, deep A ring.
    // (The producers fill the rings strictly in order, so any depth >= 2 is valid.)
    static const int force_a = [] { const char* e = getenv("KWS_A_STAGES"); return e ? atoi(e) : 0; }();       // A/B aids
    static const int force_ob = [] { const char* e = getenv("KWS_OUT_BUFS"); return e ? atoi(e) : 0; }();
    static const int force_r = [] { const char* e = getenv("KWS_RAW_STAGES"); return e ? atoi(e) : 0; }();
    auto fits_f = [&](int a, int b, int r) {
      if ((force_a && a != force_a) || (force_r && r != force_r)) return false;
      return fits(a, b, r);
    };
    const int r_min = MODE == 1 ? 3 : 2;
    auto search = [&](int ns_lo, int ns_hi, bool resident_only, int ob_lo) {
      for (int ns = ns_lo; ns <= ns_hi && !chosen; ++ns) {
        if (!set_split(ns)) continue;
        const int blocks = p.num_kb * p.n_halves;
        const int b_block = p.n_inst * ROW_BYTES;
        const int b_stream = std::max(2, std::min(4, 65536 / b_block));
        if (ob_lo == 2) {                                        // narrow layers
          // resident weights first (measured: 192->192 with resident weights and a 2-deep A ring beats streamed
          // weights with a 4-deep A ring by 8 %), then streamed weights with the deepest rings that fit
          // two epilogue warps per lane quarter alternate, so one store box each is enough -- except when the single
          // accumulator stage of a 320-512-column layer exposes the epilogue (measured: -7 % / -6 % on 256->320 and
          // 320->384 with two boxes, +15 % on the layers that lose ring depth to them); KWS_OUT_BUFS: A/B aid
          const int ob_first = force_ob ? force_ob : (p.acc_stages == 1 ? 2 : 1);
          for (int ob = ob_first == 2 ? 2 : 1; ob >= 1 && !chosen; --ob) {
            p.out_bufs = ob;
            if (blocks <= MAX_B_BLOCKS)
              for (int r = 6; r >= r_min && !chosen; --r)          // TMA prefetch depth matters more than A-ring depth
                for (int a = 4; a >= 2 && !chosen; --a)
                  if (fits_f(a, blocks, r)) { p.b_resident = 1; chosen = true; }
            for (int a = 4; a >= 2 && !chosen; --a)
              for (int res = 1; res >= (resident_only ? 1 : 0) && !chosen; --res) {
                if (res && blocks > MAX_B_BLOCKS) continue;
                for (int bs = res ? blocks : b_stream; bs >= (res ? blocks : 2) && !chosen; --bs)
                  for (int r = 6; r >= 2 && !chosen; --r)
                    if (fits_f(a, bs, r)) { p.b_resident = res; chosen = true; }
              }
          }
        } else {                                                 // split layers: deep raw ring (HBM/L2 latency) first
          for (int res = 1; res >= (resident_only ? 1 : 0) && !chosen; --res) {
            if (res && blocks > MAX_B_BLOCKS) continue;
            for (int r = 6; r >= 2 && !chosen; --r)
              for (int ob = 1; ob >= 1 && !chosen; --ob) {
                p.out_bufs = ob;
                for (int a = 4; a >= 2 && !chosen; --a)
                  for (int bs = res ? blocks : b_stream; bs >= (res ? blocks : 2) && !chosen; --bs)
                    if (fits_f(a, bs, r)) { p.b_resident = res; chosen = true; }
              }
          }
        }
      }
    };
    // Column split of the wide layers (cout > 256) over 2 adjacent CTAs: resident weight slices and two accumulator
    // stages, but the A operand is produced once per slice.  It paid (-12 / -7 / -6 % on 320->320, 384->384, 512->512)
    // while the MMA issue and the epilogue were slow; with this round's pipeline the unsplit form -- streamed weights,
    // one accumulator stage, two store boxes per epilogue warp -- is 2-7 % faster on those three layers and the
    // stride-2 ones, so the split is off by default (KWS_MAX_SPLIT=2, plus KWS_SPLIT_S2=1 for the stride-2 layers, re-enables it for A/B runs).
    static const int max_split = [] { const char* e = getenv("KWS_MAX_SPLIT"); return e ? atoi(e) : 1; }();   // A/B aid
    static const bool split_s2 = [] { const char* e = getenv("KWS_SPLIT_S2"); return e && e[0] == '1'; }();           // A/B aid
    static const bool no_two_pass = [] { const char* e = getenv("KWS_NO_TWO_PASS"); return e && e[0] == '1'; }();   // A/B aid
    p.two_pass = 0;
    if (!no_two_pass && set_split(1) && p.n_halves == 2 && p.num_kb <= MAX_STAGES) {
      // every K slab of a tile resident in the A ring, MMAs half by half (see GemmParams::two_pass): needs
      // a_stages == num_kb; streamed weights (even ring depth is not required: one issuer warp), one store box
      const int b_block = p.n_inst * ROW_BYTES;
      p.out_bufs = 1;
      for (int r = 4; r >= 2 && !chosen; --r)
        for (int bs = std::max(2, std::min(4, 65536 / b_block)); bs >= 2 && !chosen; --bs)
          if (fits_f(p.num_kb, bs, r)) { p.b_resident = 0; p.two_pass = 1; chosen = true; }
    }
    if (chosen) {
    } else if (p.cout <= 256 || max_split < 2 || (MODE != 1 && !split_s2)) {
      search(1, 1, false, 2);                                    // weights resident when they fit, else streamed
    } else {
      search(2, max_split, true, 1);
      if (!chosen) search(2, max_split, false, 1);
    }
  }
  if (!chosen) return fail(h, KWS_EUNSUPPORTED, "layer does not fit in shared memory");
  const SmemLayout lay = smem_layout(p, conv1);
  if (!(h->smem_attr_done & (2u << MODE))) {
    KWS_CUDA(h, cudaFuncSetAttribute(tc_gemm_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    h->smem_attr_done |= 2u << MODE;
  }
  const int grid = std::min(p.num_tiles, h->num_sms / p.n_split) * p.n_split;
  if (grid <= 0) return KWS_OK;
  if (!conv1 && (p.two_pass ? p.n_inst : p.ncta) % 64) {         // 32-column tail boxes of this launch's output
    const int rc = make_tensor_map(h, &p.tmap_tail, p.out_act, p.cout, p.out_rows, 1, 32, false, 32);
    if (rc) return rc;
  }
  static const int knockout = [] {
    const char* e = getenv("KWS_KNOCKOUT");
    if ((e || getenv("KWS_TRACE")) && !kProfile)
      fprintf(stderr, "kws: KWS_KNOCKOUT / KWS_TRACE need a profiling build (KWS_PROFILE_BUILD=1 python -m speech_recognition_b200.build --force)\n");
    return e && kProfile ? atoi(e) : 0;
  }();
  p.knockout = knockout;
  static const int trace_block = [] { const char* e = getenv("KWS_TRACE"); return e ? atoi(e) : -1; }();
  static int trace_launches = 0;
  const bool tracing = kProfile && MODE != 0 && trace_block == p.block_index && ++trace_launches == 2;   // a warmed-up launch
  p.trace = nullptr;
  if (tracing) {
    cudaMalloc(&p.trace, sizeof(unsigned long long) * TRACE_EVENTS * TRACE_ROLES);
    cudaMemset(p.trace, 0, sizeof(unsigned long long) * TRACE_EVENTS * TRACE_ROLES);
  }
  KWS_T0(h, MODE == 0 ? KC_CONV1 : KC_BLOCK0 + p.block_index, st);
  tc_gemm_kernel<MODE><<<grid, Roles<MODE>::THREADS, lay.total, st>>>(p);
  KWS_T1(h, st);
  if (tracing) {                                                 // profiling aid only: synchronises and writes a file
    std::vector<unsigned long long> ev(static_cast<size_t>(TRACE_EVENTS) * TRACE_ROLES);
    cudaDeviceSynchronize();
    cudaMemcpy(ev.data(), p.trace, ev.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(p.trace);
    if (FILE* f = fopen("gpurun_out/trace.bin", "wb")) {
      fwrite(ev.data(), sizeof(unsigned long long), ev.size(), f);
      fclose(f);
    }
    fprintf(stderr, "kws trace: block %d cin %d cout %d stride %d num_kb %d a_stages %d raw_stages %d b_resident %d b_stages %d "
                    "n_split %d acc_stages %d tiles %d grid %d\n",
            p.block_index, p.cin, p.cout, p.stride, p.num_kb, p.a_stages, p.raw_stages, p.b_resident, p.b_stages, p.n_split,
            p.acc_stages, p.num_tiles, grid);
  }
  if (debug_sync() && cudaDeviceSynchronize() != cudaSuccess)
    return fail(h, KWS_ECUDA, "tc_gemm_kernel<" + std::to_string(MODE) + "> cin " + std::to_string(p.cin) + " cout " +
                                  std::to_string(p.cout) + " rows_out " + std::to_string(p.rows_out) + " stages a/b/raw " +
                                  std::to_string(p.a_stages) + "/" + std::to_string(p.b_stages) + "/" +
                                  std::to_string(p.raw_stages) + " resident " + std::to_string(p.b_resident) + " split " +
                                  std::to_string(p.n_split) + " out_bufs " + std::to_string(p.out_bufs) + ": " +
                                  cudaGetErrorString(cudaGetLastError()));
  KWS_LAUNCH_CHECK(h);
  return KWS_OK;
}

int launch_conv1_block1(kws_handle* h, Model& m, const float* wav, int nb, const ViewGroups& vg, int V, __half* out,
                        cudaStream_t st) {
  const LayerDesc& d = m.layers[0];
  FusedParams p{};
  p.wav = wav; p.vg = vg; p.n_views = V;
  p.w1_img = reinterpret_cast<const uint8_t*>(m.tc_conv1);
  p.w2_img = reinterpret_cast<const uint8_t*>(m.tc_pw[0]);
  p.dw_h = m.tc_dw[0];
  p.shift1 = m.tc_shift0; p.shift2 = m.bn_shift[1];
  p.c0 = m.c0_tc; p.c1 = d.cout; p.t1 = m.t0; p.t2 = d.t_out;
  p.blocks_per_view = (p.t2 + FUSE_ROWS - 1) / FUSE_ROWS;
  p.num_units = nb * vg.n_groups * p.blocks_per_view;
  const int rows = nb * V;
  int rc = make_tensor_map(h, &p.tmap_out, out, p.c1, p.t2, std::max(rows, 2), 32, true);
  if (rc) return rc;
  rc = make_tensor_map(h, &p.tmap_out30, out, p.c1, p.t2, std::max(rows, 2), 30, true);
  if (rc) return rc;
  const FusedSmem lay = fused_smem(p.c0, p.c1);
  if (static_cast<int>(lay.total) > SMEM_LIMIT) return fail(h, KWS_EUNSUPPORTED, "fused conv1d_1 + block 1 does not fit in shared memory");
  if (!(h->smem_attr_done & 16u)) {
    KWS_CUDA(h, cudaFuncSetAttribute(conv1_block1_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    KWS_CUDA(h, cudaFuncSetAttribute(conv1_block1_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    h->smem_attr_done |= 16u;
  }
  const int grid = std::min(p.num_units, h->num_sms);
  if (grid <= 0) return KWS_OK;
  // c0 == 128: the transposed form (see the kernel); KWS_FUSE_V1=1 keeps the r01 row-major form for A/B runs
  static const bool v1 = [] { const char* e = getenv("KWS_FUSE_V1"); return e && e[0] == '1'; }();
  const bool transposed = !v1 && p.c0 == TILE_M;
  static const int fknock = [] { const char* e = getenv("KWS_FKNOCK"); return e && kProfile ? atoi(e) : 0; }();
  p.knockout = fknock;
  KWS_T0(h, KC_CONV1, st);
  if (transposed) conv1_block1_kernel<true><<<grid, FuseRoles<true>::THREADS, lay.total, st>>>(p);
  else conv1_block1_kernel<false><<<grid, FuseRoles<false>::THREADS, lay.total, st>>>(p);
  KWS_T1(h, st);
  if (debug_sync() && cudaDeviceSynchronize() != cudaSuccess)
    return fail(h, KWS_ECUDA, std::string("conv1_block1_kernel: ") + cudaGetErrorString(cudaGetLastError()));
  KWS_LAUNCH_CHECK(h);
  return KWS_OK;
}

// conv1d_1 and block 1 run as one kernel when the shapes allow it (both shipped architectures);
// KWS_NO_FUSE=1 keeps them separate (A/B and layer-0 debugging).
bool fuse_conv1_block1(const Model& m) {
  static const bool off = [] { const char* e = getenv("KWS_NO_FUSE"); return e && e[0] == '1'; }();
  const LayerDesc& d = m.layers[0];
  return !off && d.stride == 1 && d.cin == m.c0 && m.c0_tc <= 128 && d.cout % 64 == 0 &&
         2 * m.c0_tc + 2 * d.cout <= TMEM_COLS && d.t_out == m.t0 - 2;
}

}  // namespace

int model_build_tc(kws_handle* h, Model& m, const std::vector<std::vector<float>>& pw_host,
                   const std::vector<float>& conv1_host, const std::vector<std::vector<float>>& dw_host,
                   const std::vector<std::vector<float>>& scales) {
  // conv1: fold the 3 overlapping patches into 80 taps: W80[u, co] = sum_f W[f, u - 20 f, co].  The tensor-core kernels
  // work on 64-channel K slabs, so a conv1d_1 narrower than that (conv_1d_time_sliced: 32 filters) is padded to c0_tc
  // with zero filters (zero BN shift: their activation is ReLU6(0) = 0) and block 1 gets zero taps / zero weight rows for them.
  const int c0 = m.c0, c0p = m.c0_tc;
  std::vector<float> w80(static_cast<size_t>(CONV1_K) * c0p, 0.0f), scale0(c0p, 1.0f);
  std::copy(scales[0].begin(), scales[0].end(), scale0.begin());
  for (int f = 0; f < 3; ++f)
    for (int i = 0; i < 40; ++i)
      for (int co = 0; co < c0; ++co)
        w80[static_cast<size_t>(20 * f + i) * c0p + co] += conv1_host[static_cast<size_t>(f * 40 + i) * c0 + co];
  std::vector<__half> all, img;
  std::vector<size_t> offs, dw_offs;
  auto align_1k = [&]() { while (all.size() % 512) all.push_back(__float2half_rn(0.0f)); };   // 1024-byte aligned
  build_weight_image(w80.data(), scale0.data(), CONV1_K, c0p, img);
  offs.push_back(all.size()); all.insert(all.end(), img.begin(), img.end());
  for (int i = 0; i < m.n_blocks; ++i) {
    const int cin = m.layers[i].cin, cinp = i == 0 ? c0p : cin, cout = m.layers[i].cout;
    if (cinp == cin) {
      build_weight_image(pw_host[i].data(), scales[i + 1].data(), cin, cout, img);
    } else {
      std::vector<float> padded(static_cast<size_t>(cinp) * cout, 0.0f);
      std::copy(pw_host[i].begin(), pw_host[i].end(), padded.begin());      // rows cin .. cinp - 1 stay zero
      build_weight_image(padded.data(), scales[i + 1].data(), cinp, cout, img);
    }
    align_1k();
    offs.push_back(all.size()); all.insert(all.end(), img.begin(), img.end());
  }
  for (int i = 0; i < m.n_blocks; ++i) {                       // depthwise taps [3][cin] as fp16
    const int cin = m.layers[i].cin, cinp = i == 0 ? c0p : cin;
    align_1k();
    dw_offs.push_back(all.size());
    for (int j = 0; j < 3; ++j)
      for (int c = 0; c < cinp; ++c) all.push_back(__float2half_rn(c < cin ? dw_host[i][static_cast<size_t>(j) * cin + c] : 0.0f));
  }
  KWS_CUDA(h, cudaMalloc(&m.tc_blob, all.size() * sizeof(__half)));
  KWS_CUDA(h, cudaMemcpy(m.tc_blob, all.data(), all.size() * sizeof(__half), cudaMemcpyHostToDevice));
  __half* base = static_cast<__half*>(m.tc_blob);
  m.tc_conv1 = base + offs[0];
  for (int i = 0; i < m.n_blocks; ++i) { m.tc_pw[i] = base + offs[i + 1]; m.tc_dw[i] = base + dw_offs[i]; }
  return KWS_OK;
}

int launch_forward_tc(kws_handle* h, Model& m, const float* wav, int B, const ViewTable& vt,
                      float* probs_mean, int32_t* argmax, cudaStream_t st, int dbg_layer, float* dbg_out) {
  const int V = vt.n;
  ViewGroups vg{};
  {
    bool used[KWS_MAX_VIEWS] = {};
    int mem = 0;
    for (int v = 0; v < V; ++v) {
      if (used[v]) continue;
      int sm = vt.shift[v] % L; if (sm < 0) sm += L;
      vg.shift[vg.n_groups] = sm;
      vg.start[vg.n_groups] = mem;
      for (int u = v; u < V; ++u) {
        int su = vt.shift[u] % L; if (su < 0) su += L;
        if (!used[u] && su == sm) { used[u] = true; vg.view[mem] = u; vg.gain[mem] = vt.gain[u]; ++mem; }
      }
      vg.start[++vg.n_groups] = mem;
    }
  }
  const int clips_per_chunk = std::max(1, h->max_rows / V);
  const size_t need = static_cast<size_t>(clips_per_chunk) * V * m.max_act_elems * sizeof(__half);
  if (h->act_bytes < need) {
    h->tmap_cache.clear();
    for (int i = 0; i < 2; ++i) {
      if (h->act[i]) cudaFree(h->act[i]);
      h->act[i] = nullptr;
      KWS_CUDA(h, cudaMalloc(&h->act[i], need));
    }
    h->act_bytes = need;
  }
  for (int b0 = 0; b0 < B; b0 += clips_per_chunk) {
    const int nb = std::min(clips_per_chunk, B - b0);
    const int rows = nb * V;
    __half* cur = static_cast<__half*>(h->act[0]);
    __half* nxt = static_cast<__half*>(h->act[1]);
    const bool fused = dbg_layer != 0 && h->fuse_conv1_block1 && fuse_conv1_block1(m);
    if (fused) {
      int rc = launch_conv1_block1(h, m, wav + static_cast<size_t>(b0) * L, nb, vg, V, cur, st);
      if (rc) return rc;
      if (dbg_layer == 1)
        return launch_to_float(h, cur, true, dbg_out, static_cast<size_t>(rows) * m.layers[0].t_out * m.layers[0].cout, st);
    } else {
      GemmParams p{};
      p.wav = wav + static_cast<size_t>(b0) * L; p.n_views = V;
      p.vg = vg;
      p.w_img = reinterpret_cast<const uint8_t*>(m.tc_conv1);
      p.shift = m.tc_shift0;
      p.cin = CONV1_K; p.cout = m.c0_tc; p.t_out = m.t0; p.rows_out = rows * m.t0;
      p.tiles_per_group = (m.t0 + TILE_M - 1) / TILE_M;
      p.num_tiles = nb * vg.n_groups * p.tiles_per_group;
      p.num_kb = 2; p.last_ksteps = 1;
      int rc = make_tensor_map(h, &p.tmap_out, cur, m.c0_tc, m.t0, std::max(rows, 2), 32, true);   // 3-D: clip at t0
      if (rc) return rc;
      rc = launch_tc_gemm<0>(h, p, st);
      if (rc) return rc;
    }
    if (dbg_layer == 0) return launch_to_float(h, cur, true, dbg_out, static_cast<size_t>(rows) * m.t0 * m.c0_tc, st);   // c0_tc channels (zero padded)
    for (int i = fused ? 1 : 0; i < m.n_blocks; ++i) {
      LayerDesc d = m.layers[i];
      if (i == 0) d.cin = m.c0_tc;                               // zero-padded conv1d_1 channels (zero taps, zero weight rows)
      if (d.cin % SLAB_K) return fail(h, KWS_EUNSUPPORTED, "channel count must be a multiple of 64");
      GemmParams p{};
      p.dw_h = m.tc_dw[i];
      p.w_img = reinterpret_cast<const uint8_t*>(m.tc_pw[i]);
      p.shift = m.bn_shift[i + 1];
      p.cin = d.cin; p.cout = d.cout; p.stride = d.stride; p.pad_left = d.pad_left;
      p.t_in = d.t_in; p.t_out = d.t_out; p.rows_out = rows * d.t_out;
      p.num_tiles = (p.rows_out + TILE_M - 1) / TILE_M;
      p.num_kb = d.cin / SLAB_K; p.last_ksteps = 4;
      // extent of a tile inside the raw box: 127 rows * stride + one jump per clip boundary + 3 taps
      // (stride 2: a boundary advances by t_in - 2 t_out + 2 <= 2 rows, no more than a normal step)
      const int boundaries = (TILE_M + d.t_out - 2) / d.t_out;
      const int extent = (TILE_M - 1) * d.stride + (d.stride == 1 ? 2 * boundaries : 0) + 3;
      p.n_boxes = d.stride == 1 ? 1 : 2;
      p.box_rows = ((extent + p.n_boxes - 1) / p.n_boxes + 7) & ~7;
      if (p.box_rows > 256) return fail(h, KWS_EUNSUPPORTED, "tile does not fit the raw activation box");
      p.raw_stage_bytes = p.box_rows * p.n_boxes * ROW_BYTES;
      p.t_out_magic = ((1ull << 40) + d.t_out - 1) / d.t_out;
      int rc = make_tensor_map(h, &p.tmap_in, cur, d.cin, static_cast<long long>(rows) * d.t_in, 1, p.box_rows, false);
      if (rc) return rc;
      rc = make_tensor_map(h, &p.tmap_out, nxt, d.cout, static_cast<long long>(rows) * d.t_out, 1, 32, true);
      if (rc) return rc;
      p.block_index = i; p.out_act = nxt; p.out_rows = static_cast<long long>(rows) * d.t_out;
      p.in_act = cur; p.rows_in = static_cast<long long>(rows) * d.t_in;
      static const int prefetch_tiles = [] { const char* e = getenv("KWS_PREFETCH_TILES"); return e ? atoi(e) : 0; }();   // A/B aid
      p.prefetch_tiles = prefetch_tiles;
      rc = d.stride == 1 ? launch_tc_gemm<1>(h, p, st) : launch_tc_gemm<2>(h, p, st);
      if (rc) return rc;
      std::swap(cur, nxt);
      if (dbg_layer == i + 1)
        return launch_to_float(h, cur, true, dbg_out, static_cast<size_t>(rows) * d.t_out * d.cout, st);
    }
    int rc = launch_head(h, m, cur, /*act_half=*/true, nb, V,
                         probs_mean ? probs_mean + static_cast<size_t>(b0) * m.classes : nullptr,
                         argmax ? argmax + b0 : nullptr, st);
    if (rc) return rc;
  }
  return KWS_OK;
}

}  // namespace kws
