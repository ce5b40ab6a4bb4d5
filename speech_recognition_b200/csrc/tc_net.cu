// Tensor-core tier of the Depthwise1D network forward (reference model.py:34-52,67-76,805-817).
//
// One warp-specialised, persistent kernel template serves both GEMM-shaped layer types:
//   * K5 slice_conv1 : overlapping_time_slice_stack (k40 s20 SAME) + Conv1D(C0,k3,s2) + BN + ReLU6.
//                      The 3 overlapping patches of a row cover 80 consecutive samples, so the
//                      layer is a K=80 contraction against taps pre-summed on the host
//                      (W80[u] = sum_f W[f, u-20f]); the TTA view (np.roll + gain,
//                      make_submission.py:126-130) is applied while the waveform window is staged.
//   * K6 dw_pw_block : depthwise k3 FIR (CUDA cores, fp32) as the A-operand producer of the
//                      pointwise GEMM + BN + ReLU6 epilogue.
// Roles (448 threads, 1 CTA / SM):
//   warps 0-3  epilogue : tcgen05.ld accumulator -> acc*scale+shift -> ReLU6 -> fp16 -> global
//   warp  4    MMA      : one thread issues tcgen05.mma (A,B from swizzled smem, D in TMEM)
//   warp  5    B loader : cp.async.bulk of pre-swizzled fp16 weight slabs (resident when they fit)
//   warps 6-13 A producers
// Pipelines: A ring (producers <-> MMA), B ring (loader <-> MMA), accumulator stages in TMEM
// (MMA <-> epilogue), all on mbarriers; tcgen05.commit releases smem slots / publishes accumulators.
// Operands are fp16 (same tensor rate as bf16, 3 more mantissa bits; ReLU6 bounds activations to
// [0,6] so the range is safe), accumulation is fp32 in TMEM, BN scale/shift stay fp32.
#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "tc_common.cuh"

namespace kws {

using namespace tc;

namespace {

constexpr int NUM_EPI_WARPS = 4;
constexpr int NUM_PROD_WARPS = 8;
constexpr int NUM_PROD_THREADS = NUM_PROD_WARPS * 32;           // 256
constexpr int MMA_WARP = NUM_EPI_WARPS;                          // 4
constexpr int LOAD_WARP = NUM_EPI_WARPS + 1;                     // 5
constexpr int PROD_WARP0 = NUM_EPI_WARPS + 2;                    // 6
constexpr int TC_THREADS = 32 * (NUM_EPI_WARPS + 2 + NUM_PROD_WARPS);   // 448
constexpr int MAX_STAGES = 8;
constexpr int TMEM_COLS = 512;
constexpr int SMEM_LIMIT = 227 * 1024;

constexpr int CONV1_K = 80;                                      // samples per output row
constexpr int CONV1_ROW_HOP = 40;
constexpr int CONV1_WIN = CONV1_ROW_HOP * (TILE_M - 1) + CONV1_K;   // 5160 staged samples per tile

struct GemmParams {
  // A-side sources
  const __half* act_in;     // dw_pw: previous activation [rows_in, cin] fp16
  const float* wav;         // conv1: waveforms [B, 16000] fp32
  const float* dw;          // dw_pw: depthwise taps [3][cin] fp32
  ViewTable vt;             // conv1: TTA views
  int n_views;
  // B side: pre-swizzled fp16 slabs, slab kb at w_img + kb * cout * 128
  const uint8_t* w_img;
  // epilogue
  const float* scale;
  const float* shift;
  __half* out;              // [rows_out, cout] fp16
  // shapes
  int cin, cout, stride, pad_left, t_in, t_out;
  int rows_out;             // valid output rows
  int num_tiles;
  int tiles_per_group;      // conv1: tiles per clip-view (4); dw_pw: unused
  int num_kb;               // K slabs
  int last_ksteps;          // K=16 steps in the last slab (4, or 1 for conv1)
  int a_stages, b_stages, acc_stages, b_resident;
  int n_inst, n_halves;     // cout = n_inst * n_halves, n_inst <= 256
  int a_stage_bytes;        // dw_pw: 16 KB; conv1: 32 KB (both slabs of a tile)
};

struct SmemLayout {
  uint32_t a_off, b_off, aux_off, bar_off, total;
};

__host__ __device__ inline SmemLayout smem_layout(const GemmParams& p, bool conv1) {
  SmemLayout s;
  uint32_t o = 0;
  s.a_off = o; o += static_cast<uint32_t>(p.a_stages) * p.a_stage_bytes;
  s.b_off = o; o += static_cast<uint32_t>(p.b_stages) * p.cout * ROW_BYTES;
  s.aux_off = o;
  // aux: scale[cout] shift[cout] fp32, then dw taps [3*cin] fp32 (dw_pw) or the staged fp16 window (conv1)
  o += 2u * p.cout * 4u;
  o += conv1 ? static_cast<uint32_t>(((CONV1_WIN + 8) * 2 + 15) & ~15) : 3u * p.cin * 4u;
  o = (o + 15u) & ~15u;
  s.bar_off = o; o += (4 * MAX_STAGES + 4) * 8 + 16;
  s.total = o + 1024;        // slack for the manual 1024-byte alignment of the base
  return s;
}

// ------------------------------------------------------------------------------------------------
// A producers
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ void unpack8(const uint4& v, float* x) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    x[2 * i] = f.x; x[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* x) {
  uint4 v;
  __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
  return v;
}

// Depthwise producer: thread = (16-byte channel chunk c of the 64-channel slab, group g of 4
// consecutive output rows).  STRIDE is the depthwise stride (1 VALID / 2 SAME).
template <int STRIDE>
struct DwProducer {
  int c, g;
  int in_row0[4];            // index into act_in rows of tap 0 for each of the 4 output rows (may be <0)
  int t0[4];                 // t*stride - pad_left (time index of tap 0 inside the clip)
  bool valid[4];
  bool same_clip;

  __device__ __forceinline__ void begin_tile(const GemmParams& p, int tile, int ptid) {
    c = ptid & 7; g = ptid >> 3;
    const int m0 = tile * TILE_M + 4 * g;
    int rv0 = -1;
    same_clip = true;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + i;
      valid[i] = m < p.rows_out;
      const int mm = valid[i] ? m : 0;
      const int rv = mm / p.t_out, t = mm - rv * p.t_out;
      t0[i] = t * STRIDE - p.pad_left;
      in_row0[i] = rv * p.t_in + t0[i];
      if (i == 0) rv0 = rv;
      if (rv != rv0 || !valid[i]) same_clip = false;
    }
  }

  __device__ __forceinline__ void produce(const GemmParams& p, uint8_t* slab, int kb, const float* s_dw) const {
    const int ch0 = kb * SLAB_K + c * 8;
    float w[3][8];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float4 a = *reinterpret_cast<const float4*>(s_dw + j * p.cin + ch0);
      const float4 b = *reinterpret_cast<const float4*>(s_dw + j * p.cin + ch0 + 4);
      w[j][0] = a.x; w[j][1] = a.y; w[j][2] = a.z; w[j][3] = a.w;
      w[j][4] = b.x; w[j][5] = b.y; w[j][6] = b.z; w[j][7] = b.w;
    }
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[i][e] = 0.0f;
    const __half* base = p.act_in + ch0;
    if (same_clip) {
      // sliding window: input rows q = 0 .. 3*STRIDE+2 relative to tap 0 of output row 0
      constexpr int NQ = 3 * STRIDE + 3;
      uint4 v[NQ];
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int ti = t0[0] + q;
        v[q] = make_uint4(0u, 0u, 0u, 0u);
        if (ti >= 0 && ti < p.t_in)
          v[q] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(in_row0[0] + q) * p.cin));
      }
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        float x[8];
        unpack8(v[q], x);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = q - i * STRIDE;
          if (j >= 0 && j < 3) {
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[i][e] = fmaf(w[j][e], x[e], acc[i][e]);
          }
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (!valid[i]) continue;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int ti = t0[i] + j;
          if (ti < 0 || ti >= p.t_in) continue;
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + static_cast<size_t>(in_row0[i] + j) * p.cin));
          float x[8];
          unpack8(v, x);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[i][e] = fmaf(w[j][e], x[e], acc[i][e]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t r = 4 * g + i;
      *reinterpret_cast<uint4*>(slab + swz_off(r, c)) = pack8(acc[i]);
    }
  }
};

// conv1 producer: stage the 5160-sample window of (clip-view, row block) in smem as fp16 with the
// TTA view applied (coalesced scalar loads, circular index), then copy 16-byte chunks into the
// swizzled slabs: row r = samples [40 r, 40 r + 80) of the window; slab 0 = first 64, slab 1 = last 16.
struct Conv1Producer {
  __device__ __forceinline__ static void stage_window(const GemmParams& p, int tile, int ptid, __half* s_win) {
    const int rv = tile / p.tiles_per_group, jb = tile - rv * p.tiles_per_group;
    const int b = rv / p.n_views, v = rv - b * p.n_views;
    const int shift = p.vt.shift[v];
    const float gain = p.vt.gain[v];
    const float* x = p.wav + static_cast<size_t>(b) * L;
    const int p_start = CONV1_ROW_HOP * TILE_M * jb - 10;     // patch stack pads 10 samples on the left
    int sm = shift % L; if (sm < 0) sm += L;
    for (int i = ptid; i < CONV1_WIN; i += NUM_PROD_THREADS) {
      const int ps = p_start + i;
      float val = 0.0f;
      if (ps >= 0 && ps < L) {
        int src = ps - sm; if (src < 0) src += L;              // np.roll(x, shift)[ps]
        val = __fmul_rn(gain, __ldg(&x[src]));
      }
      s_win[i] = __float2half_rn(val);
    }
  }
  __device__ __forceinline__ static void fill_slabs(uint8_t* stage, int ptid, const __half* s_win) {
    // 128 rows x 10 chunks (8 in slab 0, 2 in slab 1)
    for (int task = ptid; task < TILE_M * 10; task += NUM_PROD_THREADS) {
      const int r = task / 10, ch = task - r * 10;
      const uint4 v = *reinterpret_cast<const uint4*>(s_win + CONV1_ROW_HOP * r + 8 * ch);
      uint8_t* slab = stage + (ch < 8 ? 0 : A_SLAB_BYTES);
      *reinterpret_cast<uint4*>(slab + swz_off(r, ch < 8 ? ch : ch - 8)) = v;
    }
  }
};

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int MODE>   // 0 = conv1, 1 = dw_pw stride 1, 2 = dw_pw stride 2
__global__ void __launch_bounds__(TC_THREADS, 1) tc_gemm_kernel(const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr bool kConv1 = (MODE == 0);
  const SmemLayout lay = smem_layout(p, kConv1);
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_base = smem + lay.a_off;
  uint8_t* b_base = smem + lay.b_off;
  float* s_scale = reinterpret_cast<float*>(smem + lay.aux_off);
  float* s_shift = s_scale + p.cout;
  float* s_dw = s_shift + p.cout;                              // dw_pw
  __half* s_win = reinterpret_cast<__half*>(s_shift + p.cout); // conv1
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bar_off);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + MAX_STAGES;
  uint64_t* b_full = bars + 2 * MAX_STAGES;
  uint64_t* b_empty = bars + 3 * MAX_STAGES;
  uint64_t* acc_full = bars + 4 * MAX_STAGES;
  uint64_t* acc_empty = bars + 4 * MAX_STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4 * MAX_STAGES + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup ----
  for (int i = tid; i < p.cout; i += TC_THREADS) { s_scale[i] = p.scale[i]; s_shift[i] = p.shift[i]; }
  if (!kConv1)
    for (int i = tid; i < 3 * p.cin; i += TC_THREADS) s_dw[i] = p.dw[i];
  if (warp == MMA_WARP) {
    if (lane == 0) {
      for (int i = 0; i < MAX_STAGES; ++i) {
        mbar_init(&a_full[i], NUM_PROD_THREADS);
        mbar_init(&a_empty[i], 1);
        mbar_init(&b_full[i], 1);
        mbar_init(&b_empty[i], 1);
      }
      for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], NUM_EPI_WARPS * 32); }
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
  const uint32_t b_slab_bytes = static_cast<uint32_t>(p.cout) * ROW_BYTES;

  if (warp < NUM_EPI_WARPS) {
    // =========================== epilogue ===========================
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      // output row of this thread
      const int r = warp * 32 + lane;
      long long orow;
      bool ok;
      if (kConv1) {
        const int rv = tile / p.tiles_per_group, jb = tile - rv * p.tiles_per_group;
        const int j = jb * TILE_M + r;
        ok = j < p.t_out;
        orow = static_cast<long long>(rv) * p.t_out + j;
      } else {
        orow = static_cast<long long>(tile) * TILE_M + r;
        ok = orow < p.rows_out;
      }
      __half* optr = p.out + orow * p.cout;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(acc * p.cout);
      for (int c0 = 0; c0 < p.cout; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
        float y[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float a = __uint_as_float(v[e]);
          y[e] = fminf(fmaxf(fmaf(a, s_scale[c0 + e], s_shift[c0 + e]), 0.0f), 6.0f);   // BN + ReLU6
        }
        if (ok) {
          *reinterpret_cast<uint4*>(optr + c0) = pack8(y);
          *reinterpret_cast<uint4*>(optr + c0 + 8) = pack8(y + 8);
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[acc]);
      if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp == MMA_WARP) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(TILE_M, p.n_inst, /*fp16*/ 0);
      int sa = 0; uint32_t pa = 0; int sb = 0; uint32_t pb = 0; int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d0 = tmem_base + static_cast<uint32_t>(acc * p.cout);
        if (kConv1) {
          mbar_wait(&a_full[sa], pa);                          // one stage = both slabs of the tile
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          uint32_t a_addr;
          if (kConv1) {
            a_addr = smem_u32(a_base + sa * p.a_stage_bytes + kb * A_SLAB_BYTES);
          } else {
            mbar_wait(&a_full[sa], pa);
            a_addr = smem_u32(a_base + sa * p.a_stage_bytes);
          }
          uint32_t b_addr;
          if (p.b_resident) {
            mbar_wait(&b_full[kb], 0);
            b_addr = smem_u32(b_base + kb * b_slab_bytes);
          } else {
            mbar_wait(&b_full[sb], pb);
            b_addr = smem_u32(b_base + sb * b_slab_bytes);
          }
          tc_fence_after();
          const int ksteps = (kb == p.num_kb - 1) ? p.last_ksteps : 4;
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t adesc = umma_desc_sw128(a_addr + ks * 32);
            for (int nh = 0; nh < p.n_halves; ++nh) {
              const uint64_t bdesc = umma_desc_sw128(b_addr + nh * p.n_inst * ROW_BYTES + ks * 32);
              umma_f16(d0 + nh * p.n_inst, adesc, bdesc, idesc, (kb | ks) != 0 ? 1u : 0u);
            }
          }
          if (!kConv1) {
            umma_commit(&a_empty[sa]);
            if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
          }
          if (!p.b_resident) {
            umma_commit(&b_empty[sb]);
            if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
          }
        }
        if (kConv1) {
          umma_commit(&a_empty[sa]);
          if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
        }
        umma_commit(&acc_full[acc]);
        if (++acc == p.acc_stages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp == LOAD_WARP) {
    // =========================== weight loader ===========================
    if (lane == 0) {
      if (p.b_resident) {
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_arrive_expect_tx(&b_full[kb], b_slab_bytes);
          for (uint32_t o = 0; o < b_slab_bytes; o += 16384)
            bulk_g2s(b_base + kb * b_slab_bytes + o, p.w_img + static_cast<size_t>(kb) * b_slab_bytes + o,
                     min(16384u, b_slab_bytes - o), &b_full[kb]);
        }
      } else {
        int sb = 0; uint32_t pb = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(&b_empty[sb], pb ^ 1);
            mbar_arrive_expect_tx(&b_full[sb], b_slab_bytes);
            for (uint32_t o = 0; o < b_slab_bytes; o += 16384)
              bulk_g2s(b_base + sb * b_slab_bytes + o, p.w_img + static_cast<size_t>(kb) * b_slab_bytes + o,
                       min(16384u, b_slab_bytes - o), &b_full[sb]);
            if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else {
    // =========================== A producers ===========================
    const int ptid = tid - PROD_WARP0 * 32;
    int sa = 0; uint32_t pa = 0;
    if constexpr (kConv1) {
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        Conv1Producer::stage_window(p, tile, ptid, s_win);
        asm volatile("bar.sync 1, %0;" ::"n"(NUM_PROD_THREADS) : "memory");     // window complete
        mbar_wait(&a_empty[sa], pa ^ 1);
        Conv1Producer::fill_slabs(a_base + sa * p.a_stage_bytes, ptid, s_win);
        fence_proxy_async_smem();
        mbar_arrive(&a_full[sa]);
        if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
        asm volatile("bar.sync 1, %0;" ::"n"(NUM_PROD_THREADS) : "memory");     // window free again
      }
    } else {
      DwProducer<MODE == 2 ? 2 : 1> prod;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        prod.begin_tile(p, tile, ptid);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&a_empty[sa], pa ^ 1);
          prod.produce(p, a_base + sa * p.a_stage_bytes, kb, s_dw);
          fence_proxy_async_smem();
          mbar_arrive(&a_full[sa]);
          if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
        }
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
// host side
// ------------------------------------------------------------------------------------------------

// Keras kernel [K, N] (row-major, K = input channel / tap) -> pre-swizzled fp16 slabs
// [num_kb][N rows x 128 B]; element (k, n) lands in slab k/64, row n, chunk (k%64)/8.
void build_weight_image(const float* w, int K, int N, std::vector<__half>& img) {
  const int num_kb = (K + SLAB_K - 1) / SLAB_K;
  img.assign(static_cast<size_t>(num_kb) * N * SLAB_K, __float2half_rn(0.0f));
  for (int k = 0; k < K; ++k) {
    const int kb = k / SLAB_K, kk = k % SLAB_K;
    for (int n = 0; n < N; ++n) {
      const size_t byte = static_cast<size_t>(kb) * N * ROW_BYTES + swz_off(n, kk / 8) + (kk % 8) * 2;
      img[byte / 2] = __float2half_rn(w[static_cast<size_t>(k) * N + n]);
    }
  }
}

template <int MODE>
int launch_tc_gemm(kws_handle* h, GemmParams& p, cudaStream_t st) {
  const bool conv1 = MODE == 0;
  // split cout into <= 256-wide instructions
  p.n_halves = p.cout > 256 ? 2 : 1;
  p.n_inst = p.cout / p.n_halves;
  if (p.n_inst % 16 || p.n_inst > 256 || p.cout > TMEM_COLS)
    return fail(h, KWS_EUNSUPPORTED, "unsupported channel count for the tensor-core path");
  p.acc_stages = std::min(2, TMEM_COLS / p.cout);
  p.a_stage_bytes = conv1 ? 2 * A_SLAB_BYTES : A_SLAB_BYTES;
  p.a_stages = conv1 ? 3 : 4;
  const int b_slab = p.cout * ROW_BYTES;
  // weights resident in smem when they fit next to the A ring, else a streaming ring
  p.b_stages = p.num_kb; p.b_resident = 1;
  SmemLayout lay = smem_layout(p, conv1);
  if (static_cast<int>(lay.total) > SMEM_LIMIT) {
    p.b_resident = 0;
    int budget = SMEM_LIMIT - static_cast<int>(lay.total) + p.b_stages * b_slab;
    p.b_stages = std::max(1, std::min(MAX_STAGES, budget / b_slab));
    p.b_stages = std::min(p.b_stages, 4);
    lay = smem_layout(p, conv1);
    if (static_cast<int>(lay.total) > SMEM_LIMIT || p.b_stages < 2)
      return fail(h, KWS_EUNSUPPORTED, "weight slab does not fit in shared memory");
  }
  if (p.num_kb > MAX_STAGES && p.b_resident) return fail(h, KWS_EUNSUPPORTED, "too many K slabs");
  static bool attr_set[3] = {false, false, false};
  if (!attr_set[MODE]) {
    KWS_CUDA(h, cudaFuncSetAttribute(tc_gemm_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    attr_set[MODE] = true;
  }
  const int grid = std::min(p.num_tiles, h->num_sms);
  if (grid <= 0) return KWS_OK;
  KWS_T0(h, MODE == 0 ? KC_CONV1 : KC_BLOCKS, st);
  tc_gemm_kernel<MODE><<<grid, TC_THREADS, lay.total, st>>>(p);
  KWS_T1(h, st);
  KWS_LAUNCH_CHECK(h);
  return KWS_OK;
}

}  // namespace

int model_build_tc(kws_handle* h, Model& m, const std::vector<std::vector<float>>& pw_host,
                   const std::vector<float>& conv1_host) {
  // conv1: fold the 3 overlapping patches into 80 taps: W80[u, co] = sum_f W[f, u - 20 f, co]
  std::vector<float> w80(static_cast<size_t>(CONV1_K) * m.c0, 0.0f);
  for (int f = 0; f < 3; ++f)
    for (int i = 0; i < 40; ++i)
      for (int co = 0; co < m.c0; ++co)
        w80[static_cast<size_t>(20 * f + i) * m.c0 + co] += conv1_host[static_cast<size_t>(f * 40 + i) * m.c0 + co];
  std::vector<__half> all, img;
  std::vector<size_t> offs;
  build_weight_image(w80.data(), CONV1_K, m.c0, img);
  offs.push_back(all.size()); all.insert(all.end(), img.begin(), img.end());
  for (int i = 0; i < NUM_BLOCKS; ++i) {
    build_weight_image(pw_host[i].data(), m.layers[i].cin, m.layers[i].cout, img);
    while (all.size() % 512) all.push_back(__float2half_rn(0.0f));      // 1024-byte aligned slabs
    offs.push_back(all.size()); all.insert(all.end(), img.begin(), img.end());
  }
  KWS_CUDA(h, cudaMalloc(&m.tc_blob, all.size() * sizeof(__half)));
  KWS_CUDA(h, cudaMemcpy(m.tc_blob, all.data(), all.size() * sizeof(__half), cudaMemcpyHostToDevice));
  __half* base = static_cast<__half*>(m.tc_blob);
  m.tc_conv1 = base + offs[0];
  for (int i = 0; i < NUM_BLOCKS; ++i) m.tc_pw[i] = base + offs[i + 1];
  return KWS_OK;
}

int launch_forward_tc(kws_handle* h, Model& m, const float* wav, int B, const ViewTable& vt,
                      float* probs_mean, int32_t* argmax, cudaStream_t st, int dbg_layer, float* dbg_out) {
  const int V = vt.n;
  const int clips_per_chunk = std::max(1, h->max_rows / V);
  const size_t need = static_cast<size_t>(clips_per_chunk) * V * m.max_act_elems * sizeof(__half);
  if (h->act_bytes < need) {
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
    {
      GemmParams p{};
      p.wav = wav + static_cast<size_t>(b0) * L; p.vt = vt; p.n_views = V;
      p.w_img = reinterpret_cast<const uint8_t*>(m.tc_conv1);
      p.scale = m.bn_scale[0]; p.shift = m.bn_shift[0]; p.out = cur;
      p.cin = CONV1_K; p.cout = m.c0; p.t_out = m.t0; p.rows_out = rows * m.t0;
      p.tiles_per_group = (m.t0 + TILE_M - 1) / TILE_M;
      p.num_tiles = rows * p.tiles_per_group;
      p.num_kb = 2; p.last_ksteps = 1;
      int rc = launch_tc_gemm<0>(h, p, st);
      if (rc) return rc;
    }
    if (dbg_layer == 0) return launch_to_float(h, cur, true, dbg_out, static_cast<size_t>(rows) * m.t0 * m.c0, st);
    for (int i = 0; i < NUM_BLOCKS; ++i) {
      const LayerDesc& d = m.layers[i];
      GemmParams p{};
      p.act_in = cur; p.dw = m.w_dw[i];
      p.w_img = reinterpret_cast<const uint8_t*>(m.tc_pw[i]);
      p.scale = m.bn_scale[i + 1]; p.shift = m.bn_shift[i + 1]; p.out = nxt;
      p.cin = d.cin; p.cout = d.cout; p.stride = d.stride; p.pad_left = d.pad_left;
      p.t_in = d.t_in; p.t_out = d.t_out; p.rows_out = rows * d.t_out;
      p.num_tiles = (p.rows_out + TILE_M - 1) / TILE_M;
      p.num_kb = d.cin / SLAB_K; p.last_ksteps = 4;
      if (d.cin % SLAB_K) return fail(h, KWS_EUNSUPPORTED, "channel count must be a multiple of 64");
      int rc = d.stride == 1 ? launch_tc_gemm<1>(h, p, st) : launch_tc_gemm<2>(h, p, st);
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
