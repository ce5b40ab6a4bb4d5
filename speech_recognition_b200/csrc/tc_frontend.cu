// Tensor-core tier of the STFT front end: STFT -> |.| -> mel -> log (-> DCT) in ONE kernel
// (reference input_data.py:361-381; constants from the logs_195 GraphDef).
//
// The STFT is a DFT-as-GEMM on tcgen05:  rows = (clip, frame), K = window samples (480 -> 8 slabs
// of 64, the tail zero), N = 512 columns = (re, im) of bins 0..255 with the periodic Hann window
// folded into the basis.  fp32 accuracy comes from a split-fp16 ("3-pass") product: x = x_hi + x_lo
// and basis = b_hi + b_lo as fp16 pairs, D += x_hi b_hi + x_lo b_hi + x_hi b_lo (the dropped
// x_lo b_lo term is 2^-22 relative), accumulated in fp32 in TMEM.  The accumulator of a 128-frame
// tile (128 lanes x 512 columns) fills TMEM exactly; the epilogue reads (re, im) pairs, takes the
// magnitude and applies the mel matrix as what it is -- a band matrix with at most two non-zeros
// per bin -- with two running accumulators per frame, then log(. + 1e-6) and, for MFCC, the
// DCT-II against a shared-memory basis.  Neither the spectrogram nor the mel energies touch HBM.
//
// Roles (576 threads, 1 CTA / SM, static round-robin over 128-frame tiles):
//   warps 0-7  epilogue : tcgen05.ld -> magnitude -> banded mel (linear sums to smem) | log -> (DCT) -> global.
//                         The accumulator fills TMEM, so MMA and epilogue of a tile alternate and the epilogue
//                         (256 bins x sqrt + 2 FMA per frame, latency-bound with one warp per scheduler) was 3/4
//                         of the tile time: two warps per TMEM lane quarter now take 128 bins each; the two
//                         mel filters that straddle bin 128 are summed from both halves in a fixed order.
//   warp  8    MMA      : one thread issues tcgen05.mma
//   warp  9    B loader : cp.async.bulk of pre-swizzled basis blocks (hi / lo, 256 columns x 64 k)
//   warps 10-17 A producers: implicit framing from the waveform (each sample is read from HBM once,
//                           the 3x frame overlap is served by L1/L2), hi/lo split, swizzled store
// Bin 256 (Nyquist) has no mel weight for any upper edge below the Nyquist frequency; the
// 'spectrogram' representation (all 257 bins) and exotic window sizes stay on the fp32 GEMM chain.
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "tc_common.cuh"

namespace kws {

using namespace tc;

namespace {

constexpr int F_EPI_WARPS = 8;
constexpr int F_MMA_WARP = 8;
constexpr int F_LOAD_WARP = 9;
constexpr int F_PROD_WARP0 = 10;
constexpr int F_PROD_THREADS = 256;
constexpr int F_THREADS = 32 * F_PROD_WARP0 + F_PROD_THREADS;     // 576
constexpr int F_A_STAGES = 2;                                      // stage = hi slab + lo slab (32 KB)
constexpr int F_B_STAGES = 4;                                      // max ring depth; block = 256 columns x 64 k (32 KB)
constexpr int F_NH = 256;                                          // columns per MMA instruction
constexpr int F_BINS = 256;                                        // bins on the tensor-core path
constexpr int F_B_BLOCK = F_NH * ROW_BYTES;                        // 32 KB
constexpr int F_A_STAGE = 2 * A_SLAB_BYTES;                        // 32 KB
constexpr int F_DCT_LD = 64;                                       // padded n_keep
constexpr int F_SMEM_LIMIT = 227 * 1024;

struct DftParams {
  const float* wav;          // [B, 16000]
  float* out;                // [rows_total, out_dim]
  const uint8_t* b_img;      // basis blocks, index ((kb * 2 + nh) * 2 + part), part 0 = hi, 1 = lo
  const float2* bin_tab;     // [256] {w_a, w_b}: weights of the bin for the open filter pair (m, m + 1), then
                             // [16] uint32: 2 bits per bin = how many filters finish BEFORE the bin (0..3)
  const float* dct;          // [n_mel][64] zero padded
  int frames, hop, win, n_mel, n_keep;
  int rows_total, num_tiles, num_kb, last_ksteps;
  int b_stages;              // basis ring depth (2..4, what shared memory allows)
  int m_split;               // mel filter that is open when bin 128 starts: filters m_split, m_split + 1 straddle the halves
  int floor_mode;            // 0: log(mel + 1e-6) (input_data.py:378); 1: log(max(mel, 1e-12)) (contrib_audio Mfcc)
};

struct FSmem { uint32_t a_off, b_off, tab_off, dct_off, part_off, bar_off, total; };

__host__ __device__ inline FSmem f_smem(int n_mel, bool mfcc, int b_stages) {
  FSmem s; uint32_t o = 0;
  s.a_off = o; o += F_A_STAGES * F_A_STAGE;
  s.b_off = o; o += static_cast<uint32_t>(b_stages) * F_B_BLOCK;
  s.tab_off = o; o += F_BINS * 8 + (F_BINS / 16) * 4;               // weight pairs + advance words
  s.dct_off = o; o += mfcc ? static_cast<uint32_t>(n_mel) * F_DCT_LD * 4u : 0u;
  s.part_off = o; o += static_cast<uint32_t>(n_mel + 2) * TILE_M * 4u;   // linear mel sums [n_mel][128 rows] + 2 overlap rows
  s.bar_off = o; o += (2 * F_A_STAGES + 2 * F_B_STAGES + 2) * 8 + 16;
  s.total = o + 1024;
  return s;
}

template <bool MFCC, bool FLOOR>
__global__ void __launch_bounds__(F_THREADS, 1) stft_mel_tc_kernel(const DftParams p) {
  extern __shared__ uint8_t smem_raw[];
  const FSmem lay = f_smem(p.n_mel, MFCC, p.b_stages);
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_base = smem + lay.a_off;
  uint8_t* b_base = smem + lay.b_off;
  float2* s_w = reinterpret_cast<float2*>(smem + lay.tab_off);
  uint32_t* s_adv = reinterpret_cast<uint32_t*>(s_w + F_BINS);
  float* s_dct = reinterpret_cast<float*>(smem + lay.dct_off);
  float* s_part = reinterpret_cast<float*>(smem + lay.part_off);  // [n_mel + 2][128] linear mel sums, see the epilogue
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + lay.bar_off);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + F_A_STAGES;
  uint64_t* b_full = a_empty + F_A_STAGES;
  uint64_t* b_empty = b_full + F_B_STAGES;
  uint64_t* acc_full = b_empty + F_B_STAGES;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < F_BINS; i += F_THREADS) s_w[i] = p.bin_tab[i];
  if (tid < F_BINS / 16) s_adv[tid] = reinterpret_cast<const uint32_t*>(p.bin_tab + F_BINS)[tid];
  if (MFCC)
    for (int i = tid; i < p.n_mel * F_DCT_LD; i += F_THREADS) s_dct[i] = p.dct[i];
  if (warp == F_MMA_WARP) {
    if (lane == 0) {
      for (int i = 0; i < F_A_STAGES; ++i) { mbar_init(&a_full[i], F_PROD_THREADS); mbar_init(&a_empty[i], 1); }
      for (int i = 0; i < F_B_STAGES; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
      mbar_init(acc_full, 1);
      mbar_init(acc_empty, F_EPI_WARPS * 32);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < F_EPI_WARPS) {
    // =========================== epilogue ===========================
    uint32_t acc_phase = 0;
    const int out_dim = MFCC ? p.n_keep : p.n_mel;
    const int q = warp & 3, hh = warp >> 2;                      // TMEM lane quarter; bin half [128 hh, 128 hh + 128)
    const int row = q * 32 + lane;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(acc_full, acc_phase);
      tc_fence_after();
      const long long R = static_cast<long long>(tile) * TILE_M + row;
      const bool ok = R < p.rows_total;
      float* orow = p.out + R * out_dim;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(hh * 256);
      // Banded mel with two running sums per thread: acc_a = filter m, acc_b = filter m + 1; when the table says
      // "advance" the finished filter goes to shared memory.  Half hh writes its linear sums to rows
      // [m + 2 hh] of s_part: half 0 owns filters 0 .. m_split + 1 (rows 0 .. m_split + 1), half 1 owns filters
      // m_split .. n_mel - 1 (rows m_split + 2 .. n_mel + 1), so the store address is one pointer that moves by a row
      // per finished filter and the two filters that straddle bin 128 are summed by the reader in a fixed order.
      // (r02: the first version kept a filter index, a bounds check and the straddle case inside the per-bin code;
      // unrolled 128 times that was 72 KB of instructions, and the warps sat in instruction-cache misses -- stall_no_inst
      // was a third of the samples of this loop, which ran at ~250 cycles per bin.  Now 8 bins are one unit: their
      // weight pairs are loaded and their magnitudes computed first (8 independent FMUL / FFMA / MUFU.SQRT, so the MUFU
      // and LDS latencies are paid once per unit, not per bin), then per bin two FFMAs and a predicated advance whose
      // predicate is a bit test on a register (the advance counts of 16 bins are one word); the column loop is not
      // unrolled.)
      float acc_a = 0.0f, acc_b = 0.0f;
      uint32_t di = static_cast<uint32_t>((hh ? p.m_split + 2 : 0) * TILE_M + row);   // s_part index of the open filter's sum
      const float2* wtab = s_w + hh * (F_BINS / 2);
      const uint32_t* advw = s_adv + hh * (F_BINS / 32);
      auto bins8 = [&](const uint32_t (&v)[16], const float2* wp, uint32_t bits) {   // bits: 2 per bin, bin 0 in bits 0-1
        float2 w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = wp[k];
        float mag[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {                                        // ComplexAbs, input_data.py:366
          const float re = __uint_as_float(v[2 * k]), im = __uint_as_float(v[2 * k + 1]);
          asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(mag[k]) : "f"(fmaf(re, re, im * im)));   // MUFU.SQRT: 1 ulp-class, far inside the 1e-4 tier
        }
        if ((bits & 0xAAAAu) == 0u) {
          // no bin of the unit finishes more than one filter (always, unless the filters are narrower than a bin):
          // branch-free -- a GPU does not predict branches, and a warp-uniform branch per bin cost more than the math
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const bool adv = ((bits >> (2 * k)) & 1u) != 0u;
            if (adv) s_part[di] = acc_a;                                     // predicated store
            di += adv ? TILE_M : 0;
            acc_a = adv ? acc_b : acc_a;
            acc_b = adv ? 0.0f : acc_b;
            acc_a = fmaf(mag[k], w[k].x, acc_a);
            acc_b = fmaf(mag[k], w[k].y, acc_b);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (bits & (3u << (2 * k))) {                                    // warp-uniform
              int adv = static_cast<int>((bits >> (2 * k)) & 3u);
#pragma unroll 1
              do { s_part[di] = acc_a; di += TILE_M; acc_a = acc_b; acc_b = 0.0f; } while (--adv);
            }
            acc_a = fmaf(mag[k], w[k].x, acc_a);
            acc_b = fmaf(mag[k], w[k].y, acc_b);
          }
        }
      };
      {                                                          // next TMEM load (16 columns = 8 bins) in flight during the math
        uint32_t va[16], vb[16];
        tmem_ld16(taddr, va);
#pragma unroll 1
        for (int c0 = 0; c0 < 256; c0 += 32) {
          const uint32_t bits = advw[c0 >> 5];
          tmem_ld_wait();
          tmem_ld16(taddr + c0 + 16, vb);
          bins8(va, wtab + c0 / 2, bits);
          tmem_ld_wait();
          if (c0 + 32 < 256) tmem_ld16(taddr + c0 + 32, va);
          bins8(vb, wtab + c0 / 2 + 8, bits >> 16);
        }
      }
      tc_fence_before();
      mbar_arrive(acc_empty);                                    // TMEM is free for the next tile
      acc_phase ^= 1;
      {                                                          // flush: half 0 its two open filters, half 1 every filter that is left
        const int m_open = static_cast<int>(di / TILE_M) - 2 * hh;
        const int m_end = hh ? p.n_mel : min(p.n_mel, m_open + 2);
        for (int m = m_open; m < m_end; ++m) { s_part[di] = acc_a; di += TILE_M; acc_a = acc_b; acc_b = 0.0f; }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(F_EPI_WARPS * 32) : "memory");    // all linear mel sums of the tile are in smem
      auto log_mel = [&](int m) {
        float v = m <= p.m_split + 1 ? s_part[m * TILE_M + row] : 0.0f;           // first bin half's share
        if (m >= p.m_split) v += s_part[(m + 2) * TILE_M + row];                  // second half's share
        return FLOOR ? logf(fmaxf(v, 1e-12f)) : logf(v + 1e-6f);             // TF mfcc.cc / input_data.py:378
      };
      if (MFCC) {                                                // this thread: DCT outputs [32 hh, 32 hh + 32) of its row
        float dctacc[F_DCT_LD / 2];
#pragma unroll
        for (int k = 0; k < F_DCT_LD / 2; ++k) dctacc[k] = 0.0f;
        for (int m = 0; m < p.n_mel; ++m) {
          const float lm = log_mel(m);
          const float4* d4 = reinterpret_cast<const float4*>(s_dct + m * F_DCT_LD + hh * (F_DCT_LD / 2));
#pragma unroll
          for (int k = 0; k < F_DCT_LD / 8; ++k) {
            const float4 d = d4[k];
            dctacc[4 * k] = fmaf(lm, d.x, dctacc[4 * k]);
            dctacc[4 * k + 1] = fmaf(lm, d.y, dctacc[4 * k + 1]);
            dctacc[4 * k + 2] = fmaf(lm, d.z, dctacc[4 * k + 2]);
            dctacc[4 * k + 3] = fmaf(lm, d.w, dctacc[4 * k + 3]);
          }
        }
        if (ok) {
#pragma unroll
          for (int k = 0; k < F_DCT_LD / 2; ++k)
            if (hh * (F_DCT_LD / 2) + k < p.n_keep) orow[hh * (F_DCT_LD / 2) + k] = dctacc[k];
        }
      } else if (ok) {                                           // this thread: log-mel outputs of its half of the filters
        const int m_half = (p.n_mel + 1) / 2;
        for (int m = hh * m_half; m < min(p.n_mel, (hh + 1) * m_half); ++m) orow[m] = log_mel(m);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(F_EPI_WARPS * 32) : "memory");    // smem sums may be overwritten
    }
  } else if (warp == F_MMA_WARP) {
    // =========================== MMA issuer ===========================
    // all 32 lanes walk the loop (uniform control flow and registers); one elected lane issues.  Each
    // (A slab, basis block) pass is one predicated PTX sequence (tc_common.cuh umma_slab_commit).
    {
      const uint32_t idesc = umma_idesc_f16(TILE_M, F_NH, /*fp16*/ 0);
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(a_base)), b_lo0 = umma_desc_lo(smem_u32(b_base));
      const uint32_t a_full0 = smem_u32(a_full), a_empty0 = smem_u32(a_empty), b_full0 = smem_u32(b_full), b_empty0 = smem_u32(b_empty);
      const int num_kb = p.num_kb, last_ksteps = p.last_ksteps, b_stages = p.b_stages, num_tiles = p.num_tiles;
      int sa = 0; uint32_t pa = 0; int sb = 0; uint32_t pb = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(acc_empty, acc_phase ^ 1);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_addr(a_full0 + 8u * sa, pa);
          const uint32_t a_hi = a_lo0 + static_cast<uint32_t>(sa) * (F_A_STAGE >> 4);
          const uint32_t a_lo = a_hi + (A_SLAB_BYTES >> 4);
          const uint32_t ksteps = (kb == num_kb - 1) ? static_cast<uint32_t>(last_ksteps) : 4u;
          for (int nh = 0; nh < 2; ++nh) {
            const uint32_t d = tmem_base + nh * F_NH;
            mbar_wait_addr(b_full0 + 8u * sb, pb);
            tc_fence_after();
            uint32_t b = b_lo0 + static_cast<uint32_t>(sb) * (F_B_BLOCK >> 4);
            umma_slab_commit(d, a_hi, b, idesc, kb != 0 ? 1u : 0u, ksteps, 0u, 0u);          // x_hi * b_hi
            umma_slab_commit(d, a_lo, b, idesc, 1u, ksteps, b_empty0 + 8u * sb, 0u);            // x_lo * b_hi
            if (++sb == b_stages) { sb = 0; pb ^= 1; }
            mbar_wait_addr(b_full0 + 8u * sb, pb);
            tc_fence_after();
            b = b_lo0 + static_cast<uint32_t>(sb) * (F_B_BLOCK >> 4);
            umma_slab_commit(d, a_hi, b, idesc, 1u, ksteps, b_empty0 + 8u * sb,                  // x_hi * b_lo
                             nh == 1 ? a_empty0 + 8u * sa : 0u);
            if (++sb == b_stages) { sb = 0; pb ^= 1; }
          }
          if (++sa == F_A_STAGES) { sa = 0; pa ^= 1; }
        }
        umma_commit_elect(smem_u32(acc_full));
        acc_phase ^= 1;
      }
    }
  } else if (warp == F_LOAD_WARP) {
    // =========================== basis loader ===========================
    if (lane == 0) {
      int sb = 0; uint32_t pb = 0;
      const int blocks = p.num_kb * 4;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int j = 0; j < blocks; ++j) {
          mbar_wait(&b_empty[sb], pb ^ 1);
          mbar_arrive_expect_tx(&b_full[sb], F_B_BLOCK);
          const uint8_t* src = p.b_img + static_cast<size_t>(j) * F_B_BLOCK;
          uint8_t* dst = b_base + sb * F_B_BLOCK;
          bulk_g2s(dst, src, 16384, &b_full[sb]);
          bulk_g2s(dst + 16384, src + 16384, 16384, &b_full[sb]);
          if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else {
    // =========================== A producers ===========================
    const int ptid = tid - F_PROD_WARP0 * 32;
    const int c = ptid & 7, r0 = ptid >> 3;                      // rows r0 + 32 i
    int sa = 0; uint32_t pa = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const float* src[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long R = static_cast<long long>(tile) * TILE_M + r0 + 32 * i;
        if (R < p.rows_total) {
          const long long b = R / p.frames;
          const int f = static_cast<int>(R - b * p.frames);
          src[i] = p.wav + b * L + static_cast<long long>(p.hop) * f;     // frames[f, k] = x[hop f + k]
        } else {
          src[i] = nullptr;
        }
      }
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const int k0 = kb * SLAB_K + c * 8;
        float4 x[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          x[i][0] = x[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (src[i] != nullptr && k0 < p.win) {                 // win % 8 == 0: a chunk is all in or all out
            x[i][0] = __ldg(reinterpret_cast<const float4*>(src[i] + k0));
            x[i][1] = __ldg(reinterpret_cast<const float4*>(src[i] + k0 + 4));
          }
        }
        mbar_wait(&a_empty[sa], pa ^ 1);
        uint8_t* hi_slab = a_base + sa * F_A_STAGE;
        uint8_t* lo_slab = hi_slab + A_SLAB_BYTES;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float xs[8] = {x[i][0].x, x[i][0].y, x[i][0].z, x[i][0].w, x[i][1].x, x[i][1].y, x[i][1].z, x[i][1].w};
          uint4 hv, lv;
          __half2* hh = reinterpret_cast<__half2*>(&hv);
          __half2* ll = reinterpret_cast<__half2*>(&lv);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const __half2 h2 = __floats2half2_rn(xs[2 * e], xs[2 * e + 1]);
            const float2 back = __half22float2(h2);
            hh[e] = h2;
            ll[e] = __floats2half2_rn(xs[2 * e] - back.x, xs[2 * e + 1] - back.y);
          }
          const uint32_t off = swz_off(r0 + 32 * i, c);
          *reinterpret_cast<uint4*>(hi_slab + off) = hv;
          *reinterpret_cast<uint4*>(lo_slab + off) = lv;
        }
        fence_proxy_async_smem();
        mbar_arrive(&a_full[sa]);
        if (++sa == F_A_STAGES) { sa = 0; pa ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == F_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// Basis images and the banded-mel table.  basis = host fp32 [win, 2 * n_bins] (cos w, -sin w),
// mel = host fp32 [n_bins, n_mel], dct = host fp32 [n_mel, n_keep] of frontend_build.
int frontend_build_tc(kws_handle* h, const float* basis, const float* mel, const float* dct) {
  Frontend& fe = h->fe;
  fe.tc_ok = false;
  // shapes served by the tensor-core kernel; everything else stays on the fp32 chain
  if (fe.win % 8 || fe.hop % 4 || fe.n_fft != 512 || fe.n_keep > F_DCT_LD || fe.n_mel > 128) return KWS_OK;
  const int nb = fe.n_bins, num_kb = (fe.win + SLAB_K - 1) / SLAB_K;
  // ---- banded mel: every bin feeds at most two adjacent filters (m, m+1), m non-decreasing ----
  std::vector<float> tab(static_cast<size_t>(F_BINS) * 2 + F_BINS / 16, 0.0f);   // weight pairs, then the advance words
  std::vector<uint32_t> advw(F_BINS / 16, 0u);
  int m_cur = 0;
  fe.tc_m_split = 0;
  for (int j = 0; j < nb; ++j) {
    if (j == F_BINS / 2) fe.tc_m_split = m_cur;                   // filter open when the second bin half starts
    int lo = -1, hi = -1;
    for (int m = 0; m < fe.n_mel; ++m)
      if (mel[static_cast<size_t>(j) * fe.n_mel + m] != 0.0f) { if (lo < 0) lo = m; hi = m; }
    if (j >= F_BINS) { if (lo >= 0) return KWS_OK; continue; }     // weight on the Nyquist bin: fp32 chain
    int adv = 0;
    if (lo >= 0) {
      if (hi - lo > 1) return KWS_OK;                               // not banded: fp32 chain
      adv = std::max(0, hi - 1 - m_cur);
      if (adv > 3) return KWS_OK;                                   // two bits per bin in the advance words: fp32 chain
      m_cur += adv;
      if (lo < m_cur || hi > m_cur + 1) return KWS_OK;
      tab[2 * j + 0] = mel[static_cast<size_t>(j) * fe.n_mel + m_cur];
      tab[2 * j + 1] = m_cur + 1 < fe.n_mel ? mel[static_cast<size_t>(j) * fe.n_mel + m_cur + 1] : 0.0f;
    }
    advw[j / 16] |= static_cast<uint32_t>(adv) << (2 * (j % 16));
  }
  std::memcpy(&tab[static_cast<size_t>(F_BINS) * 2], advw.data(), advw.size() * sizeof(uint32_t));
  // ---- split-fp16 basis blocks ----
  const size_t n_blocks = static_cast<size_t>(num_kb) * 4;
  std::vector<__half> img(n_blocks * F_B_BLOCK / 2, __float2half_rn(0.0f));
  for (int kb = 0; kb < num_kb; ++kb)
    for (int nh = 0; nh < 2; ++nh)
      for (int n = 0; n < F_NH; ++n) {
        const int col = nh * F_NH + n;                              // (re, im) of bin col / 2
        for (int kk = 0; kk < SLAB_K; ++kk) {
          const int k = kb * SLAB_K + kk;
          const float v = k < fe.win ? basis[static_cast<size_t>(k) * 2 * nb + col] : 0.0f;
          const __half vh = __float2half_rn(v);
          const __half vl = __float2half_rn(v - __half2float(vh));
          const size_t blk = (static_cast<size_t>(kb) * 2 + nh) * 2;
          const size_t byte = swz_off(n, kk / 8) + (kk % 8) * 2;
          img[(blk * F_B_BLOCK + byte) / 2] = vh;
          img[((blk + 1) * F_B_BLOCK + byte) / 2] = vl;
        }
      }
  std::vector<float> dct_pad(static_cast<size_t>(fe.n_mel) * F_DCT_LD, 0.0f);
  for (int n = 0; n < fe.n_mel; ++n)
    for (int k = 0; k < fe.n_keep; ++k) dct_pad[static_cast<size_t>(n) * F_DCT_LD + k] = dct[static_cast<size_t>(n) * fe.n_keep + k];
  const size_t img_bytes = img.size() * sizeof(__half);
  const size_t tab_bytes = tab.size() * sizeof(float);
  const size_t dct_bytes = dct_pad.size() * sizeof(float);
  KWS_CUDA(h, cudaMalloc(&fe.tc_blob, img_bytes + tab_bytes + dct_bytes));
  uint8_t* base = static_cast<uint8_t*>(fe.tc_blob);
  KWS_CUDA(h, cudaMemcpy(base, img.data(), img_bytes, cudaMemcpyHostToDevice));
  KWS_CUDA(h, cudaMemcpy(base + img_bytes, tab.data(), tab_bytes, cudaMemcpyHostToDevice));
  KWS_CUDA(h, cudaMemcpy(base + img_bytes + tab_bytes, dct_pad.data(), dct_bytes, cudaMemcpyHostToDevice));
  fe.tc_basis = base;
  fe.tc_bin_tab = reinterpret_cast<float*>(base + img_bytes);
  fe.tc_dct = reinterpret_cast<float*>(base + img_bytes + tab_bytes);
  fe.tc_kblocks = num_kb;
  fe.tc_last_ksteps = (fe.win - (num_kb - 1) * SLAB_K + 15) / 16;
  fe.tc_ok = true;
  return KWS_OK;
}

int launch_features_tc(kws_handle* h, const float* wav, int B, int kind, float* out, cudaStream_t st) {
  Frontend& fe = h->fe;
  if (!fe.tc_ok || kind == KWS_FEAT_SPEC || reinterpret_cast<uintptr_t>(wav) % 16)
    return launch_features_f32(h, wav, B, kind, out, st);
  DftParams p{};
  p.wav = wav; p.out = out;
  p.b_img = fe.tc_basis;
  p.bin_tab = reinterpret_cast<const float2*>(fe.tc_bin_tab);
  p.dct = fe.tc_dct;
  p.frames = fe.frames; p.hop = fe.hop; p.win = fe.win; p.n_mel = fe.n_mel; p.n_keep = fe.n_keep;
  p.rows_total = B * fe.frames;
  p.num_tiles = (p.rows_total + TILE_M - 1) / TILE_M;
  p.num_kb = fe.tc_kblocks; p.last_ksteps = fe.tc_last_ksteps;
  p.floor_mode = fe.flavour == 1;
  const bool mfcc = kind == KWS_FEAT_MFCC;
  p.m_split = fe.tc_m_split;
  p.b_stages = F_B_STAGES;
  while (p.b_stages > 2 && static_cast<int>(f_smem(fe.n_mel, mfcc, p.b_stages).total) > F_SMEM_LIMIT) --p.b_stages;
  const FSmem lay = f_smem(fe.n_mel, mfcc, p.b_stages);
  if (static_cast<int>(lay.total) > F_SMEM_LIMIT) return launch_features_f32(h, wav, B, kind, out, st);
  if (!(h->smem_attr_done & 1u)) {
    KWS_CUDA(h, cudaFuncSetAttribute(stft_mel_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM_LIMIT));
    KWS_CUDA(h, cudaFuncSetAttribute(stft_mel_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM_LIMIT));
    KWS_CUDA(h, cudaFuncSetAttribute(stft_mel_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM_LIMIT));
    KWS_CUDA(h, cudaFuncSetAttribute(stft_mel_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM_LIMIT));
    h->smem_attr_done |= 1u;
  }
  const int grid = std::min(p.num_tiles, h->num_sms);
  KWS_T0(h, KC_DFT, st);
  if (p.floor_mode) {
    if (mfcc) stft_mel_tc_kernel<true, true><<<grid, F_THREADS, lay.total, st>>>(p);
    else stft_mel_tc_kernel<false, true><<<grid, F_THREADS, lay.total, st>>>(p);
  } else {
    if (mfcc) stft_mel_tc_kernel<true, false><<<grid, F_THREADS, lay.total, st>>>(p);
    else stft_mel_tc_kernel<false, false><<<grid, F_THREADS, lay.total, st>>>(p);
  }
  KWS_T1(h, st);
  KWS_LAUNCH_CHECK(h);
  return KWS_OK;
}

}  // namespace kws
