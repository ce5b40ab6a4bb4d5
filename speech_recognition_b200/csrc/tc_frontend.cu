// Tensor-core tier of the STFT front end: STFT -> |.| -> mel -> log (-> DCT) in ONE kernel
// (reference input_data.py:361-381; constants from the logs_195 GraphDef).
//
// The STFT is a DFT-as-GEMM on tcgen05 with one radix-2 step taken out of it.  With y[n] = w[n] x[n] (periodic Hann,
// applied in fp32 by the producers exactly like the reference's `frames * window`) and the 512-point transform
//     X[2j]     = sum_{n < 256} (y[n] + y[n + 256]) e^{-2 pi i (2j) n / 512}
//     X[2j + 1] = sum_{n < 256} (y[n] - y[n + 256]) e^{-2 pi i (2j + 1) n / 512}
// the even and the odd bins are two K = 256 contractions with N = 256 columns each ((re, im) of 128 bins) instead of
// one K = 480, N = 512 contraction: half the tensor work and half the basis bytes streamed from L2 (r02; until then the
// window was folded into a [480 x 512] basis).  rows = (clip, frame); per tile 8 stages = (K slab 0..3) x (sum, difference),
// stage s accumulates into TMEM columns [256 (s & 1), +256).  fp32 accuracy comes from a split-fp16 ("3-pass") product:
// a = a_hi + a_lo and basis = b_hi + b_lo as fp16 pairs, D += a_hi b_hi + a_lo b_hi + a_hi b_lo (the dropped a_lo b_lo
// term is 2^-22 relative), accumulated in fp32 in TMEM.  The two accumulators of a 128-frame tile fill TMEM exactly; the
// epilogue reads (re, im) pairs alternately from the even and the odd half -- bins in ascending order -- takes the magnitude
// and applies the mel matrix as what it is -- a band matrix with at most two non-zeros per bin -- with two running
// accumulators per frame, then log(. + 1e-6) and, for MFCC, the DCT-II against a shared-memory basis.  Neither the
// spectrogram nor the mel energies touch HBM.
//
// Roles (576 threads, 1 CTA / SM, static round-robin over 128-frame tiles):
//   warps 0-7  epilogue : tcgen05.ld -> magnitude -> banded mel (linear sums to smem) | log -> (DCT) -> global.
//                         The accumulators fill TMEM, so MMA and the first epilogue phase of a tile alternate; two warps
//                         per TMEM lane quarter take 128 bins each, and the two mel filters that straddle bin 128 are
//                         summed from both halves in a fixed order.
//   warp  8    MMA      : one thread issues tcgen05.mma
//   warp  9    B loader : cp.async.bulk of pre-swizzled basis blocks (hi / lo, 256 columns x 64 k)
//   warps 10-17 A producers: implicit framing from the waveform (each sample is read from HBM once, the 3x frame
//                           overlap is served by L1/L2), window, sum / difference of the two window halves, hi/lo
//                           split, swizzled store
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
constexpr int F_NFFT = 512;
constexpr int F_KB = F_NFFT / 2 / SLAB_K;                          // K slabs of the folded window (4)
constexpr int F_STAGES_PER_TILE = 2 * F_KB;                        // (slab, sum | difference)
constexpr int F_B_BLOCK = F_NH * ROW_BYTES;                        // 32 KB
constexpr int F_A_STAGE = 2 * A_SLAB_BYTES;                        // 32 KB
constexpr int F_DCT_LD = 64;                                       // padded n_keep
constexpr int F_PART_LD = TILE_M + 1;                              // row stride of the mel sums (conflict-free both ways)
constexpr int F_SMEM_LIMIT = 227 * 1024;

struct DftParams {
  const float* wav;          // [B, 16000]
  float* out;                // [rows_total, out_dim]
  const uint8_t* b_img;      // basis blocks, index ((kb * 2 + odd) * 2 + part), part 0 = hi, 1 = lo
  const float2* bin_tab;     // [256] {w_a, w_b}: weights of the bin for the open filter pair (m, m + 1), then
                             // [16] uint32: 2 bits per bin = how many filters finish BEFORE the bin (0..3)
  const float* dct;          // [n_mel][64] zero padded
  const float* hann;         // [512] window, zero from `win` on
  int frames, hop, win, n_mel, n_keep;
  int rows_total, num_tiles;
  int b_stages;              // basis ring depth (2..4, what shared memory allows)
  int m_split;               // mel filter that is open when bin 128 starts: filters m_split, m_split + 1 straddle the halves
  uint32_t mel_magic;        // ceil(2^32 / n_mel): idx / n_mel == __umulhi(idx, mel_magic) for idx < 128 * n_mel
  int floor_mode;            // 0: log(mel + 1e-6) (input_data.py:378); 1: log(max(mel, 1e-12)) (contrib_audio Mfcc)
};

struct FSmem { uint32_t a_off, b_off, tab_off, hann_off, dct_off, part_off, bar_off, total; };

__host__ __device__ inline FSmem f_smem(int n_mel, bool mfcc, int b_stages) {
  FSmem s; uint32_t o = 0;
  s.a_off = o; o += F_A_STAGES * F_A_STAGE;
  s.b_off = o; o += static_cast<uint32_t>(b_stages) * F_B_BLOCK;
  s.tab_off = o; o += F_BINS * 8 + (F_BINS / 16) * 4;               // weight pairs + advance words
  s.hann_off = o; o += F_NFFT * 4;
  s.dct_off = o; o += mfcc ? static_cast<uint32_t>(n_mel) * F_DCT_LD * 4u : 0u;
  s.part_off = o; o += static_cast<uint32_t>(n_mel + 2) * F_PART_LD * 4u;   // linear mel sums, see the epilogue
  o = (o + 15u) & ~15u;
  s.bar_off = o; o += (2 * F_A_STAGES + 2 * F_B_STAGES + 2) * 8 + 16;
  s.total = o + 1024;
  return s;
}

// a[8] fp32 -> fp16 hi / lo parts (a = hi + lo to 22 bits)
__device__ __forceinline__ void split8(const float (&a)[8], uint4& hv, uint4& lv) {
  __half2* hh = reinterpret_cast<__half2*>(&hv);
  __half2* ll = reinterpret_cast<__half2*>(&lv);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __half2 h2 = __floats2half2_rn(a[2 * e], a[2 * e + 1]);
    const float2 back = __half22float2(h2);
    hh[e] = h2;
    ll[e] = __floats2half2_rn(a[2 * e] - back.x, a[2 * e + 1] - back.y);
  }
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
  float* s_hann = reinterpret_cast<float*>(smem + lay.hann_off);
  float* s_dct = reinterpret_cast<float*>(smem + lay.dct_off);
  float* s_part = reinterpret_cast<float*>(smem + lay.part_off);  // [n_mel + 2][129] linear mel sums, see the epilogue
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
  for (int i = tid; i < F_NFFT; i += F_THREADS) s_hann[i] = p.hann[i];
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
    const int n_mel = p.n_mel, m_split = p.m_split, ms1 = p.m_split + 1;
    const uint32_t mel_magic = p.mel_magic;
    const float2* wtab = s_w + hh * (F_BINS / 2);
    const uint32_t* advw = s_adv + hh * (F_BINS / 32);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(acc_full, acc_phase);
      tc_fence_after();
      // even bins 2j: (re, im) in columns 2j, 2j + 1; odd bins 2j + 1: columns 256 + 2j, 256 + 2j + 1
      const uint32_t tp = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(hh * (F_BINS / 2));
      // Banded mel with two running sums per thread: acc_a = filter m, acc_b = filter m + 1; when the table says
      // "advance" the finished filter goes to shared memory.  Half hh writes its linear sums to rows
      // [m + 2 hh] of s_part: half 0 owns filters 0 .. m_split + 1 (rows 0 .. m_split + 1), half 1 owns filters
      // m_split .. n_mel - 1 (rows m_split + 2 .. n_mel + 1), so the store index is one counter that moves by a row
      // per finished filter and the two filters that straddle bin 128 are summed by the reader in a fixed order.
      // (r02: the first version kept a filter index, a bounds check and the straddle case inside the per-bin code;
      // unrolled 128 times that was 72 KB of instructions, and the warps sat in instruction-cache misses -- stall_no_inst
      // was a third of the samples of this loop, which ran at ~250 cycles per bin.  Now 16 bins are one unit: their
      // magnitudes first (independent FMUL / FFMA / MUFU.SQRT, so the MUFU latency is paid once per unit and the TMEM
      // loads of the next unit fly under the rest), then per bin two FFMAs and a predicated advance whose predicate is
      // a bit test on a register (the advance counts of 16 bins are one word); the unit loop is not unrolled.)
      float acc_a = 0.0f, acc_b = 0.0f;
      uint32_t di = static_cast<uint32_t>((hh ? m_split + 2 : 0) * F_PART_LD + row);   // s_part index of the open filter's sum
      auto accumulate8 = [&](const float (&mag)[16], const int off, const float2* wp, const uint32_t bits) {
        float2 w[8];                                              // bits: 2 per bin, bin 0 of the eight in bits 0-1
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = wp[k];
        if ((bits & 0xAAAAu) == 0u) {
          // no bin of the eight finishes more than one filter (always, unless the filters are narrower than a bin):
          // branch-free -- a GPU does not predict branches, and a warp-uniform branch per bin cost more than the math
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const bool adv = ((bits >> (2 * k)) & 1u) != 0u;
            if (adv) s_part[di] = acc_a;                                     // predicated store
            di += adv ? F_PART_LD : 0;
            acc_a = adv ? acc_b : acc_a;
            acc_b = adv ? 0.0f : acc_b;
            acc_a = fmaf(mag[off + k], w[k].x, acc_a);
            acc_b = fmaf(mag[off + k], w[k].y, acc_b);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (bits & (3u << (2 * k))) {                                    // warp-uniform
              int adv = static_cast<int>((bits >> (2 * k)) & 3u);
#pragma unroll 1
              do { s_part[di] = acc_a; di += F_PART_LD; acc_a = acc_b; acc_b = 0.0f; } while (--adv);
            }
            acc_a = fmaf(mag[off + k], w[k].x, acc_a);
            acc_b = fmaf(mag[off + k], w[k].y, acc_b);
          }
        }
      };
      {
        uint32_t ve[16], vo[16];                                 // 8 even and 8 odd bins = 16 consecutive bins
        tmem_ld16(tp, ve);
        tmem_ld16(tp + F_NH, vo);
#pragma unroll 1
        for (int u = 0; u < F_BINS / 32; ++u) {
          tmem_ld_wait();
          float mag[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {                                      // ComplexAbs, input_data.py:366
            const float er = __uint_as_float(ve[2 * j]), ei = __uint_as_float(ve[2 * j + 1]);
            const float orr = __uint_as_float(vo[2 * j]), oi = __uint_as_float(vo[2 * j + 1]);
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(mag[2 * j]) : "f"(fmaf(er, er, ei * ei)));   // MUFU.SQRT: 1 ulp-class,
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(mag[2 * j + 1]) : "f"(fmaf(orr, orr, oi * oi)));   // far inside the 1e-4 tier
          }
          if (u + 1 < F_BINS / 32) {                             // the next unit's loads fly under the accumulation
            tmem_ld16(tp + 16 * (u + 1), ve);
            tmem_ld16(tp + F_NH + 16 * (u + 1), vo);
          }
          const uint32_t bits = advw[u];
          accumulate8(mag, 0, wtab + 16 * u, bits & 0xffffu);
          accumulate8(mag, 8, wtab + 16 * u + 8, bits >> 16);
        }
      }
      tc_fence_before();
      mbar_arrive(acc_empty);                                    // TMEM is free for the next tile
      acc_phase ^= 1;
      {                                                          // flush: half 0 its two open filters, half 1 every filter that is left
        const int m_open = static_cast<int>(di / F_PART_LD) - 2 * hh;
        const int m_end = hh ? n_mel : min(n_mel, m_open + 2);
        for (int m = m_open; m < m_end; ++m) { s_part[di] = acc_a; di += F_PART_LD; acc_a = acc_b; acc_b = 0.0f; }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(F_EPI_WARPS * 32) : "memory");    // all linear mel sums of the tile are in smem
      // filter m: first bin half's share in row m (m <= m_split + 1), second half's share in row m + 2 (m >= m_split)
      auto log_mel = [&](int m) {
        float v = m <= ms1 ? s_part[m * F_PART_LD + row] : 0.0f;
        if (m >= m_split) v += s_part[(m + 2) * F_PART_LD + row];
        return FLOOR ? logf(fmaxf(v, 1e-12f)) : logf(v + 1e-6f);             // TF mfcc.cc / input_data.py:378
      };
      const long long R = static_cast<long long>(tile) * TILE_M + row;
      if (MFCC) {                                                // this thread: DCT outputs [32 hh, 32 hh + 32) of its row
        float dctacc[F_DCT_LD / 2];
#pragma unroll
        for (int k = 0; k < F_DCT_LD / 2; ++k) dctacc[k] = 0.0f;
        for (int m = 0; m < n_mel; ++m) {
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
        if (R < p.rows_total) {
          float* orow = p.out + R * out_dim;
#pragma unroll
          for (int k = 0; k < F_DCT_LD / 2; ++k)
            if (hh * (F_DCT_LD / 2) + k < p.n_keep) orow[hh * (F_DCT_LD / 2) + k] = dctacc[k];
        }
      } else {
        // this thread: the logs of its half of the filters of its row, written back in place (filter m ends up in row
        // m, or m + 2 past the straddling pair); then the tile's [rows x n_mel] block, contiguous in the output, leaves
        // with coalesced stores (one strided 4-byte store per filter and thread was 32 sectors per instruction)
        const int m_half = (n_mel + 1) / 2;
        const int m_hi = min(n_mel, (hh + 1) * m_half);
#pragma unroll 4
        for (int m = hh * m_half; m < m_hi; ++m) {
          const float lm = log_mel(m);
          s_part[(m <= ms1 ? m : m + 2) * F_PART_LD + row] = lm;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(F_EPI_WARPS * 32) : "memory");
        const long long rows_left = static_cast<long long>(p.rows_total) - static_cast<long long>(tile) * TILE_M;
        const uint32_t total = static_cast<uint32_t>(rows_left < TILE_M ? rows_left : TILE_M) * static_cast<uint32_t>(n_mel);
        float* otile = p.out + static_cast<long long>(tile) * TILE_M * n_mel;
        for (uint32_t idx = static_cast<uint32_t>(tid); idx < total; idx += F_EPI_WARPS * 32) {
          const uint32_t r = __umulhi(idx, mel_magic);
          const int m = static_cast<int>(idx - r * static_cast<uint32_t>(n_mel));
          otile[idx] = s_part[(m <= ms1 ? m : m + 2) * F_PART_LD + r];
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(F_EPI_WARPS * 32) : "memory");    // smem sums may be overwritten
    }
  } else if (warp == F_MMA_WARP) {
    // =========================== MMA issuer ===========================
    // all 32 lanes walk the loop (uniform control flow and registers); one elected lane issues.  Each
    // (A slab, basis block) pass is one predicated PTX sequence (tc_common.cuh umma_slab4_commit).
    {
      const uint32_t idesc = umma_idesc_f16(TILE_M, F_NH, /*fp16*/ 0);
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(a_base)), b_lo0 = umma_desc_lo(smem_u32(b_base));
      const uint32_t a_full0 = smem_u32(a_full), a_empty0 = smem_u32(a_empty), b_full0 = smem_u32(b_full), b_empty0 = smem_u32(b_empty);
      const int b_stages = p.b_stages, num_tiles = p.num_tiles;
      int sa = 0; uint32_t pa = 0; int sb = 0; uint32_t pb = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(acc_empty, acc_phase ^ 1);
        for (int s = 0; s < F_STAGES_PER_TILE; ++s) {            // s = 2 (K slab) + (0: sum -> even bins, 1: difference -> odd bins)
          mbar_wait_addr(a_full0 + 8u * sa, pa);
          const uint32_t a_hi = a_lo0 + static_cast<uint32_t>(sa) * (F_A_STAGE >> 4);
          const uint32_t a_lo = a_hi + (A_SLAB_BYTES >> 4);
          const uint32_t d = tmem_base + static_cast<uint32_t>((s & 1) * F_NH);
          mbar_wait_addr(b_full0 + 8u * sb, pb);
          tc_fence_after();
          uint32_t b = b_lo0 + static_cast<uint32_t>(sb) * (F_B_BLOCK >> 4);
          umma_slab4_commit(d, a_hi, b, idesc, s >= 2 ? 1u : 0u, 0u, 0u);                     // a_hi * b_hi
          umma_slab4_commit(d, a_lo, b, idesc, 1u, b_empty0 + 8u * sb, 0u);                   // a_lo * b_hi
          if (++sb == b_stages) { sb = 0; pb ^= 1; }
          mbar_wait_addr(b_full0 + 8u * sb, pb);
          tc_fence_after();
          b = b_lo0 + static_cast<uint32_t>(sb) * (F_B_BLOCK >> 4);
          umma_slab4_commit(d, a_hi, b, idesc, 1u, b_empty0 + 8u * sb, a_empty0 + 8u * sa);   // a_hi * b_lo
          if (++sb == b_stages) { sb = 0; pb ^= 1; }
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
      const int blocks = F_STAGES_PER_TILE * 2;
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
    // thread = (16-byte chunk c of a K slab, rows r0 + 32 i): per slab it loads samples [64 kb + 8 c, + 8) of both window
    // halves, applies the window, and writes the sum stage (-> even bins) and the difference stage (-> odd bins)
    const int ptid = tid - F_PROD_WARP0 * 32;
    const int c = ptid & 7, r0 = ptid >> 3;
    const int win = p.win;
    int sa = 0; uint32_t pa = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      uint32_t src[4];                                           // element offset of the frame's first sample (B * L < 2^32), ~0 = no row
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long R = static_cast<long long>(tile) * TILE_M + r0 + 32 * i;
        if (R < p.rows_total) {
          const long long b = R / p.frames;
          const int f = static_cast<int>(R - b * p.frames);
          src[i] = static_cast<uint32_t>(b * L + static_cast<long long>(p.hop) * f);     // frames[f, k] = x[hop f + k]
        } else {
          src[i] = 0xffffffffu;
        }
      }
      for (int kb = 0; kb < F_KB; ++kb) {
        const int n0 = kb * SLAB_K + c * 8;                      // win % 8 == 0: a chunk is all inside or all outside the window
        const bool in_a = n0 < win, in_b = n0 + F_NFFT / 2 < win;
        float ya[4][8], yb[4][8];
        {
          const float4 wa0 = *reinterpret_cast<const float4*>(s_hann + n0), wa1 = *reinterpret_cast<const float4*>(s_hann + n0 + 4);
          const float4 wb0 = *reinterpret_cast<const float4*>(s_hann + n0 + F_NFFT / 2);
          const float4 wb1 = *reinterpret_cast<const float4*>(s_hann + n0 + F_NFFT / 2 + 4);
          const float wa[8] = {wa0.x, wa0.y, wa0.z, wa0.w, wa1.x, wa1.y, wa1.z, wa1.w};
          const float wb[8] = {wb0.x, wb0.y, wb0.z, wb0.w, wb1.x, wb1.y, wb1.z, wb1.w};
          float4 xa[4][2], xb[4][2];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            xa[i][0] = xa[i][1] = xb[i][0] = xb[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* x = p.wav + src[i] + n0;
            if (src[i] != 0xffffffffu && in_a) {
              xa[i][0] = __ldg(reinterpret_cast<const float4*>(x));
              xa[i][1] = __ldg(reinterpret_cast<const float4*>(x + 4));
            }
            if (src[i] != 0xffffffffu && in_b) {
              xb[i][0] = __ldg(reinterpret_cast<const float4*>(x + F_NFFT / 2));
              xb[i][1] = __ldg(reinterpret_cast<const float4*>(x + F_NFFT / 2 + 4));
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {                          // frames * window in fp32 (input_data.py:361-365)
            const float a[8] = {xa[i][0].x, xa[i][0].y, xa[i][0].z, xa[i][0].w, xa[i][1].x, xa[i][1].y, xa[i][1].z, xa[i][1].w};
            const float b[8] = {xb[i][0].x, xb[i][0].y, xb[i][0].z, xb[i][0].w, xb[i][1].x, xb[i][1].y, xb[i][1].z, xb[i][1].w};
#pragma unroll
            for (int e = 0; e < 8; ++e) { ya[i][e] = __fmul_rn(a[e], wa[e]); yb[i][e] = __fmul_rn(b[e], wb[e]); }
          }
        }
#pragma unroll
        for (int sgn = 0; sgn < 2; ++sgn) {
          mbar_wait(&a_empty[sa], pa ^ 1);
          uint8_t* hi_slab = a_base + sa * F_A_STAGE;
          uint8_t* lo_slab = hi_slab + A_SLAB_BYTES;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = sgn == 0 ? __fadd_rn(ya[i][e], yb[i][e]) : __fsub_rn(ya[i][e], yb[i][e]);
            uint4 hv, lv;
            split8(v, hv, lv);
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == F_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// Basis images, window and the banded-mel table.  hann = host fp32 [win], mel = host fp32 [n_bins, n_mel],
// dct = host fp32 [n_mel, n_keep] of frontend_build.
int frontend_build_tc(kws_handle* h, const float* hann, const float* mel, const float* dct) {
  Frontend& fe = h->fe;
  fe.tc_ok = false;
  // shapes served by the tensor-core kernel; everything else stays on the fp32 chain
  if (fe.win % 8 || fe.hop % 4 || fe.n_fft != F_NFFT || fe.n_keep > F_DCT_LD || fe.n_mel > 128 || fe.n_mel < 2) return KWS_OK;
  const int nb = fe.n_bins;
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
  // ---- split-fp16 basis blocks: stage s = 2 kb + odd, rows n = the 256 columns (re, im) of bins 2 j + odd, ----
  // ---- k = folded sample index 64 kb + kk; the window is NOT in the basis (the producers apply it)        ----
  const size_t n_blocks = static_cast<size_t>(F_STAGES_PER_TILE) * 2;
  std::vector<__half> img(n_blocks * F_B_BLOCK / 2, __float2half_rn(0.0f));
  const double kPi = 3.14159265358979323846;
  for (int kb = 0; kb < F_KB; ++kb)
    for (int odd = 0; odd < 2; ++odd)
      for (int n = 0; n < F_NH; ++n) {
        const int bin = 2 * (n / 2) + odd;
        for (int kk = 0; kk < SLAB_K; ++kk) {
          const int k = kb * SLAB_K + kk;
          const int ph = (bin * k) % F_NFFT;
          const double ang = 2.0 * kPi * ph / F_NFFT;
          const float v = static_cast<float>((n & 1) ? -std::sin(ang) : std::cos(ang));
          const __half vh = __float2half_rn(v);
          const __half vl = __float2half_rn(v - __half2float(vh));
          const size_t blk = (static_cast<size_t>(kb) * 2 + odd) * 2;
          const size_t byte = swz_off(n, kk / 8) + (kk % 8) * 2;
          img[(blk * F_B_BLOCK + byte) / 2] = vh;
          img[((blk + 1) * F_B_BLOCK + byte) / 2] = vl;
        }
      }
  std::vector<float> hann_pad(F_NFFT, 0.0f);
  for (int i = 0; i < fe.win; ++i) hann_pad[i] = hann[i];
  std::vector<float> dct_pad(static_cast<size_t>(fe.n_mel) * F_DCT_LD, 0.0f);
  for (int n = 0; n < fe.n_mel; ++n)
    for (int k = 0; k < fe.n_keep; ++k) dct_pad[static_cast<size_t>(n) * F_DCT_LD + k] = dct[static_cast<size_t>(n) * fe.n_keep + k];
  const size_t img_bytes = img.size() * sizeof(__half);
  const size_t tab_bytes = tab.size() * sizeof(float);
  const size_t dct_bytes = dct_pad.size() * sizeof(float);
  const size_t hann_bytes = hann_pad.size() * sizeof(float);
  KWS_CUDA(h, cudaMalloc(&fe.tc_blob, img_bytes + tab_bytes + dct_bytes + hann_bytes));
  uint8_t* base = static_cast<uint8_t*>(fe.tc_blob);
  KWS_CUDA(h, cudaMemcpy(base, img.data(), img_bytes, cudaMemcpyHostToDevice));
  KWS_CUDA(h, cudaMemcpy(base + img_bytes, tab.data(), tab_bytes, cudaMemcpyHostToDevice));
  KWS_CUDA(h, cudaMemcpy(base + img_bytes + tab_bytes, dct_pad.data(), dct_bytes, cudaMemcpyHostToDevice));
  KWS_CUDA(h, cudaMemcpy(base + img_bytes + tab_bytes + dct_bytes, hann_pad.data(), hann_bytes, cudaMemcpyHostToDevice));
  fe.tc_basis = base;
  fe.tc_bin_tab = reinterpret_cast<float*>(base + img_bytes);
  fe.tc_dct = reinterpret_cast<float*>(base + img_bytes + tab_bytes);
  fe.tc_hann = reinterpret_cast<float*>(base + img_bytes + tab_bytes + dct_bytes);
  fe.tc_ok = true;
  return KWS_OK;
}

int launch_features_tc(kws_handle* h, const float* wav, int B, int kind, float* out, cudaStream_t st) {
  Frontend& fe = h->fe;
  if (!fe.tc_ok || kind == KWS_FEAT_SPEC || reinterpret_cast<uintptr_t>(wav) % 16)
    return launch_features_f32(h, wav, B, kind, out, st);
  constexpr int kMaxB = 131072;                                    // the producers address samples with 32-bit offsets
  if (B > kMaxB) {
    const size_t odim = static_cast<size_t>(fe.frames) * (kind == KWS_FEAT_MFCC ? fe.n_keep : fe.n_mel);
    for (int b0 = 0; b0 < B; b0 += kMaxB) {
      const int rc = launch_features_tc(h, wav + static_cast<size_t>(b0) * L, std::min(kMaxB, B - b0), kind, out + b0 * odim, st);
      if (rc) return rc;
    }
    return KWS_OK;
  }
  DftParams p{};
  p.wav = wav; p.out = out;
  p.b_img = fe.tc_basis;
  p.bin_tab = reinterpret_cast<const float2*>(fe.tc_bin_tab);
  p.dct = fe.tc_dct;
  p.hann = fe.tc_hann;
  p.frames = fe.frames; p.hop = fe.hop; p.win = fe.win; p.n_mel = fe.n_mel; p.n_keep = fe.n_keep;
  p.rows_total = B * fe.frames;
  p.num_tiles = (p.rows_total + TILE_M - 1) / TILE_M;
  p.mel_magic = static_cast<uint32_t>((0x100000000ull + fe.n_mel - 1) / fe.n_mel);
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
