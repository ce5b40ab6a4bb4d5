// K1 augment_mix -- replaces the Mul / tf_roll / Mul / Add chain of
// AudioProcessor.prepare_processing_graph (reference input_data.py:338-359,
// utils.py:56-73) for a whole batch:
//
//   out[b,t] = fl32(bg[b,t] * bg_vol[b]) + fl32(wav[b,(t - shift[b]) mod L] * fg_vol[b])
//
// HBM-bound (192,020 algorithmic bytes per clip).  One CTA per (clip, 4000-sample
// tile): the rolled waveform window and the noise-bank window are fetched with
// 16-byte cp.async chunks from their ALIGNED-DOWN source addresses into shared
// memory (L is a multiple of the chunk size, so a chunk never straddles the
// wrap-around of the circular shift), the misalignment (a CTA-uniform 0..3 sample
// offset) is resolved on the shared-memory side with two conflict-free 128-bit
// reads per output quad, and the result leaves as coalesced streaming float4
// stores.  __fmul_rn/__fadd_rn keep TF's separate roundings (no FMA contraction),
// so the output is bit-identical to the reference arithmetic.
#include "common.cuh"

namespace kws {

namespace {

constexpr int AUG_TILES = 4;
constexpr int AUG_TILE = L / AUG_TILES;     // 4000 output samples per CTA
constexpr int AUG_THREADS = 256;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

template <typename T> struct Quad;
template <> struct Quad<float>   { using type = float4; };
template <> struct Quad<int16_t> { using type = short4; };

__device__ __forceinline__ void unpack(const float4& q, float scale, float* w) {
  (void)scale; w[0] = q.x; w[1] = q.y; w[2] = q.z; w[3] = q.w;
}
__device__ __forceinline__ void unpack(const short4& q, float scale, float* w) {
  // np.float32(pcm) / divisor: 32768 (DecodeWav) is exact, 32767 (scipy paths) needs the true division
  w[0] = __fdiv_rn(static_cast<float>(q.x), scale);
  w[1] = __fdiv_rn(static_cast<float>(q.y), scale);
  w[2] = __fdiv_rn(static_cast<float>(q.z), scale);
  w[3] = __fdiv_rn(static_cast<float>(q.w), scale);
}

// window of 4 consecutive values starting r (CTA-uniform, 0..3) into an 8-value buffer
__device__ __forceinline__ void window4(const float* w8, int r, float* x) {
  switch (r) {
    case 0:  x[0] = w8[0]; x[1] = w8[1]; x[2] = w8[2]; x[3] = w8[3]; break;
    case 1:  x[0] = w8[1]; x[1] = w8[2]; x[2] = w8[3]; x[3] = w8[4]; break;
    case 2:  x[0] = w8[2]; x[1] = w8[3]; x[2] = w8[4]; x[3] = w8[5]; break;
    default: x[0] = w8[3]; x[1] = w8[4]; x[2] = w8[5]; x[3] = w8[6]; break;
  }
}

template <typename TIn>
__global__ void __launch_bounds__(AUG_THREADS)
augment_mix_kernel(const TIn* __restrict__ wav, float pcm_scale,
                   const int32_t* __restrict__ shift, const int32_t* __restrict__ bg_file,
                   const int32_t* __restrict__ bg_off, const float* __restrict__ bg_vol,
                   const float* __restrict__ fg_vol, const float* __restrict__ bank,
                   long long bank_len, const long long* __restrict__ file_offsets, int n_files,
                   float* __restrict__ out, int B, int clamp) {
  using QuadT = typename Quad<TIn>::type;
  constexpr int VEC = 16 / sizeof(TIn);                 // samples per 16-byte chunk
  constexpr int WAV_SLOTS = AUG_TILE + 2 * VEC;
  __shared__ __align__(16) TIn   s_wav[WAV_SLOTS];
  __shared__ __align__(16) float s_bg[AUG_TILE + 8];

  const int b = blockIdx.x / AUG_TILES;
  const int tile = blockIdx.x - b * AUG_TILES;
  if (b >= B) return;
  const int t0 = tile * AUG_TILE;
  const int tid = threadIdx.x;

  // ---- rolled waveform window: source index of output t is (t - shift) mod L ----
  int sm = (shift != nullptr ? shift[b] : 0) % L;      // all five parameter arrays NULL = identity (PCM decode only)
  if (sm < 0) sm += L;
  int j0 = t0 - sm;
  if (j0 < 0) j0 += L;
  const int jal = j0 & ~(VEC - 1);
  const int head = j0 - jal;                            // 0..VEC-1
  const int n_wchunks = (head + AUG_TILE + VEC - 1) / VEC;
  const TIn* row = wav + static_cast<size_t>(b) * L;
  for (int c = tid; c < n_wchunks; c += AUG_THREADS) {
    int src_chunk = jal / VEC + c;
    if (src_chunk >= L / VEC) src_chunk -= L / VEC;     // wrap-around stays chunk aligned
    cp_async16(&s_wav[c * VEC], row + src_chunk * VEC);
  }

  // ---- noise-bank window ----
  const int bf = bg_file != nullptr ? bg_file[b] : -1;
  const bool has_bg = (bf >= 0) && (bf < n_files) && (bank != nullptr);
  int bhead = 0;
  if (has_bg) {
    const long long start = file_offsets[bf] + static_cast<long long>(bg_off[b]) + t0;
    const long long a0 = start & ~3LL;
    bhead = static_cast<int>(start - a0);
    const int n_bchunks = (bhead + AUG_TILE + 3) / 4;
    for (int c = tid; c < n_bchunks; c += AUG_THREADS) {
      const long long idx = a0 + 4LL * c;
      if (idx >= 0 && idx + 4 <= bank_len) {
        cp_async16(&s_bg[c * 4], bank + idx);
      } else {                                          // ragged end of the bank
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const long long k = idx + e;
          s_bg[c * 4 + e] = (k >= 0 && k < bank_len) ? bank[k] : 0.0f;
        }
      }
    }
  }
  cp_async_wait_all();
  __syncthreads();

  const float fv = fg_vol != nullptr ? fg_vol[b] : 1.0f;
  const float bv = bg_vol != nullptr ? bg_vol[b] : 0.0f;
  const int wq = head >> 2, wr = head & 3;
  const int br = bhead;                                 // 0..3
  const QuadT* sq = reinterpret_cast<const QuadT*>(s_wav);
  const float4* sb = reinterpret_cast<const float4*>(s_bg);
  float4* orow = reinterpret_cast<float4*>(out + static_cast<size_t>(b) * L + t0);

  for (int i = tid; i < AUG_TILE / 4; i += AUG_THREADS) {
    float w8[8], x[4], g[4];
    unpack(sq[wq + i], pcm_scale, w8);
    unpack(sq[wq + i + 1], pcm_scale, w8 + 4);
    window4(w8, wr, x);
    if (has_bg) {
      float g8[8];
      unpack(sb[i], 1.0f, g8);
      unpack(sb[i + 1], 1.0f, g8 + 4);
      window4(g8, br, g);
    } else {
      g[0] = g[1] = g[2] = g[3] = 0.0f;                 // np.zeros background (input_data.py:498)
    }
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float fgv = __fmul_rn(x[e], fv);            // tf.multiply(wav, foreground_volume)
      const float bgv = __fmul_rn(g[e], bv);            // tf.multiply(background, volume)
      float v = __fadd_rn(bgv, fgv);                    // tf.add(background_mul, shifted)
      if (clamp) v = fminf(fmaxf(v, -1.0f), 1.0f);
      o[e] = v;
    }
    __stcs(&orow[i], make_float4(o[0], o[1], o[2], o[3]));
  }
}

}  // namespace

int launch_augment(kws_handle* h, const float* wav, const int16_t* pcm, float pcm_scale,
                   const int32_t* shift, const int32_t* bg_file, const int32_t* bg_off,
                   const float* bg_vol, const float* fg_vol, float* out, int B, int clamp,
                   cudaStream_t st) {
  if (B == 0) return KWS_OK;
  const long long* fo = reinterpret_cast<const long long*>(h->file_offsets_d);
  dim3 grid(static_cast<unsigned>(B) * AUG_TILES), block(AUG_THREADS);
  KWS_T0(h, KC_AUGMENT, st);
  if (pcm != nullptr) {
    augment_mix_kernel<int16_t><<<grid, block, 0, st>>>(pcm, pcm_scale, shift, bg_file, bg_off, bg_vol,
                                                        fg_vol, h->bank, h->bank_len, fo, h->n_files,
                                                        out, B, clamp);
  } else {
    augment_mix_kernel<float><<<grid, block, 0, st>>>(wav, 1.0f, shift, bg_file, bg_off, bg_vol, fg_vol,
                                                      h->bank, h->bank_len, fo, h->n_files, out, B, clamp);
  }
  KWS_T1(h, st);
  KWS_LAUNCH_CHECK(h);
  return KWS_OK;
}

}  // namespace kws
