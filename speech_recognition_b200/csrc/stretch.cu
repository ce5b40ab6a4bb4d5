// Speed-TTA view: phase-vocoder time stretch of a batch of clips (reference create_tta_set.py:10-22, consumed by
// make_submission.py:131-140):
//
//     data = np.float32(pcm) / 32767;  data = librosa.effects.time_stretch(data, rate);  out = np.int16(data[-16000:] * 32767)
//
// librosa is not vendored by the reference; the algorithm is the published librosa 0.5.x one (see oracle/stretch.py,
// parity unpinned): STFT (n_fft 2048, hop 512, periodic Hann, reflect-padded, double-precision transform rounded to
// complex64) -> phase vocoder (linear magnitude interpolation, float32 phase accumulator advanced in double) -> ISTFT
// (single-precision transform, windowed overlap-add, division by the window sum-square, n_fft / 2 trimmed at both ends).
//
// One CTA per clip, two per SM.  Two real frames share one complex transform (z = f0 + i f1 on the way in, Hermitian
// packing Z = A + i B on the way out), so a clip is 16 double-precision and 18 single-precision radix-2 FFTs of 2048
// points in shared memory.  The STFT (magnitude + angle, 2 x 32 x 1025 floats) parks in a per-CTA slice of an L2-resident
// scratch buffer; each thread owns fixed frequency bins, so the phase accumulator of a bin lives in a register across the
// 36 output frames; the overlap-add buffer (19,968 floats) stays in shared memory until the clip is written out.
// CUDA-core work, a few MFLOP per clip: it only has to outrun the network that consumes its output.
#include <cmath>
#include <vector>

#include "common.cuh"

namespace kws {

namespace {

constexpr int ST_N = 2048, ST_HOP = 512, ST_BINS = ST_N / 2 + 1, ST_LOG2 = 11;
constexpr int ST_THREADS = 256;
constexpr int ST_FRAMES_IN = 1 + L / ST_HOP;                     // 32 (centered STFT of 16000 samples)
constexpr int ST_MAX_FRAMES_OUT = 48;                            // rate >= 0.67
constexpr int ST_SCRATCH_FLOATS = 2 * ST_FRAMES_IN * ST_BINS;    // magnitude + angle per CTA

struct StretchParams {
  const int16_t* pcm;        // [B, 16000]
  int16_t* out;              // [B, 16000]
  int B;
  float divisor;             // 32767 (create_tta_set.py:18,22)
  int n_out;                 // output frames = len(arange(0, 32, rate))
  const double* steps;       // [n_out] time steps (float64, as np.arange produces them)
  const double* tw;          // [1024] (cos, -sin)(2 pi k / 2048) interleaved
  const double* win;         // [2048] periodic Hann, float64
  const float* wss;          // [2048 + 512 (n_out - 1)] window sum-square, accumulated in float32
  float* scratch;            // [grid][ST_SCRATCH_FLOATS]
};

__device__ __forceinline__ int bitrev11(int i) { return static_cast<int>(__brev(static_cast<unsigned>(i)) >> (32 - ST_LOG2)); }

// in-place radix-2 decimation-in-time FFT of 2048 complex points whose input was stored in bit-reversed order;
// kInverse conjugates the twiddles (the caller applies 1 / N)
template <typename T2, typename T, bool kInverse>
__device__ __forceinline__ void fft2048(T2* z, const double* __restrict__ tw, int tid) {
#pragma unroll 1
  for (int s = 0; s < ST_LOG2; ++s) {
    const int half = 1 << s;
    __syncthreads();
#pragma unroll
    for (int q = 0; q < ST_N / 2 / ST_THREADS; ++q) {
      const int j = tid + q * ST_THREADS;
      const int pos = j & (half - 1);
      const int i0 = ((j >> s) << (s + 1)) + pos, i1 = i0 + half;
      const int k = pos << (ST_LOG2 - 1 - s);                    // twiddle exp(-2 pi i pos / (2 half)) = table[k]
      const T wr = static_cast<T>(__ldg(&tw[2 * k])), wi = static_cast<T>(kInverse ? -__ldg(&tw[2 * k + 1]) : __ldg(&tw[2 * k + 1]));
      const T2 a = z[i0], b = z[i1];
      const T tr = b.x * wr - b.y * wi, ti = b.x * wi + b.y * wr;
      z[i0] = T2{a.x + tr, a.y + ti};
      z[i1] = T2{a.x - tr, a.y - ti};
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(ST_THREADS, 2) time_stretch_kernel(const StretchParams p) {
  extern __shared__ __align__(16) unsigned char st_smem[];
  double2* zd = reinterpret_cast<double2*>(st_smem);                                  // 32 KB: forward transforms
  float2* zf = reinterpret_cast<float2*>(st_smem);                                    // (first 16 KB again: inverse transforms)
  float* ola = reinterpret_cast<float*>(st_smem + ST_N * sizeof(double2));           // overlap-add, 2048 + 512 (n_out - 1) floats
  const int tid = threadIdx.x;
  float* mag = p.scratch + static_cast<size_t>(blockIdx.x) * ST_SCRATCH_FLOATS;       // [32][1025]
  float* ang = mag + ST_FRAMES_IN * ST_BINS;
  const int ola_len = ST_N + ST_HOP * (p.n_out - 1);

  for (int clip = blockIdx.x; clip < p.B; clip += gridDim.x) {
    const int16_t* x = p.pcm + static_cast<size_t>(clip) * L;
    // ---------------- STFT: frames 2 f and 2 f + 1 in one double-precision transform ----------------
    for (int fp = 0; fp < ST_FRAMES_IN / 2; ++fp) {
      __syncthreads();                                            // the previous pair's spectrum has been read
      for (int i = tid; i < ST_N; i += ST_THREADS) {
        double v[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          int s = ST_HOP * (2 * fp + h) + i - ST_N / 2;           // sample index before the reflect padding
          if (s < 0) s = -s;
          if (s >= L) s = 2 * (L - 1) - s;
          const float xv = __fdiv_rn(static_cast<float>(x[s]), p.divisor);            // np.float32(data) / 32767
          v[h] = __ldg(&p.win[i]) * static_cast<double>(xv);      // float64 window x float32 frame
        }
        zd[bitrev11(i)] = double2{v[0], v[1]};
      }
      fft2048<double2, double, false>(zd, p.tw, tid);
      for (int k = tid; k < ST_BINS; k += ST_THREADS) {
        const double2 a = zd[k], b = zd[(ST_N - k) & (ST_N - 1)];
        // F0 = (Z[k] + conj(Z[N-k])) / 2, F1 = (Z[k] - conj(Z[N-k])) / (2 i); stored as complex64
        const float r0 = static_cast<float>(0.5 * (a.x + b.x)), i0 = static_cast<float>(0.5 * (a.y - b.y));
        const float r1 = static_cast<float>(0.5 * (a.y + b.y)), i1 = static_cast<float>(0.5 * (b.x - a.x));
        mag[(2 * fp) * ST_BINS + k] = hypotf(r0, i0);
        ang[(2 * fp) * ST_BINS + k] = atan2f(i0, r0);
        mag[(2 * fp + 1) * ST_BINS + k] = hypotf(r1, i1);
        ang[(2 * fp + 1) * ST_BINS + k] = atan2f(i1, r1);
      }
    }
    for (int i = tid; i < ola_len; i += ST_THREADS) ola[i] = 0.0f;
    __syncthreads();                                              // mag / ang of this CTA are complete (same threads read them: block-scope visibility)
    // ---------------- phase vocoder + ISTFT, two output frames per single-precision transform ----------------
    constexpr int BPT = (ST_BINS + ST_THREADS - 1) / ST_THREADS;  // bins per thread: k = tid + 256 m
    float acc[BPT];                                               // phase accumulators (float32, as np.angle(D[:, 0]))
#pragma unroll
    for (int m = 0; m < BPT; ++m) {
      const int k = tid + m * ST_THREADS;
      acc[m] = k < ST_BINS ? ang[k] : 0.0f;
    }
    const double phi_step = M_PI * static_cast<double>(ST_HOP) / static_cast<double>(ST_BINS - 1);   // np.linspace(0, pi * hop, 1025)
    for (int tp = 0; tp < (p.n_out + 1) / 2; ++tp) {
#pragma unroll
      for (int m = 0; m < BPT; ++m) {
        const int k = tid + m * ST_THREADS;
        if (k < ST_BINS) {
          float2 spec[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int t = 2 * tp + h;
            if (t < p.n_out) {
              const double step = __ldg(&p.steps[t]);
              const int c0 = static_cast<int>(step);              // columns int(step), int(step) + 1 of the zero-padded STFT
              const double alpha = step - floor(step);            // np.mod(step, 1.0)
              const float m0 = c0 < ST_FRAMES_IN ? mag[c0 * ST_BINS + k] : 0.0f, m1 = c0 + 1 < ST_FRAMES_IN ? mag[(c0 + 1) * ST_BINS + k] : 0.0f;
              const float a0 = c0 < ST_FRAMES_IN ? ang[c0 * ST_BINS + k] : 0.0f, a1 = c0 + 1 < ST_FRAMES_IN ? ang[(c0 + 1) * ST_BINS + k] : 0.0f;
              const double mg = (1.0 - alpha) * static_cast<double>(m0) + alpha * static_cast<double>(m1);
              float sn, cs;
              sincosf(acc[m], &sn, &cs);                          // np.exp(1j * phase_acc) in complex64
              spec[h] = float2{static_cast<float>(mg * static_cast<double>(cs)), static_cast<float>(mg * static_cast<double>(sn))};
              const double phi = (k == ST_BINS - 1) ? M_PI * ST_HOP : phi_step * k;   // linspace hits its end point exactly
              double dphase = static_cast<double>(a1 - a0) - phi; // float32 difference, then float64
              dphase = dphase - 2.0 * M_PI * rint(dphase / (2.0 * M_PI));              // np.round: half to even
              acc[m] = static_cast<float>(static_cast<double>(acc[m]) + (phi + dphase));   // in-place add on a float32 array
            } else {
              spec[h] = float2{0.0f, 0.0f};
            }
          }
          // Hermitian packing: ifft(A + i B) = a + i b for real a, b.  The reference builds spec = concat(S, conj(S[-2:0:-1]))
          // and keeps ifft(spec).real: the imaginary parts of bins 0 and N/2 contribute to the imaginary part only, so they
          // are dropped here (packed, they would leak into the OTHER frame's real part).
          float2 A = spec[0], Bv = spec[1];
          if (k == 0 || k == ST_BINS - 1) { A.y = 0.0f; Bv.y = 0.0f; }   // they only feed the discarded imaginary part of ifft(spec)
          zf[bitrev11(k)] = float2{A.x - Bv.y, A.y + Bv.x};
          if (k > 0 && k < ST_BINS - 1) zf[bitrev11(ST_N - k)] = float2{A.x + Bv.y, Bv.x - A.y};   // conj(A) + i conj(B)
        }
      }
      fft2048<float2, float, true>(zf, p.tw, tid);
      // windowed overlap-add in frame order: y = float32(float64(y) + win * ytmp)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int t = 2 * tp + h;
        if (t < p.n_out) {
          for (int i = tid; i < ST_N; i += ST_THREADS) {
            const float2 v = zf[i];
            const float r = (h == 0 ? v.x : v.y) * (1.0f / ST_N);  // ifft(...).real, single precision
            float* y = ola + ST_HOP * t + i;
            *y = static_cast<float>(static_cast<double>(*y) + __ldg(&p.win[i]) * static_cast<double>(r));
          }
          __syncthreads();                                        // frame t + 1 overlaps frame t at other threads' samples
        }
      }
    }
    // ---------------- normalise, trim, keep the last 16000 samples, int16 ----------------
    const int trimmed = ola_len - ST_N;                           // y[n_fft / 2 : -n_fft / 2]
    const int first = ST_N / 2 + (trimmed - L);                   // data[-16000:]
    int16_t* o = p.out + static_cast<size_t>(clip) * L;
    for (int i = tid; i < L; i += ST_THREADS) {
      float y = ola[first + i];
      const float w = __ldg(&p.wss[first + i]);
      if (w > 1.17549435e-38f) y = __fdiv_rn(y, w);               // util.tiny(float32)
      const float s = __fmul_rn(y, 32767.0f);                     // data * 32767 in float32
      o[i] = static_cast<int16_t>(static_cast<int>(s));           // np.int16(...): truncation toward zero
    }
    __syncthreads();
  }
}

}  // namespace

int launch_time_stretch(kws_handle* h, const int16_t* pcm, int B, double rate, float divisor, int16_t* out, cudaStream_t st) {
  if (B == 0) return KWS_OK;
  if (!(rate > 0.0) || rate > 1.0) return fail(h, KWS_EUNSUPPORTED, "time stretch supports 0 < rate <= 1 (create_tta_set.py uses 0.9)");
  // time steps exactly as np.arange(0, n_frames, rate, dtype=float64): start + i * step
  std::vector<double> steps;
  {
    const double r = rate;                                        // a double, like the Python float the reference passes: float(0.9) * 10 < 9
    const int n = static_cast<int>(std::ceil(static_cast<double>(ST_FRAMES_IN) / r));
    for (int i = 0; i < n; ++i) steps.push_back(i * r);
  }
  const int n_out = static_cast<int>(steps.size());
  if (n_out > ST_MAX_FRAMES_OUT) return fail(h, KWS_EUNSUPPORTED, "time stretch rate too small");
  const int ola_len = ST_N + ST_HOP * (n_out - 1);
  if (ola_len - ST_N < L) return fail(h, KWS_EUNSUPPORTED, "stretched clip shorter than one clip");
  const int grid = std::min(B, 2 * h->num_sms);
  if (h->stretch_rate != rate || !h->stretch_ws) {                // tables of this rate (built once)
    std::vector<double> tab(2 * (ST_N / 2) + ST_N + n_out);
    double* tw = tab.data(); double* win = tw + ST_N; double* stp = win + ST_N;
    for (int k = 0; k < ST_N / 2; ++k) { tw[2 * k] = std::cos(2.0 * M_PI * k / ST_N); tw[2 * k + 1] = -std::sin(2.0 * M_PI * k / ST_N); }
    for (int i = 0; i < ST_N; ++i) win[i] = 0.5 - 0.5 * std::cos(2.0 * M_PI * i / ST_N);   // get_window('hann', fftbins=True)
    for (int i = 0; i < n_out; ++i) stp[i] = steps[i];
    std::vector<float> wss(ola_len, 0.0f);                        // librosa.filters.window_sumsquare, dtype float32
    for (int f = 0; f < n_out; ++f)
      for (int i = 0; i < ST_N && ST_HOP * f + i < ola_len; ++i)
        wss[ST_HOP * f + i] = static_cast<float>(static_cast<double>(wss[ST_HOP * f + i]) + win[i] * win[i]);
    const size_t tab_bytes = tab.size() * sizeof(double), wss_bytes = (wss.size() * sizeof(float) + 15) / 16 * 16;
    const size_t scratch_bytes = static_cast<size_t>(2 * h->num_sms) * ST_SCRATCH_FLOATS * sizeof(float);
    if (h->stretch_ws) { cudaFree(h->stretch_ws); h->stretch_ws = nullptr; }
    KWS_CUDA(h, cudaMalloc(&h->stretch_ws, tab_bytes + wss_bytes + scratch_bytes));
    KWS_CUDA(h, cudaMemcpyAsync(h->stretch_ws, tab.data(), tab_bytes, cudaMemcpyHostToDevice, st));
    KWS_CUDA(h, cudaMemcpyAsync(static_cast<char*>(h->stretch_ws) + tab_bytes, wss.data(), wss.size() * sizeof(float),
                                cudaMemcpyHostToDevice, st));
    KWS_CUDA(h, cudaStreamSynchronize(st));                       // the host vectors go out of scope
    h->stretch_rate = rate; h->stretch_n_out = n_out;
  }
  StretchParams p{};
  char* base = static_cast<char*>(h->stretch_ws);
  p.tw = reinterpret_cast<const double*>(base);
  p.win = p.tw + ST_N;
  p.steps = p.win + ST_N;
  const size_t tab_bytes = (2 * (ST_N / 2) + ST_N + n_out) * sizeof(double);
  p.wss = reinterpret_cast<const float*>(base + tab_bytes);
  p.scratch = reinterpret_cast<float*>(base + tab_bytes + (static_cast<size_t>(ola_len) * sizeof(float) + 15) / 16 * 16);
  p.pcm = pcm; p.out = out; p.B = B; p.divisor = divisor; p.n_out = n_out;
  const size_t smem = ST_N * sizeof(double2) + static_cast<size_t>(ola_len) * sizeof(float);
  if (!h->stretch_attr_done) {
    KWS_CUDA(h, cudaFuncSetAttribute(time_stretch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    h->stretch_attr_done = true;                                  // (two CTAs per SM up to 113 KB: rates >= 0.89)
  }
  if (smem > 160 * 1024) return fail(h, KWS_EUNSUPPORTED, "time stretch rate too small for the overlap-add buffer");
  KWS_T0(h, KC_OTHER, st);
  time_stretch_kernel<<<grid, ST_THREADS, smem, st>>>(p);
  KWS_T1(h, st);
  KWS_LAUNCH_CHECK(h);
  return KWS_OK;
}

}  // namespace kws
