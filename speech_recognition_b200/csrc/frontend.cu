// Stage 1b front end: STFT -> magnitude -> mel -> log -> DCT-II (MFCC).
// Replaces tf.contrib.signal.stft / linear_to_mel_weight_matrix / tensordot / log /
// mfccs_from_log_mel_spectrograms of AudioProcessor.prepare_processing_graph
// (reference input_data.py:361-381; constants from the logs_195 GraphDef).
//
// The STFT is a DFT-as-GEMM: the periodic Hann window is folded into a
// [window, 2*bins] (cos, -sin) basis, frames are gathered implicitly from the
// waveform by the GEMM's A loader (the clip is read once although frames overlap
// 3x), the epilogue takes the magnitude.  mel and DCT are two more GEMMs with the
// log fused into the mel epilogue.  This file holds the basis construction and the
// fp32 (KWS_PREC_FP32) chain; tc_frontend.cu holds the tcgen05 chain.
#include <cmath>

#include "common.cuh"
#include "gemm_f32.cuh"

namespace kws {

int frontend_build_tc(kws_handle* h, const float* hann, const float* mel, const float* dct);   // tc_frontend.cu

int frontend_build(kws_handle* h, int win, int hop, int n_mel, int n_keep, float f_lo, float f_hi,
                   int sample_rate, int flavour) {
  if (win <= 0 || hop <= 0 || win > L || n_mel <= 0 || n_keep <= 0 || n_keep > n_mel)
    return fail(h, KWS_EINVAL, "bad front-end configuration");
  Frontend& fe = h->fe;
  if (fe.blob) { cudaFree(fe.blob); fe.blob = nullptr; }
  if (fe.tc_blob) { cudaFree(fe.tc_blob); fe.tc_blob = nullptr; }
  fe = Frontend();
  fe.win = win; fe.hop = hop; fe.n_mel = n_mel; fe.n_keep = n_keep; fe.flavour = flavour;
  int n_fft = 1;
  while (n_fft < win) n_fft *= 2;                       // fft_length=None -> next pow2 (graph: stft/Const=512)
  fe.n_fft = n_fft; fe.n_bins = n_fft / 2 + 1;
  fe.frames = 1 + (L - win) / hop;                      // model.py:1803-1808

  // periodic Hann in fp32 exactly as the graph nodes stft/hann_window/*
  std::vector<float> hann(win);
  if (flavour == 1) {
    // contrib_audio AudioSpectrogram (TF spectrogram.cc GetPeriodicHann): double precision, 0.5 - 0.5 cos(2 pi i / N)
    const double kPi2 = 2.0 * 3.14159265358979323846;
    for (int i = 0; i < win; ++i) hann[i] = static_cast<float>(0.5 - 0.5 * std::cos(kPi2 * i / win));
  } else {
    const float two_pi = 6.2831854820251465f;
    const float denom = static_cast<float>(win + (1 - win % 2) - 1);
    for (int i = 0; i < win; ++i) {
      const float arg = (two_pi * static_cast<float>(i)) / denom;
      const float c = static_cast<float>(std::cos(static_cast<double>(arg)));
      hann[i] = 0.5f - 0.5f * c;
    }
  }
  const int nb = fe.n_bins;
  std::vector<float> host;
  auto pad4 = [&]() { while (host.size() % 4) host.push_back(0.f); };
  // DFT basis [win, 2*nb]: col 2j = cos(2 pi j k / n_fft) w[k], col 2j+1 = -sin(...) w[k]
  const size_t o_basis = host.size();
  host.resize(o_basis + static_cast<size_t>(win) * 2 * nb);
  const double kPi = 3.14159265358979323846;
  for (int k = 0; k < win; ++k)
    for (int j = 0; j < nb; ++j) {
      const int ph = static_cast<int>((static_cast<long long>(j) * k) % n_fft);
      const double a = 2.0 * kPi * ph / n_fft;
      host[o_basis + static_cast<size_t>(k) * 2 * nb + 2 * j] = static_cast<float>(std::cos(a) * hann[k]);
      host[o_basis + static_cast<size_t>(k) * 2 * nb + 2 * j + 1] = static_cast<float>(-std::sin(a) * hann[k]);
    }
  pad4();
  // mel matrix in float64, cast at the end (graph nodes linear_to_mel_weight_matrix/*)
  const size_t o_mel = host.size();
  host.resize(o_mel + static_cast<size_t>(nb) * n_mel, 0.f);
  if (flavour == 1) {
    // contrib_audio Mfcc (TF mfcc_mel_filterbank.cc): HTK-style triangular bank on sqrt(power), channel
    // centres equally spaced in mel between f_lo and f_hi, each bin split between its two neighbouring
    // channels with weights w and 1 - w; bins below start_index / above end_index are dropped.
    auto mel = [](double f) { return 1127.0 * std::log1p(f / 700.0); };
    const double mel_low = mel(f_lo), mel_hi = mel(f_hi);
    const double spacing = (mel_hi - mel_low) / (n_mel + 1);
    std::vector<double> center(n_mel + 1);
    for (int i = 0; i < n_mel + 1; ++i) center[i] = mel_low + spacing * (i + 1);
    const double hz_per_sbin = 0.5 * sample_rate / (nb - 1);
    const int start_index = static_cast<int>(1.5 + f_lo / hz_per_sbin);
    const int end_index = static_cast<int>(f_hi / hz_per_sbin);
    int channel = 0;
    for (int i = 0; i < nb; ++i) {
      const double melf = mel(i * hz_per_sbin);
      if (i < start_index || i > end_index) continue;
      while (channel < n_mel && center[channel] < melf) ++channel;
      const int ch = channel - 1;                              // band_mapper_[i]
      const double w = ch >= 0 ? (center[ch + 1] - melf) / (center[ch + 1] - center[ch])
                               : (center[0] - melf) / (center[0] - mel_low);
      if (ch >= 0) host[o_mel + static_cast<size_t>(i) * n_mel + ch] = static_cast<float>(w);
      if (ch + 1 < n_mel) host[o_mel + static_cast<size_t>(i) * n_mel + ch + 1] = static_cast<float>(1.0 - w);
    }
  } else {
    auto mel = [](double f) { return 1127.0 * std::log(1.0 + f / 700.0); };
    const double nyq = sample_rate / 2.0;
    const double m_lo = mel(f_lo), m_hi = mel(f_hi);
    std::vector<double> edges(n_mel + 2);
    for (int i = 0; i < n_mel + 2; ++i) edges[i] = m_lo + (m_hi - m_lo) * i / (n_mel + 1);
    for (int bin = 1; bin < nb; ++bin) {                 // row 0 (DC) stays zero (Pad [[1,0],[0,0]])
      const double sm = mel(nyq * bin / (nb - 1));
      for (int j = 0; j < n_mel; ++j) {
        const double lo = edges[j], ce = edges[j + 1], up = edges[j + 2];
        const double w = std::fmax(0.0, std::fmin((sm - lo) / (ce - lo), (up - sm) / (up - ce)));
        host[o_mel + static_cast<size_t>(bin) * n_mel + j] = static_cast<float>(w);
      }
    }
  }
  pad4();
  // DCT-II basis [n_mel, n_keep]: 2 cos(pi k (2n+1) / 2M) * rsqrt(2M)
  const size_t o_dct = host.size();
  host.resize(o_dct + static_cast<size_t>(n_mel) * n_keep);
  if (flavour == 1) {
    // TF mfcc_dct.cc: out[k] = sum_n sqrt(2 / M) cos(k pi / M (n + 0.5)) in[n]
    const double fnorm = std::sqrt(2.0 / n_mel), arg = kPi / n_mel;
    for (int n = 0; n < n_mel; ++n)
      for (int k = 0; k < n_keep; ++k)
        host[o_dct + static_cast<size_t>(n) * n_keep + k] = static_cast<float>(fnorm * std::cos(k * arg * (n + 0.5)));
  } else {
    const double sc = 1.0 / std::sqrt(2.0 * n_mel);
    for (int n = 0; n < n_mel; ++n)
      for (int k = 0; k < n_keep; ++k)
        host[o_dct + static_cast<size_t>(n) * n_keep + k] =
            static_cast<float>(2.0 * std::cos(kPi * k * (2 * n + 1) / (2.0 * n_mel)) * sc);
  }
  pad4();
  KWS_CUDA(h, cudaMalloc(&fe.blob, host.size() * sizeof(float)));
  KWS_CUDA(h, cudaMemcpy(fe.blob, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
  fe.dft_basis = fe.blob + o_basis;
  fe.mel_w = fe.blob + o_mel;
  fe.dct_w = fe.blob + o_dct;
  int rc = frontend_build_tc(h, hann.data(), host.data() + o_mel, host.data() + o_dct);
  if (rc) return rc;
  fe.configured = true;
  return KWS_OK;
}

int launch_features_f32(kws_handle* h, const float* wav, int B, int kind, float* out, cudaStream_t st) {
  Frontend& fe = h->fe;
  const int CH = 2048;                                  // clips per pass (bounds the workspace)
  const int nb = fe.n_bins;
  if (kind != KWS_FEAT_SPEC) {
    const size_t need = static_cast<size_t>(std::min(B, CH)) * fe.frames * nb * sizeof(float);
    int rc = ensure_bytes(h, reinterpret_cast<void**>(&h->spec_ws), &h->spec_ws_bytes, need);
    if (rc) return rc;
  }
  if (kind == KWS_FEAT_MFCC) {
    const size_t need = static_cast<size_t>(std::min(B, CH)) * fe.frames * fe.n_mel * sizeof(float);
    int rc = ensure_bytes(h, reinterpret_cast<void**>(&h->mel_ws), &h->mel_ws_bytes, need);
    if (rc) return rc;
  }
  const int out_dim = kind == KWS_FEAT_SPEC ? nb : (kind == KWS_FEAT_LOGMEL ? fe.n_mel : fe.n_keep);
  for (int b0 = 0; b0 < B; b0 += CH) {
    const int n = std::min(CH, B - b0);
    const int M = n * fe.frames;
    float* o = out + static_cast<size_t>(b0) * fe.frames * out_dim;
    float* spec = kind == KWS_FEAT_SPEC ? o : h->spec_ws;
    KWS_T0(h, KC_DFT, st);
    launch_gemm_f32(LoadFrames{wav + static_cast<size_t>(b0) * L, fe.frames, fe.hop}, fe.dft_basis, M,
                    2 * nb, fe.win, EpiMagnitude{spec, nb, fe.flavour == 1 && kind == KWS_FEAT_SPEC}, st);
    KWS_T1(h, st);
    KWS_LAUNCH_CHECK(h);
    if (kind == KWS_FEAT_SPEC) continue;
    float* lm = kind == KWS_FEAT_LOGMEL ? o : h->mel_ws;
    KWS_T0(h, KC_MELDCT, st);
    launch_gemm_f32(LoadPlain{spec, nb}, fe.mel_w, M, fe.n_mel, nb, EpiLog{lm, fe.n_mel, fe.flavour == 1}, st);
    KWS_T1(h, st);
    KWS_LAUNCH_CHECK(h);
    if (kind == KWS_FEAT_LOGMEL) continue;
    KWS_T0(h, KC_MELDCT, st);
    launch_gemm_f32(LoadPlain{lm, fe.n_mel}, fe.dct_w, M, fe.n_keep, fe.n_mel, EpiStore{o, fe.n_keep}, st);
    KWS_T1(h, st);
    KWS_LAUNCH_CHECK(h);
  }
  return KWS_OK;
}

}  // namespace kws
