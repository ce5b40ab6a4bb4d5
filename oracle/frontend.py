"""Oracle, stage 1b: STFT -> |.| -> mel -> log -> MFCC (TEST INFRASTRUCTURE ONLY).

Restates input_data.py:361-381 (``tf.contrib.signal.stft`` /
``linear_to_mel_weight_matrix`` / ``mfccs_from_log_mel_spectrograms`` of TF 1.4,
un-vendored).  Constants follow the reference's own serialized graph
(logs_195 GraphDef nodes ``stft/*``, ``linear_to_mel_weight_matrix/*``,
``mfccs_from_log_mel_spectrograms/*``); ``tests/golden/graph_frontend.npz`` holds
the outputs of that graph evaluated node-by-node and pins this file.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def hann_window_periodic(frame_length: int) -> np.ndarray:
    """Graph nodes stft/hann_window/*: fp32, periodic=True.
    w[i] = 0.5 - 0.5*cos(fl32(2*pi) * i / fl32(N')),  N' = N + (1 - N%2) - 1 (= N for even N)."""
    even = 1 - frame_length % 2
    n = F32(frame_length + even - 1)
    two_pi = F32(6.2831854820251465)
    i = np.arange(frame_length, dtype=F32)
    arg = ((two_pi * i).astype(F32) / n).astype(F32)
    c = np.cos(arg, dtype=F32)
    return (F32(0.5) - (F32(0.5) * c).astype(F32)).astype(F32)


def frame(x: np.ndarray, frame_length: int, frame_step: int) -> np.ndarray:
    """tf.contrib.signal.frame, pad_end=False: frames[f,i] = x[f*step + i]."""
    L = x.shape[-1]
    n = max(0, 1 + (L - frame_length) // frame_step)
    idx = (np.arange(n)[:, None] * frame_step + np.arange(frame_length)[None, :])
    return x[..., idx]


def next_pow2(n: int) -> int:
    p = 1
    while p < n:
        p *= 2
    return p


def spectrogram(x: np.ndarray, frame_length: int = 480, frame_step: int = 160,
                fft_dtype=np.float64) -> np.ndarray:
    """input_data.py:361-366: |RFFT(pad(frames * hann, fft_length))|, magnitude (not power).
    fft_length=None => next power of two (512 for 480; graph node stft/Const).
    ``fft_dtype``: float64 gives the correctly-rounded reference value; float32
    mimics TF's fp32 FFT up to its (unknowable) butterfly order."""
    x = np.asarray(x, F32)
    w = hann_window_periodic(frame_length)
    fr = (frame(x, frame_length, frame_step) * w).astype(F32)
    n_fft = next_pow2(frame_length)
    pad = [(0, 0)] * (fr.ndim - 1) + [(0, n_fft - frame_length)]
    fr = np.pad(fr, pad)
    spec = np.fft.rfft(fr.astype(fft_dtype), n=n_fft, axis=-1)
    return np.abs(spec).astype(F32)


def hertz_to_mel(f):
    return 1127.0 * np.log(1.0 + np.asarray(f, np.float64) / 700.0)


def linear_to_mel_weight_matrix(num_mel_bins: int, num_spectrogram_bins: int = 257,
                                sample_rate: float = 16000.0,
                                lower_edge_hertz: float = 80.0,
                                upper_edge_hertz: float = 7600.0) -> np.ndarray:
    """input_data.py:367-373; float64 internally, cast to f32 at the end
    (graph: every linear_to_mel_weight_matrix/* Const is float64, final Cast)."""
    nyquist = sample_rate / 2.0
    lin = np.linspace(0.0, nyquist, num_spectrogram_bins, dtype=np.float64)[1:]
    spec_mel = hertz_to_mel(lin)[:, None]
    edges = np.linspace(hertz_to_mel(lower_edge_hertz), hertz_to_mel(upper_edge_hertz),
                        num_mel_bins + 2, dtype=np.float64)
    lower, center, upper = edges[None, :-2], edges[None, 1:-1], edges[None, 2:]
    lower_slopes = (spec_mel - lower) / (center - lower)
    upper_slopes = (upper - spec_mel) / (upper - center)
    w = np.maximum(0.0, np.minimum(lower_slopes, upper_slopes))
    w = np.pad(w, [(1, 0), (0, 0)])
    return w.astype(F32)


def log_mel(spec: np.ndarray, mel_w: np.ndarray, acc_dtype=np.float64) -> np.ndarray:
    """input_data.py:374-378: log(spec @ W + 1e-6) in fp32."""
    mel = (spec.astype(acc_dtype) @ mel_w.astype(acc_dtype)).astype(F32)
    return np.log((mel + F32(1e-6)).astype(F32), dtype=F32)


def mfcc_from_log_mel(lm: np.ndarray, fft_dtype=np.float64) -> np.ndarray:
    """tf.contrib.signal.mfccs_from_log_mel_spectrograms (TF 1.4), graph nodes
    mfccs_from_log_mel_spectrograms/*: DCT-II through an RFFT of length 2M:
        scale[k] = 2*exp(-j*pi*k/(2M));  y = Re(RFFT(pad(x, 2M))[:M] * scale) * rsqrt(2M)
    """
    lm = np.asarray(lm, F32)
    M = lm.shape[-1]
    k = np.arange(M, dtype=F32)
    arg = ((F32(-3.1415927410125732) * k).astype(F32) / (F32(2.0) * F32(M))).astype(F32)
    scale = (2.0 * np.exp(1j * arg.astype(np.float64)))
    spec = np.fft.rfft(lm.astype(fft_dtype), n=2 * M, axis=-1)[..., :M]
    y = np.real(spec * scale).astype(F32)
    rs = F32(1.0) / np.sqrt(F32(2 * M), dtype=F32)
    return (y * rs).astype(F32)


def features(x: np.ndarray, *, window_size_samples: int = 480,
             window_stride_samples: int = 160, dct_coefficient_count: int = 40,
             num_log_mel_features: int | None = None, sample_rate: int = 16000,
             kind: str = "mfcc", fft_dtype=np.float64) -> np.ndarray:
    """x [B,16000] f32 -> 'spec' [B,98,257] | 'logmel' [B,98,M] | 'mfcc' [B,98,K]
    (K = num_log_mel_features, input_data.py:379-381)."""
    spec = spectrogram(x, window_size_samples, window_stride_samples, fft_dtype)
    if kind == "spec":
        return spec
    W = linear_to_mel_weight_matrix(dct_coefficient_count, spec.shape[-1], sample_rate)
    lm = log_mel(spec, W)
    if kind == "logmel":
        return lm
    mf = mfcc_from_log_mel(lm, fft_dtype)
    K = num_log_mel_features if num_log_mel_features is not None else dct_coefficient_count
    return mf[..., :K]


# ------------------------------------------------------------------------------------------------
# Native contrib_audio flavour (audio.py:15-23; exp-106 graph nodes AudioSpectrogram / Mfcc).
# TF 1.4's C++ kernels are not vendored in the reference; this restates their published algorithm
# (tensorflow/core/kernels/spectrogram.cc, mfcc.cc, mfcc_mel_filterbank.cc, mfcc_dct.cc), which
# computes in double precision and casts to float at the end.  PARITY UNPINNED: no golden vector
# of these ops exists in the reference.
# ------------------------------------------------------------------------------------------------
def contrib_audio_spectrogram(x, window_size=480, stride=160, magnitude_squared=True):
    """x [B, L] f32 -> [B, frames, fft/2+1]; periodic Hann in double, zero-pad to the next power of two."""
    x = np.asarray(x, F32).astype(np.float64)
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(window_size) / window_size)
    fr = frame(x, window_size, stride) * w
    n_fft = next_pow2(window_size)
    spec = np.fft.rfft(fr, n=n_fft, axis=-1)
    p = spec.real ** 2 + spec.imag ** 2
    return (p if magnitude_squared else np.sqrt(p)).astype(F32)


def contrib_mel_filterbank(input_length=257, sample_rate=16000.0, channels=40, lower=20.0, upper=4000.0):
    """(band_mapper, weights, start_index, end_index) of MfccMelFilterbank::Initialize."""
    mel = lambda f: 1127.0 * np.log1p(f / 700.0)  # noqa: E731
    mel_low, mel_hi = mel(lower), mel(upper)
    spacing = (mel_hi - mel_low) / (channels + 1)
    center = mel_low + spacing * (np.arange(channels + 1) + 1)
    hz_per_sbin = 0.5 * sample_rate / (input_length - 1)
    start, end = int(1.5 + lower / hz_per_sbin), int(upper / hz_per_sbin)
    band = np.full(input_length, -2, np.int64)
    wts = np.zeros(input_length, np.float64)
    ch = 0
    for i in range(input_length):
        melf = mel(i * hz_per_sbin)
        if i < start or i > end:
            continue
        while ch < channels and center[ch] < melf:
            ch += 1
        band[i] = ch - 1
        c = band[i]
        wts[i] = (center[c + 1] - melf) / (center[c + 1] - center[c]) if c >= 0 else \
            (center[0] - melf) / (center[0] - mel_low)
    return band, wts, start, end


def contrib_mfcc(power_spec, sample_rate=16000, dct_coefficient_count=40, channels=40, lower=20.0,
                 upper=4000.0, return_log_mel=False):
    """contrib_audio.mfcc on a squared-magnitude spectrogram [.., frames, bins] -> [.., frames, dct_count]."""
    p = np.asarray(power_spec, F32).astype(np.float64)
    bins = p.shape[-1]
    band, wts, start, end = contrib_mel_filterbank(bins, sample_rate, channels, lower, upper)
    out = np.zeros(p.shape[:-1] + (channels,), np.float64)
    for i in range(start, end + 1):                       # MfccMelFilterbank::Compute
        sv = np.sqrt(p[..., i])
        wv = sv * wts[i]
        c = band[i]
        if c >= 0:
            out[..., c] += wv
        if c + 1 < channels:
            out[..., c + 1] += sv - wv
    lm = np.log(np.maximum(out, 1e-12))                   # kFilterbankFloor
    if return_log_mel:
        return lm.astype(F32)
    n = channels                                          # MfccDct
    cos = np.sqrt(2.0 / n) * np.cos(np.arange(dct_coefficient_count)[:, None] * (np.pi / n) * (np.arange(n)[None, :] + 0.5))
    return (lm @ cos.T).astype(F32)
