"""Oracle, speed-TTA view: phase-vocoder time stretch (TEST INFRASTRUCTURE ONLY).

create_tta_set.py:10-22 builds the "slow" test set that make_submission.py:131-140 (use_speed_tta) consumes:

    data = np.float32(data) / 32767
    data = librosa.effects.time_stretch(data, 0.9)
    data = data[-16000:]
    wf.write(out_fn, rate, np.int16(data * 32767))

librosa is a third-party dependency that is absent from /root/reference and not installable here (no pinned
version in the reference either; README.md:31 only names the function; the code is from January 2018 =
librosa 0.5.1).  This file restates the PUBLISHED algorithm of librosa 0.5.x `effects.time_stretch` =
`core.stft` (n_fft 2048, hop 512, periodic Hann, centered with reflect padding, complex64 output of a
double-precision FFT) -> `core.phase_vocoder` (linear magnitude interpolation, phase advance accumulation in the
dtype of `np.angle(D[:, 0])` = float32) -> `core.istft` (single-precision inverse FFT, windowed overlap-add,
division by the window sum-square where it exceeds `tiny`, trim n_fft // 2 at both ends).  PARITY UNPINNED: there
is no librosa output anywhere in the reference to check it against; `tests/test_stretch_cpu.py` cross-checks the
STFT / ISTFT halves against torch.stft / torch.istft and the rate-1.0 identity.
"""
from __future__ import annotations

import numpy as np

N_FFT = 2048
HOP = N_FFT // 4


def hann_periodic(n: int = N_FFT) -> np.ndarray:
    """scipy.signal.get_window('hann', n, fftbins=True), float64."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def stft(y: np.ndarray) -> np.ndarray:
    """librosa.core.stft defaults -> complex64 [1 + n_fft/2, n_frames] (Fortran order in librosa; values only here)."""
    y = np.asarray(y)
    win = hann_periodic()
    yp = np.pad(y, N_FFT // 2, mode="reflect")
    n_frames = 1 + (len(yp) - N_FFT) // HOP
    frames = np.stack([yp[i * HOP:i * HOP + N_FFT] for i in range(n_frames)], axis=1)   # [n_fft, n_frames], y's dtype
    spec = np.fft.fft(win[:, None] * frames, axis=0)[: 1 + N_FFT // 2]                   # float64 window -> double FFT
    return spec.astype(np.complex64)


def phase_vocoder(D: np.ndarray, rate: float) -> np.ndarray:
    """librosa.core.phase_vocoder (0.5.x), dtype for dtype."""
    n_fft = 2 * (D.shape[0] - 1)
    hop = n_fft // 4
    time_steps = np.arange(0, D.shape[1], rate, dtype=np.float64)
    out = np.zeros((D.shape[0], len(time_steps)), D.dtype)
    phi_advance = np.linspace(0, np.pi * hop, D.shape[0])            # float64
    phase_acc = np.angle(D[:, 0])                                    # float32 for complex64 input
    D = np.pad(D, [(0, 0), (0, 2)], mode="constant")
    for t, step in enumerate(time_steps):
        cols = D[:, int(step):int(step + 2)]
        alpha = np.mod(step, 1.0)
        mag = (1.0 - alpha) * np.abs(cols[:, 0]) + alpha * np.abs(cols[:, 1])           # float64
        out[:, t] = mag * np.exp(1.0j * phase_acc)                   # complex64 exp, product rounded to complex64
        dphase = np.angle(cols[:, 1]) - np.angle(cols[:, 0]) - phi_advance              # float64
        dphase = dphase - 2.0 * np.pi * np.round(dphase / (2.0 * np.pi))
        phase_acc += phi_advance + dphase                            # in place: rounded to float32 every step
    return out


def window_sumsquare(n_frames: int, dtype=np.float32) -> np.ndarray:
    """librosa.filters.window_sumsquare(window='hann', norm=None): accumulated in `dtype`."""
    n = N_FFT + HOP * (n_frames - 1)
    x = np.zeros(n, dtype=dtype)
    win_sq = hann_periodic() ** 2
    for i in range(n_frames):
        s = i * HOP
        x[s:min(n, s + N_FFT)] += win_sq[: max(0, min(N_FFT, n - s))]
    return x


def istft(S: np.ndarray, dtype=np.float32) -> np.ndarray:
    """librosa.core.istft defaults (center=True, length=None)."""
    n_frames = S.shape[1]
    win = hann_periodic()
    y = np.zeros(N_FFT + HOP * (n_frames - 1), dtype=dtype)
    for i in range(n_frames):
        spec = S[:, i].flatten()
        spec = np.concatenate((spec, spec[-2:0:-1].conj()), 0)
        # scipy.fftpack.ifft keeps single precision for complex64 input
        ytmp = win * np.fft.ifft(spec.astype(np.complex128)).real.astype(np.float32)
        y[i * HOP:i * HOP + N_FFT] = y[i * HOP:i * HOP + N_FFT] + ytmp
    wss = window_sumsquare(n_frames, dtype=dtype)
    nz = wss > np.finfo(dtype).tiny
    y[nz] /= wss[nz]
    return y[N_FFT // 2: -(N_FFT // 2)]


def time_stretch(y: np.ndarray, rate: float) -> np.ndarray:
    """librosa.effects.time_stretch(y, rate)."""
    y = np.asarray(y, np.float32)
    return istft(phase_vocoder(stft(y), rate), dtype=y.dtype)


def stretched_len(n: int, rate: float) -> int:
    n_frames = 1 + n // HOP
    return HOP * (len(np.arange(0, n_frames, rate)) - 1)


def create_tta_clip(pcm: np.ndarray, tta_speed: float = 0.9, samples: int = 16000) -> np.ndarray:
    """create_tta_set.py:16-22 for one clip of int16 PCM -> int16 PCM of the slowed clip (last `samples` samples)."""
    data = np.float32(pcm) / 32767
    data = time_stretch(data, tta_speed)
    data = data[-samples:]
    return np.int16(data * 32767)


def create_tta_batch(pcm: np.ndarray, tta_speed: float = 0.9) -> np.ndarray:
    return np.stack([create_tta_clip(p, tta_speed) for p in pcm])
