"""Oracle of the speed-TTA view (oracle/stretch.py = the published librosa 0.5.x time_stretch): cross-checks of its
halves against torch's independent STFT / ISTFT, the rate-1.0 identity, output lengths and the properties of a
phase vocoder (a stationary tone keeps its frequency, an onset is delayed by 1 / rate)."""
import numpy as np
import torch

from oracle import stretch
from speech_recognition_b200 import synth


def test_stft_istft_against_torch():
    x = synth.make_clips(3, seed=11)
    win = torch.from_numpy(stretch.hann_periodic())
    for i in range(3):
        S = stretch.stft(x[i])
        T = torch.stft(torch.from_numpy(x[i]).double(), 2048, hop_length=512, window=win, center=True,
                       pad_mode="reflect", return_complex=True).numpy()
        assert S.shape == T.shape == (1025, 32)
        assert np.abs(S - T).max() < 1e-5 * np.abs(T).max()
        y = stretch.istft(S)
        yt = torch.istft(torch.from_numpy(T), 2048, hop_length=512, window=win, center=True).numpy()
        assert y.shape == (15872,)                       # 512 * 31: librosa's istft(length=None) drops the ragged tail
        assert np.abs(y - yt[:15872]).max() < 2e-6
        assert np.abs(y[1024:-1024] - x[i][1024:15872 - 1024]).max() < 2e-6   # perfect reconstruction away from the edges


def test_rate_one_is_identity_and_lengths():
    x = synth.make_clips(2, seed=12)
    y = stretch.time_stretch(x[0], 1.0)
    assert np.abs(y[1024:-1024] - x[0][1024:15872 - 1024]).max() < 2e-4      # float32 phase accumulator noise only
    assert stretch.stretched_len(16000, 0.9) == 17920 == len(stretch.time_stretch(x[1], 0.9))
    pcm = np.int16(np.round(x * 32767))
    out = stretch.create_tta_batch(pcm)
    assert out.dtype == np.int16 and out.shape == (2, 16000)


def test_tone_keeps_pitch_and_amplitude():
    t = np.arange(16000) / 16000.0
    x = (0.25 * np.sin(2 * np.pi * 1000.0 * t)).astype(np.float32)
    y = stretch.time_stretch(x, 0.9)
    mid = y[4000:14000]
    spec = np.abs(np.fft.rfft(mid * np.hanning(len(mid))))
    f = np.fft.rfftfreq(len(mid), 1 / 16000.0)
    assert abs(f[spec.argmax()] - 1000.0) < 2.0
    # a plain (not phase-locked) vocoder keeps the inter-bin phase relation of FRAME 0 -- here the reflect-padded edge
    # frame -- so the three main-lobe bins of this bin-centred tone no longer add up coherently: amplitude 2/3, exactly
    assert abs(np.sqrt(2 * np.mean(mid ** 2)) - 0.25 * 2 / 3) < 2e-3
    # an onset at 0.5 s moves to 0.5 / 0.9 s
    x2 = x * (t >= 0.5)
    y2 = stretch.time_stretch(x2, 0.9)
    env = np.abs(y2)
    onset = np.argmax(env > 0.125) / 16000.0
    assert abs(onset - 0.5 / 0.9) < 0.03
