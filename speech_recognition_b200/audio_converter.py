"""``AudioConverter`` -- the native contrib_audio front end of the reference (audio.py:7-28):
``decode_wav`` -> ``audio_spectrogram(window 480, stride 160, magnitude_squared=True)`` ->
``mfcc(dct_coefficient_count=40)``.  The reference runs it one file at a time through a TF
session (``load(fn, sess)``); here a batch of decoded clips goes through libkws.so in one call.
The TF kernels (spectrogram.cc, mfcc.cc, mfcc_mel_filterbank.cc, mfcc_dct.cc) are not vendored in
the reference, so the arithmetic follows their published algorithm -- parity unpinned."""
from __future__ import annotations

import numpy as np

from .engine import Engine


class AudioConverter:
    def __init__(self, desired_samples=16000, window_size_samples=480, window_stride_samples=160,
                 engine: Engine | None = None, device: int = 0, sample_rate: int = 16000):
        if desired_samples != 16000:
            raise ValueError("libkws is built for 1 s / 16 kHz clips (desired_samples=16000)")
        self.engine = engine if engine is not None else Engine(device=device)
        self.window_size_samples, self.window_stride_samples = window_size_samples, window_stride_samples
        self.sample_rate = sample_rate

    def _configure(self):
        self.engine.frontend_config_contrib(self.window_size_samples, self.window_stride_samples, self.sample_rate,
                                            20.0, 4000.0, 40, 40)

    def load_batch(self, clips: np.ndarray) -> np.ndarray:
        """clips f32 [B,16000] (decoded PCM / 32768) -> mfcc f32 [B, 98, 40]."""
        import torch
        self._configure()
        x = torch.from_numpy(np.ascontiguousarray(clips, np.float32)).to(f"cuda:{self.engine.device}")
        return self.engine.features(x, "mfcc").cpu().numpy()

    def load(self, clip: np.ndarray, sess=None) -> np.ndarray:
        """One decoded clip -> [1, 98, 40] like ``sess.run(self.mfcc, ...)`` (audio.py:25-28); ``sess`` is ignored."""
        return self.load_batch(np.asarray(clip, np.float32).reshape(1, -1))
