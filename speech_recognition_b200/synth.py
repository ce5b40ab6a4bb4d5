"""Deterministic synthetic inputs for tests and benchmarks.

The reference's dataset, noise wavs and trained checkpoints are not available
(no network; /root/reference/.MISSING_LARGE_BLOBS), so every measurement and
parity test runs on:

* clips      -- 1 s / 16 kHz, int16-quantised then /32768 (exactly representable
                like decoded PCM, input_data.py:334-336)
* noise bank -- 6 coloured-noise "files" of 60 s, amplitude /3 (generate_noise.py:16)
* weights    -- Glorot-uniform kernels + BatchNorm statistics calibrated on a fixed
                batch of these clips (``data/synth_bn_<arch>.npz``, produced once by
                ``tools/calibrate_synth_bn.py``) so that the signal actually
                propagates through the 12 layers, as it does in a trained net.
"""
from __future__ import annotations

import os

import numpy as np

from .arch import ARCHS, weight_shapes

SEED = 59185            # input_data.py:46 (RANDOM_SEED) reused as the base seed
SAMPLES = 16000
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def make_clips(n: int, seed: int = SEED, return_pcm: bool = False):
    """n synthetic clips [n,16000] f32 (and optionally the int16 PCM)."""
    rs = np.random.RandomState(seed)
    t = (np.arange(SAMPLES, dtype=np.float64) / SAMPLES)[None, :]
    out = np.empty((n, SAMPLES), np.int16)
    step = 512
    for s0 in range(0, n, step):
        m = min(step, n - s0)
        x = rs.normal(0.0, 0.002, (m, SAMPLES))                       # (a) silence-like floor
        kind = rs.randint(0, 4, m)                                    # 0 silence, 1 burst, 2/3 tones
        start = rs.uniform(0.0, 0.5, (m, 1))
        dur = rs.uniform(0.3, 0.8, (m, 1))
        env = ((t >= start) & (t < start + dur)).astype(np.float64)
        ramp = np.clip((t - start) / 0.02, 0, 1) * np.clip((start + dur - t) / 0.05, 0, 1)
        env = env * ramp
        # (b) band-limited noise burst
        white = rs.normal(0.0, 1.0, (m, SAMPLES))
        spec = np.fft.rfft(white, axis=1)
        f = np.fft.rfftfreq(SAMPLES, 1.0 / SAMPLES)[None, :]
        lo = rs.uniform(100.0, 2000.0, (m, 1))
        hi = lo + rs.uniform(300.0, 3000.0, (m, 1))
        burst = np.fft.irfft(spec * ((f >= lo) & (f <= hi)), n=SAMPLES, axis=1)
        burst /= (np.abs(burst).max(axis=1, keepdims=True) + 1e-9)
        amp_b = rs.uniform(0.05, 0.4, (m, 1))
        # (c) 2-4 decaying sinusoids
        nsin = rs.randint(2, 5, m)
        freqs = rs.uniform(100.0, 4000.0, (m, 4))
        amps = rs.uniform(0.05, 0.35, (m, 4)) * (np.arange(4)[None, :] < nsin[:, None])
        decay = rs.uniform(1.0, 6.0, (m, 4))
        phase = rs.uniform(0, 2 * np.pi, (m, 4))
        tone = np.zeros((m, SAMPLES))
        for j in range(4):
            tone += amps[:, j:j + 1] * np.exp(-decay[:, j:j + 1] * np.maximum(t - start, 0)) * \
                np.sin(2 * np.pi * freqs[:, j:j + 1] * t + phase[:, j:j + 1])
        x += (kind[:, None] == 1) * amp_b * burst * env
        x += (kind[:, None] >= 2) * tone * env
        x -= 0.00064                                                  # train.py:16 mean
        out[s0:s0 + m] = np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)
    clips = out.astype(np.float32) * np.float32(1.0 / 32768.0)
    return (clips, out) if return_pcm else clips


def make_word_clips(n: int, n_classes: int = 12, seed: int = SEED, device="cpu", return_labels: bool = False):
    """n synthetic "keyword" clips [n,16000] f32 (torch tensor on `device`), int16-quantised like decoded PCM.

    Class 0 is silence (noise floor only, `_silence_`); class c >= 1 is a fixed three-partial chirp pattern with
    an amplitude-modulated envelope (a stand-in for one spoken word: the pattern is the word, the per-clip
    amplitude / onset / pitch jitter / phases / noise floor are the speaker).  Unlike `make_clips` (random mixtures with
    no class structure) these clips can be CLASSIFIED, which is what the label-agreement tests need: a network
    TRAINED on them (`trained_weights(arch)`, tools/train_synth_ckpt.py) is confident on its inputs the way the
    reference's trained checkpoints are on speech.  Written with torch
    ops so that 100k+ distinct clips can be generated on the GPU in a second (tests); the label of clip i is
    (seed-dependent) `labels[i]`."""
    import torch
    g = torch.Generator(device=device).manual_seed(int(seed))
    f32 = torch.float32
    labels = torch.randint(0, n_classes, (n,), generator=g, device=device)
    # per-class pattern table (host RandomState: identical on every device)
    rs = np.random.RandomState(1000 + n_classes)
    f0 = rs.uniform(200.0, 3800.0, (n_classes, 3))
    slope = rs.uniform(-800.0, 800.0, (n_classes, 3))
    rel = rs.uniform(0.3, 1.0, (n_classes, 3))
    dur = rs.uniform(0.25, 0.6, (n_classes,))
    am = rs.uniform(3.0, 12.0, (n_classes,))
    tab = lambda a: torch.as_tensor(a, dtype=f32, device=device)[labels]          # noqa: E731
    f0, slope, rel, dur, am = tab(f0), tab(slope), tab(rel), tab(dur), tab(am)
    u = lambda lo, hi, shape: lo + (hi - lo) * torch.rand(shape, generator=g, device=device, dtype=f32)   # noqa: E731
    amp = u(0.08, 0.35, (n, 1))
    start = u(0.1, 0.35, (n, 1))
    jitter = 1.0 + 0.01 * torch.randn((n, 1), generator=g, device=device, dtype=f32)
    phase = u(0.0, 2 * np.pi, (n, 3))
    t = (torch.arange(SAMPLES, device=device, dtype=f32) / SAMPLES)[None, :]
    tt = torch.clamp(t - start, min=0.0)
    env = torch.clamp(tt / 0.02, 0, 1) * torch.clamp((start + dur[:, None] - t) / 0.05, 0, 1)
    env = env * (0.75 + 0.25 * torch.sin(2 * np.pi * am[:, None] * tt))
    x = 0.002 * torch.randn((n, SAMPLES), generator=g, device=device, dtype=f32)
    tone = torch.zeros((n, SAMPLES), device=device, dtype=f32)
    for j in range(3):
        ph = 2 * np.pi * jitter * (f0[:, j:j + 1] * tt + 0.5 * slope[:, j:j + 1] * tt * tt) + phase[:, j:j + 1]
        tone += rel[:, j:j + 1] * torch.sin(ph)
    x += (labels != 0).to(f32)[:, None] * amp * env * tone / 3.0
    x -= 0.00064
    pcm = torch.clamp(torch.round(x * 32768.0), -32768, 32767)
    clips = (pcm * (1.0 / 32768.0)).to(f32)
    return (clips, labels) if return_labels else clips


def make_noise_bank(seconds: int = 60, seed: int = SEED + 7):
    """6 coloured-noise files -> (bank f32 [sum_len], file_offsets i64 [7])."""
    rs = np.random.RandomState(seed)
    n = seconds * SAMPLES
    files = []
    f = np.fft.rfftfreq(n, 1.0 / SAMPLES)
    f[0] = f[1]
    for exponent in (0.0, -0.5, -1.0, 0.5, 1.0, -0.25):               # white, pink, brown, blue, violet, ...
        spec = np.fft.rfft(rs.normal(0.0, 1.0, n)) * f ** exponent
        x = np.fft.irfft(spec, n=n)
        x = x / np.abs(x).max() / 3.0                                 # generate_noise.py:16
        files.append(np.round(x * 32768.0).astype(np.float32) * np.float32(1.0 / 32768.0))
    offsets = np.zeros(len(files) + 1, np.int64)
    offsets[1:] = np.cumsum([len(x) for x in files])
    return np.concatenate(files).astype(np.float32), offsets


def raw_synthetic_weights(arch: int, seed: int | None = None, head_gain: float | None = None,
                          attn_gain: float = 2.0):
    """Seeded Glorot-uniform kernels with *uncalibrated* BatchNorm statistics."""
    rs = np.random.RandomState(arch if seed is None else seed)
    if head_gain is None:
        head_gain = {"attn_mean": 32.0, "gap_dense": 4.0, "max_avg_dense": 2.0}.get(ARCHS[arch]["pool"], 6.0)
    w = {}
    for name, shp in weight_shapes(arch).items():
        if name.endswith("/gamma"):
            v = rs.uniform(0.5, 1.5, shp)
        elif name.endswith("/moving_variance"):
            v = rs.uniform(0.5, 1.5, shp)
        elif name.endswith("/beta") or name.endswith("/moving_mean") or name.endswith("/bias"):
            v = rs.normal(0.0, 0.1, shp)
        elif "depthwise_kernel" in name:
            v = rs.uniform(-1.0, 1.0, shp)
        else:
            rf = int(np.prod(shp[:-2])) if len(shp) > 2 else 1
            lim = np.sqrt(6.0 / (rf * shp[-2] + rf * shp[-1]))
            v = rs.uniform(-lim, lim, shp)
            if name == "dense_2/kernel" or (name == "dense_1/kernel" and ARCHS[arch]["pool"] == "max_avg_dense"):
                v = v * head_gain      # trained-classifier contrast instead of a near-uniform softmax
            if name == "dense_1/kernel" and ARCHS[arch]["pool"] != "max_avg_dense":
                v = v * attn_gain
        w[name] = v.astype(np.float32)
    return w


def trained_weights(arch: int):
    """A TRAINED checkpoint (195, 106, or 206 = the 195 architecture from another seed) keyed by Keras variable name (``data/trained_<arch>.npz``,
    stored as float16): tools/train_synth_ckpt.py trained the torch restatement of the reference model on the
    synthetic keyword task of `make_word_clips` (the reference's own checkpoints are not in the mount).
    Random weights make a chaotic, undecided network (it amplifies input and rounding noise and its softmax sits
    near ties); this one behaves like the reference's: ~100 % accurate and confident on its data."""
    path = os.path.join(_DATA, f"trained_{arch}.npz")
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run tools/train_synth_ckpt.py")
    shapes = weight_shapes(arch)
    with np.load(path) as z:
        w = {k: z[k].astype(np.float32) for k in z.files}
    assert set(w) == set(shapes) and all(w[k].shape == tuple(shapes[k]) for k in shapes)
    return {k: w[k] for k in shapes}


def synthetic_weights(arch: int, calibrated: bool = True):
    """Synthetic weights keyed by Keras variable name.  With ``calibrated`` the
    BatchNorm moving statistics come from ``data/synth_bn_<arch>.npz``."""
    key = 195 if arch == 206 else arch
    w = raw_synthetic_weights(arch)
    if arch == 1663:                      # steffeNet: raw BatchNorm statistics (ReLU6 and the residual sums keep it in range)
        return w
    if calibrated:
        path = os.path.join(_DATA, f"synth_bn_{key}.npz")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run tools/calibrate_synth_bn.py")
        with np.load(path) as z:
            # exp 206 shares the 195 architecture; its own seed only changes the kernels,
            # so the calibrated statistics are re-derived per seed in the file.
            tag = f"s{arch}/"
            for name in z.files:
                if name.startswith(tag):
                    w[name[len(tag):]] = z[name].astype(np.float32)
    return w


def make_params(n: int, noise_offsets, seed: int = SEED + 1, silence_fraction: float = 0.1):
    """Pre-drawn training-mode augmentation parameters (utils.py:8-12 defaults):
    shift in [-500,0] w.p. 0.3, background w.p. 0.3 volume U(0,0.15) (silence clips:
    w.p. 0.9 volume U(0,0.3)), foreground w.p. 0.3 volume 1+U(-0.15,0.15).
    Draw order = input_data.py:457-514; vectorised per clip with one RandomState."""
    rs = np.random.RandomState(seed)
    lengths = np.diff(np.asarray(noise_offsets))
    p = dict(time_shift=np.zeros(n, np.int32), bg_index=np.zeros(n, np.int32),
             bg_offset=np.zeros(n, np.int32), bg_volume=np.zeros(n, np.float32),
             fg_volume=np.ones(n, np.float32))
    is_silence = np.arange(n) % max(1, int(round(1.0 / silence_fraction))) == 0
    for i in range(n):
        if rs.uniform(0.0, 1.0) < 0.3:
            p["time_shift"][i] = rs.randint(-500, 1)
        bi = rs.randint(len(lengths))
        p["bg_index"][i] = bi
        p["bg_offset"][i] = rs.randint(0, int(lengths[bi]) - SAMPLES)
        if rs.uniform(0, 1) < 0.3:
            p["bg_volume"][i] = rs.uniform(0, 0.15)
        elif is_silence[i] and rs.uniform(0, 1) < 0.9:
            p["bg_volume"][i] = rs.uniform(0, 0.3)
        if is_silence[i]:
            p["fg_volume"][i] = 0.0
        else:
            if rs.uniform(0, 1) < 0.3:
                p["fg_volume"][i] = 1.0 + rs.uniform(-0.15, 0.15)
            rs.uniform(0, 1)                                          # flip draw (flip_frequency=0)
    return p
