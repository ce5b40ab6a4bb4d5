"""Oracle, stage 1a: waveform augmentation (TEST INFRASTRUCTURE ONLY).

Restates, op for op in fp32:
  * ``utils.tf_roll``                         (utils.py:56-73)
  * the mix in ``prepare_processing_graph``   (input_data.py:338-359)
  * the per-clip parameter draw of ``get_data`` (input_data.py:457-514)
  * WAV sample decoding conventions          (input_data.py:334-336 -> 1/32768;
    make_submission_on_rpi.py:97, create_pseudo_with_thresh.py:48 -> 1/32767)
"""
from __future__ import annotations

import numpy as np

DESIRED_SAMPLES = 16000


def decode_pcm16(pcm: np.ndarray, desired_samples: int = DESIRED_SAMPLES,
                 scale: str = "tf") -> np.ndarray:
    """int16 PCM [B, L] -> f32 [B, desired_samples], zero-pad / crop on the right.

    ``scale='tf'``   : x / 32768  (TF DecodeWav, input_data.py:334-336)
    ``scale='scipy'``: x / 32767  (make_submission_on_rpi.py:97)
    """
    pcm = np.asarray(pcm, dtype=np.int16)
    if pcm.ndim == 1:
        pcm = pcm[None]
    out = np.zeros((pcm.shape[0], desired_samples), np.float32)
    n = min(desired_samples, pcm.shape[1])
    if scale == "tf":
        out[:, :n] = pcm[:, :n].astype(np.float32) * np.float32(1.0 / 32768.0)
    elif scale == "scipy":
        out[:, :n] = pcm[:, :n].astype(np.float32) / np.float32(32767)
    else:
        raise ValueError(scale)
    return out


def tf_roll(x: np.ndarray, shift: int) -> np.ndarray:
    """utils.py:56-73.  Both branches reduce to ``np.roll(x, shift)`` along the
    sample axis: roll_left  (shift>=0): concat(a[L-s:], a[:L-s]), s = shift % L
                 roll_right (shift<0) : concat(a[s:], a[:s]),     s = (-shift) % L
    i.e. out[t] = x[(t - shift) mod L] with wrap-around (NOT zero padding, despite
    the comment at input_data.py:342)."""
    x = np.asarray(x)
    L = x.shape[-1]
    shift = int(shift)
    if shift >= 0:
        s = shift % L
        return np.concatenate([x[..., L - s:], x[..., :L - s]], axis=-1)
    s = (-shift) % L
    return np.concatenate([x[..., s:], x[..., :s]], axis=-1)


def augment_mix(wav: np.ndarray, time_shift: np.ndarray, background: np.ndarray,
                background_volume: np.ndarray, foreground_volume: np.ndarray,
                clamp: bool = False) -> np.ndarray:
    """input_data.py:338-359 for a batch.

    out[b,t] = fl32(bg[b,t]*bv[b]) + roll(fl32(wav[b]*fv[b]), shift[b])[t]
    Every op is a separate fp32 rounding (TF Mul, roll, Mul, Add are separate
    ops).  ``clamp`` reproduces the exp-106-era ``clip_by_value(-1, 1)``
    (removed at HEAD, input_data.py:356).
    """
    wav = np.asarray(wav, np.float32)
    background = np.asarray(background, np.float32)
    B, L = wav.shape
    fv = np.asarray(foreground_volume, np.float32).reshape(B, 1)
    bv = np.asarray(background_volume, np.float32).reshape(B, 1)
    scaled = (wav * fv).astype(np.float32)
    shifted = np.empty_like(scaled)
    ts = np.asarray(time_shift).reshape(B)
    for b in range(B):
        shifted[b] = tf_roll(scaled[b], int(ts[b]))
    bg_mul = (background * bv).astype(np.float32)
    out = (bg_mul + shifted).astype(np.float32)
    if clamp:
        out = np.clip(out, np.float32(-1.0), np.float32(1.0))
    return out


def gather_background(noise_bank: np.ndarray, file_offsets: np.ndarray,
                      bg_index: np.ndarray, bg_offset: np.ndarray,
                      desired_samples: int = DESIRED_SAMPLES) -> np.ndarray:
    """background_samples[offset:offset+16000] for each clip (input_data.py:482-488).
    ``noise_bank`` is the concatenation of the background wavs, ``file_offsets``
    [n_files+1] their start positions.  bg_index < 0 means "no background"
    (np.zeros, input_data.py:498)."""
    B = len(bg_index)
    out = np.zeros((B, desired_samples), np.float32)
    for b in range(B):
        if bg_index[b] >= 0:
            s = int(file_offsets[bg_index[b]]) + int(bg_offset[b])
            out[b] = noise_bank[s:s + desired_samples]
    return out


def draw_params(rs, *, how_many: int, offset: int, n_candidates: int,
                candidate_is_silence, mode: str,
                background_frequency: float, background_volume_range: float,
                foreground_frequency: float, foreground_volume_range: float,
                time_shift_frequency: float, time_shift_range,
                background_lengths=None, n_pseudo: int = 0,
                pseudo_is_silence=None, pseudo_frequency: float = 0.0,
                flip_frequency: float = 0.0, silence_volume_range: float = 0.0,
                desired_samples: int = DESIRED_SAMPLES):
    """Per-clip parameter draw in the EXACT order of input_data.py:457-514.

    ``rs`` is anything with ``uniform``/``randint`` of np.random's signature (the
    reference uses the global ``np.random``).  Returns a dict of arrays:
    sample_index i64, from_pseudo bool, time_shift i32, bg_index i32 (-1 = none),
    bg_offset i32, bg_volume f32, fg_volume f32.
    """
    if how_many == -1:
        sample_count = n_candidates
    else:
        sample_count = max(0, min(how_many, n_candidates - offset))   # :435-438
    use_background = bool(background_lengths is not None and len(background_lengths)) \
        and (mode == 'training')                                       # :453
    pick_deterministically = (mode != 'training')                      # :454
    out = dict(
        sample_index=np.zeros(sample_count, np.int64),
        from_pseudo=np.zeros(sample_count, bool),
        time_shift=np.zeros(sample_count, np.int32),
        bg_index=np.full(sample_count, -1, np.int32),
        bg_offset=np.zeros(sample_count, np.int32),
        bg_volume=np.zeros(sample_count, np.float32),
        fg_volume=np.zeros(sample_count, np.float32),
    )
    for n, i in enumerate(range(offset, offset + sample_count)):      # :457
        if how_many == -1 or pick_deterministically:                   # :459
            sample_index, pseudo = i, False
        else:
            if rs.uniform(0, 1) < pseudo_frequency:                    # :463
                sample_index, pseudo = int(rs.randint(n_pseudo)), True
            else:
                sample_index, pseudo = int(rs.randint(n_candidates)), False
        is_silence = bool(pseudo_is_silence[sample_index] if pseudo
                          else candidate_is_silence[sample_index])
        if rs.uniform(0.0, 1.0) < time_shift_frequency:               # :471
            time_shift = int(rs.randint(time_shift_range[0], time_shift_range[1] + 1))
        else:
            time_shift = 0
        bg_index, bg_offset = -1, 0
        if use_background:                                             # :481
            bg_index = int(rs.randint(len(background_lengths)))
            bg_offset = int(rs.randint(
                0, int(background_lengths[bg_index]) - desired_samples))
            if rs.uniform(0, 1) < background_frequency:                # :489
                background_volume = rs.uniform(0, background_volume_range)
            else:
                background_volume = 0.0
                if is_silence and rs.uniform(0, 1) < 0.9:              # :494-496
                    background_volume = rs.uniform(0, silence_volume_range)
        else:
            background_volume = 0.0                                    # :498-499
        if is_silence:                                                 # :503
            foreground_volume = 0.0
        else:
            foreground_volume = 1.0
            if rs.uniform(0, 1) < foreground_frequency:                # :508
                foreground_volume = 1.0 + rs.uniform(
                    -foreground_volume_range, foreground_volume_range)
            if rs.uniform(0, 1) < flip_frequency:                      # :512
                foreground_volume *= -1.0
        out['sample_index'][n] = sample_index
        out['from_pseudo'][n] = pseudo
        out['time_shift'][n] = time_shift
        out['bg_index'][n] = bg_index
        out['bg_offset'][n] = bg_offset
        out['bg_volume'][n] = np.float32(background_volume)   # fed to a tf.float32 placeholder
        out['fg_volume'][n] = np.float32(foreground_volume)
    return out
