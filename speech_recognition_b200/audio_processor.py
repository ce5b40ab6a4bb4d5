"""``AudioProcessor`` -- the reference's batch assembly surface (input_data.py:162-175,
383-541) on top of libkws.so.

Dataset indexing / WAV file IO of the reference (input_data.py:182-309) is host-side
bookkeeping outside the hot path: here the partitions are handed over as in-memory
arrays.  ``get_data`` keeps the reference's signature, draws the augmentation
parameters from the *global* ``np.random`` state in the reference's order
(input_data.py:457-514), and runs the whole batch through one C-ABI call instead of
one ``sess.run`` per clip.
"""
from __future__ import annotations

import numpy as np

from .engine import Engine

SILENCE_LABEL = '_silence_'
SILENCE_INDEX = 0


def draw_augmentation_params(data_index, background_data, model_settings, how_many, offset,
                             background_frequency, background_volume_range, foreground_frequency,
                             foreground_volume_range, time_shift_frequency, time_shift_range, mode,
                             pseudo_frequency=0.0, flip_frequency=0.0, silence_volume_range=0.0):
    """The parameter draw of get_data (input_data.py:457-514) without the compute."""
    ms = model_settings
    cand_clips, cand_labels = data_index[mode]
    pseudo = data_index['pseudo']
    n_cand = len(cand_labels)
    sample_count = n_cand if how_many == -1 else max(0, min(how_many, n_cand - offset))
    desired = ms['desired_samples']
    use_background = bool(background_data) and (mode == 'training')
    pick_deterministically = (mode != 'training')
    idx = np.zeros(sample_count, np.int64)
    from_pseudo = np.zeros(sample_count, bool)
    p = dict(time_shift=np.zeros(sample_count, np.int32), bg_index=np.full(sample_count, -1, np.int32),
             bg_offset=np.zeros(sample_count, np.int32), bg_volume=np.zeros(sample_count, np.float32),
             fg_volume=np.zeros(sample_count, np.float32))
    labels = np.zeros(sample_count, np.int64)
    for n, i in enumerate(range(offset, offset + sample_count)):
        if how_many == -1 or pick_deterministically:
            sample_index, is_pseudo = i, False
        elif np.random.uniform(0, 1) < pseudo_frequency:
            sample_index, is_pseudo = np.random.randint(len(pseudo[1])), True
        else:
            sample_index, is_pseudo = np.random.randint(n_cand), False
        label = int(pseudo[1][sample_index] if is_pseudo else cand_labels[sample_index])
        if np.random.uniform(0.0, 1.0) < time_shift_frequency:
            time_shift = np.random.randint(time_shift_range[0], time_shift_range[1] + 1)
        else:
            time_shift = 0
        if use_background:
            bg_index = np.random.randint(len(background_data))
            bg_offset = np.random.randint(0, len(background_data[bg_index]) - desired)
            if np.random.uniform(0, 1) < background_frequency:
                bg_volume = np.random.uniform(0, background_volume_range)
            else:
                bg_volume = 0.0
                if label == SILENCE_INDEX and np.random.uniform(0, 1) < 0.9:
                    bg_volume = np.random.uniform(0, silence_volume_range)
            p['bg_index'][n], p['bg_offset'][n] = bg_index, bg_offset
        else:
            bg_volume = 0.0
        if label == SILENCE_INDEX:
            fg_volume = 0.0
        else:
            fg_volume = 1.0
            if np.random.uniform(0, 1) < foreground_frequency:
                fg_volume = 1.0 + np.random.uniform(-foreground_volume_range, foreground_volume_range)
            if np.random.uniform(0, 1) < flip_frequency:
                fg_volume *= -1.0
        idx[n], from_pseudo[n], labels[n] = sample_index, is_pseudo, label
        p['time_shift'][n], p['bg_volume'][n], p['fg_volume'][n] = time_shift, bg_volume, fg_volume
    return idx, from_pseudo, labels, p



class AudioProcessor:
    def __init__(self, model_settings, output_representation='raw', data=None, background_data=None,
                 engine: Engine | None = None, device: int = 0, precision="tc"):
        """data: {'training'|'validation'|'testing'|'pseudo': (clips f32 [n,16000], label_index int [n])}
        background_data: list of 1-D float32 arrays (decoded _background_noise_ wavs)."""
        import torch
        self.model_settings = model_settings
        self.output_representation = output_representation
        if output_representation not in ('raw', 'spec', 'mfcc', 'mfcc_and_raw'):
            raise ValueError(output_representation)
        self.engine = engine if engine is not None else Engine(device=device, precision=precision)
        self.data_index = {m: None for m in ('training', 'validation', 'testing', 'pseudo')}
        for mode, (clips, labels) in (data or {}).items():
            self.data_index[mode] = (np.ascontiguousarray(clips, np.float32), np.asarray(labels, np.int64))
        self.background_data = [np.ascontiguousarray(b, np.float32).ravel() for b in (background_data or [])]
        if self.background_data:
            offs = np.zeros(len(self.background_data) + 1, np.int64)
            offs[1:] = np.cumsum([len(b) for b in self.background_data])
            bank = torch.from_numpy(np.concatenate(self.background_data)).to(f"cuda:{self.engine.device}")
            self.engine.set_noise_bank(bank, offs)
        ms = model_settings
        self.engine.frontend_config(ms['window_size_samples'], ms['window_stride_samples'],
                                    ms['dct_coefficient_count'], ms['num_log_mel_features'],
                                    80.0, 7600.0, ms['sample_rate'])      # input_data.py:368

    def set_size(self, mode):
        d = self.data_index[mode]
        return 0 if d is None else len(d[1])

    def draw(self, *args, **kw):
        """The parameter draw of get_data without the compute (see draw_augmentation_params)."""
        return draw_augmentation_params(self.data_index, self.background_data, self.model_settings, *args, **kw)

    def get_data(self, how_many, offset, background_frequency, background_volume_range,
                 foreground_frequency, foreground_volume_range, time_shift_frequency, time_shift_range,
                 mode, sess=None, pseudo_frequency=0.0, flip_frequency=0.0, silence_volume_range=0.0):
        """Returns (data float64 [n, dim], one-hot labels float64 [n, label_count]) -- or
        ([mfcc, raw], labels) for 'mfcc_and_raw' -- like input_data.py:538-541.  ``sess`` is unused."""
        ms = self.model_settings
        idx, from_pseudo, labels, p = self.draw(
            how_many, offset, background_frequency, background_volume_range, foreground_frequency,
            foreground_volume_range, time_shift_frequency, time_shift_range, mode, pseudo_frequency,
            flip_frequency, silence_volume_range)
        n = len(idx)
        clips = np.empty((n, ms['desired_samples']), np.float32)
        cand = self.data_index[mode][0]
        if n:
            clips[~from_pseudo] = cand[idx[~from_pseudo]]
            if from_pseudo.any():
                clips[from_pseudo] = self.data_index['pseudo'][0][idx[from_pseudo]]
        onehot = np.zeros((n, ms['label_count']))
        onehot[np.arange(n), labels] = 1

        def run(kind):
            if n == 0:
                dim = ms['desired_samples'] if kind == 'raw' else int(np.prod(self.engine.feature_shape(kind)))
                return np.zeros((0, dim))
            out = self.engine.get_data_host(clips, p['time_shift'], p['bg_index'], p['bg_offset'],
                                            p['bg_volume'], p['fg_volume'], kind=kind)
            return out.astype(np.float64)       # the reference fills np.zeros((n, dim)) == float64
        rep = self.output_representation
        if rep == 'mfcc_and_raw':
            # one augmentation run feeds both outputs, as sess.run([mfcc_, background_clamp_]) does (input_data.py:520-531)
            raw = run('raw')
            if n == 0:
                return [run('mfcc'), raw], onehot
            mfcc = self.engine.features_host(raw.astype(np.float32), kind='mfcc').astype(np.float64)
            return [mfcc, raw], onehot
        return run(rep), onehot
