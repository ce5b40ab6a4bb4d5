"""Thin object wrapper around a kws_t handle.  Device-pointer methods take torch
CUDA tensors (torch is plumbing: allocation, streams, torch.distributed); *_host
methods take NumPy arrays and go through the host-buffer C entry points."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import KwsError, FEAT_RAW, FEAT_SPEC, FEAT_LOGMEL, FEAT_MFCC, PREC_FP32, PREC_TC  # noqa: F401

SAMPLES = 16000
_KIND = {"raw": FEAT_RAW, "spec": FEAT_SPEC, "logmel": FEAT_LOGMEL, "mfcc": FEAT_MFCC}


def _views(views):
    views = list(views)
    n = len(views)
    sh = (C.c_int32 * n)(*[int(v[0]) for v in views])
    ga = (C.c_float * n)(*[float(v[1]) for v in views])
    return sh, ga, n


def _hp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Engine:
    def __init__(self, device: int = 0, max_rows: int = 2048, precision: str | int = "tc"):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.kws_create(C.byref(h), int(device), int(max_rows))
        if rc != 0:
            raise KwsError(f"kws_create failed ({rc}): {self.lib.kws_last_error(None).decode()}")
        self.h = h
        self.device = device
        self._keep = {}
        self.set_precision(precision)

    # -- plumbing --
    def _check(self, rc):
        if rc != 0:
            raise KwsError(f"libkws error {rc}: {self.lib.kws_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.kws_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _stream():
        import torch
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    @staticmethod
    def _dp(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None

    def set_precision(self, precision):
        p = {"fp32": PREC_FP32, "tc": PREC_TC}.get(precision, precision)
        self._check(self.lib.kws_set_precision(self.h, int(p)))
        self.precision = p

    def set_fusion(self, on: bool):
        """conv1d_1 + block 1 as one kernel (default) or two (tensor-core tier)."""
        self._check(self.lib.kws_set_fusion(self.h, int(bool(on))))

    @property
    def launch_count(self) -> int:
        return int(self.lib.kws_launch_count(self.h))

    KERNEL_CLASSES = ("augment", "dft", "mel_dct", "slice_conv1", "dw_pw_blocks", "head", "other")

    def timing_enable(self, on=True):
        self._check(self.lib.kws_timing_enable(self.h, int(bool(on))))

    def timing_read(self):
        """{class: (total_ms, launches)} since the last read (synchronises)."""
        ms = (C.c_double * 18)()
        cnt = (C.c_int64 * 18)()
        self._check(self.lib.kws_timing_read(self.h, ms, cnt, 18))
        out = {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.KERNEL_CLASSES)}
        self.last_block_ms = [(float(ms[7 + i]), int(cnt[7 + i])) for i in range(11)]   # per dw+pw block (tc tier)
        return out

    # -- stage 1a --
    def set_noise_bank(self, bank_t, file_offsets):
        """bank_t: torch CUDA f32 tensor (kept alive here); file_offsets: int64 [n_files+1]."""
        fo = np.ascontiguousarray(file_offsets, np.int64)
        self._keep["bank"] = bank_t
        self._check(self.lib.kws_set_noise_bank(self.h, self._dp(bank_t), fo.ctypes.data_as(C.POINTER(C.c_int64)),
                                                len(fo) - 1))

    def augment(self, wav_t, shift_t, bg_file_t, bg_off_t, bg_vol_t, fg_vol_t, out_t=None, clamp=False,
                pcm_divisor=None):
        import torch
        B = wav_t.shape[0]
        if out_t is None:
            out_t = torch.empty((B, SAMPLES), dtype=torch.float32, device=wav_t.device)
        if wav_t.dtype == torch.int16:
            self._check(self.lib.kws_augment_pcm16(self.h, self._dp(wav_t), float(pcm_divisor or 32768.0),
                                                   self._dp(shift_t), self._dp(bg_file_t), self._dp(bg_off_t),
                                                   self._dp(bg_vol_t), self._dp(fg_vol_t), self._dp(out_t), B,
                                                   int(clamp), self._stream()))
        else:
            self._check(self.lib.kws_augment(self.h, self._dp(wav_t), self._dp(shift_t), self._dp(bg_file_t),
                                             self._dp(bg_off_t), self._dp(bg_vol_t), self._dp(fg_vol_t),
                                             self._dp(out_t), B, int(clamp), self._stream()))
        return out_t

    # -- stage 1b --
    def frontend_config(self, window_size_samples=480, window_stride_samples=160, n_mel=40, n_keep=None,
                        lower_edge_hertz=80.0, upper_edge_hertz=7600.0, sample_rate=16000):
        n_keep = n_mel if n_keep is None else n_keep
        self._check(self.lib.kws_frontend_config(self.h, window_size_samples, window_stride_samples, n_mel,
                                                 n_keep, lower_edge_hertz, upper_edge_hertz, sample_rate))
        self.fe = dict(frames=int(self.lib.kws_frontend_frames(self.h)),
                       bins=(1 << (window_size_samples - 1).bit_length()) // 2 + 1, n_mel=n_mel, n_keep=n_keep)

    def frontend_config_contrib(self, window_size=480, stride=160, sample_rate=16000, lower_frequency_limit=20.0,
                                upper_frequency_limit=4000.0, filterbank_channel_count=40, dct_coefficient_count=40):
        """contrib_audio.audio_spectrogram(magnitude_squared=True) + contrib_audio.mfcc flavour (audio.py:15-23)."""
        self._check(self.lib.kws_frontend_config_contrib(self.h, window_size, stride, sample_rate,
                                                         lower_frequency_limit, upper_frequency_limit,
                                                         filterbank_channel_count, dct_coefficient_count))
        self.fe = dict(frames=int(self.lib.kws_frontend_frames(self.h)),
                       bins=(1 << (window_size - 1).bit_length()) // 2 + 1, n_mel=filterbank_channel_count,
                       n_keep=dct_coefficient_count)

    def feature_shape(self, kind):
        k = _KIND.get(kind, kind)
        d = {FEAT_SPEC: self.fe["bins"], FEAT_LOGMEL: self.fe["n_mel"], FEAT_MFCC: self.fe["n_keep"]}[k]
        return self.fe["frames"], d

    def features(self, wav_t, kind="mfcc", out_t=None):
        import torch
        k = _KIND.get(kind, kind)
        B = wav_t.shape[0]
        fr, d = self.feature_shape(k)
        if out_t is None:
            out_t = torch.empty((B, fr, d), dtype=torch.float32, device=wav_t.device)
        self._check(self.lib.kws_features(self.h, self._dp(wav_t), B, k, self._dp(out_t), self._stream()))
        return out_t

    # -- stage 2 --
    def load_model(self, slot: int, arch: int, weights: dict):
        names = list(weights.keys())
        arrs = [np.ascontiguousarray(weights[n], np.float32) for n in names]
        ts = (_lib.TensorH * len(names))()
        for i, (n, a) in enumerate(zip(names, arrs)):
            ts[i].name = n.encode()
            ts[i].data = a.ctypes.data_as(C.POINTER(C.c_float))
            ts[i].numel = a.size
        self._check(self.lib.kws_model_load(self.h, slot, int(arch), ts, len(names)))
        return int(self.lib.kws_model_classes(self.h, slot))

    def classes(self, slot=0):
        return int(self.lib.kws_model_classes(self.h, slot))

    def forward(self, wav_t, views=((0, 1.0),), slot=0, want_probs=True, want_argmax=True):
        import torch
        B = wav_t.shape[0]
        Cn = self.classes(slot)
        sh, ga, n = _views(views)
        probs = torch.empty((B, Cn), dtype=torch.float32, device=wav_t.device) if want_probs else None
        amax = torch.empty((B,), dtype=torch.int32, device=wav_t.device) if want_argmax else None
        self._check(self.lib.kws_forward(self.h, slot, self._dp(wav_t), B, sh, ga, n, self._dp(probs),
                                         self._dp(amax), self._stream()))
        return probs, amax

    def debug_activation(self, wav_t, layer: int, shape, views=((0, 1.0),), slot=0):
        """fp32 activation after `layer` (0 = conv1d_1, i = block i); shape = (T, C) of that layer."""
        import torch
        B = wav_t.shape[0]
        sh, ga, n = _views(views)
        out = torch.empty((B * n, shape[0], shape[1]), dtype=torch.float32, device=wav_t.device)
        self._check(self.lib.kws_debug_activation(self.h, slot, self._dp(wav_t), B, sh, ga, n, int(layer),
                                                  self._dp(out), self._stream()))
        return out

    # -- driver math --
    def convert_classes(self, probs_t, class_map, n_out=12):
        import torch
        B, Cin = probs_t.shape
        cm = (C.c_int32 * Cin)(*[int(c) for c in class_map])
        out = torch.empty((B, n_out), dtype=torch.float32, device=probs_t.device)
        u8 = torch.empty((B, n_out), dtype=torch.uint8, device=probs_t.device)
        self._check(self.lib.kws_convert_classes(self.h, self._dp(probs_t), B, Cin, cm, n_out, self._dp(out),
                                                 self._dp(u8), self._stream()))
        return out, u8

    def select(self, probs_u8_t, thresh: float):
        import torch
        B, Cn = probs_u8_t.shape
        label = torch.empty((B,), dtype=torch.int32, device=probs_u8_t.device)
        keep = torch.empty((B,), dtype=torch.uint8, device=probs_u8_t.device)
        self._check(self.lib.kws_select(self.h, self._dp(probs_u8_t), B, Cn, float(thresh), self._dp(label),
                                        self._dp(keep), self._stream()))
        return label, keep

    def vote(self, labels_t, min_count=3):
        import torch
        M, B = labels_t.shape
        voted = torch.empty((B,), dtype=torch.int32, device=labels_t.device)
        clear = torch.empty((B,), dtype=torch.uint8, device=labels_t.device)
        self._check(self.lib.kws_vote(self.h, self._dp(labels_t), M, B, int(min_count), self._dp(voted),
                                      self._dp(clear), self._stream()))
        return voted, clear

    # -- host-buffer entry points (NumPy in / NumPy out) --
    def predict_host(self, wav: np.ndarray, views=((0, 1.0),), slot=0, probs_out=None, argmax_out=None):
        wav = np.ascontiguousarray(wav, np.float32)
        B = wav.shape[0]
        Cn = self.classes(slot)
        sh, ga, n = _views(views)
        probs = probs_out if probs_out is not None else np.empty((B, Cn), np.float32)
        amax = argmax_out if argmax_out is not None else np.empty((B,), np.int32)
        self._check(self.lib.kws_predict_host(self.h, slot, _hp(wav), B, sh, ga, n, _hp(probs), _hp(amax)))
        return probs, amax

    def get_data_host(self, wav, shift, bg_file, bg_off, bg_vol, fg_vol, kind="raw", out=None):
        wav = np.ascontiguousarray(wav, np.float32)
        B = wav.shape[0]
        k = _KIND.get(kind, kind)
        dim = SAMPLES if k == FEAT_RAW else int(np.prod(self.feature_shape(k)))
        if out is None:
            out = np.empty((B, dim), np.float32)
        args = [np.ascontiguousarray(shift, np.int32), np.ascontiguousarray(bg_file, np.int32),
                np.ascontiguousarray(bg_off, np.int32), np.ascontiguousarray(bg_vol, np.float32),
                np.ascontiguousarray(fg_vol, np.float32)]
        self._check(self.lib.kws_get_data_host(self.h, _hp(wav), *[_hp(a) for a in args], B, 0, k, _hp(out)))
        return out

    def pipeline_host(self, wav, params, feat_kind="logmel", views=((0, 1.0),), slot=0,
                      feat_out=None, probs_out=None, argmax_out=None):
        """augment -> features -> TTA forward on host buffers (the north-star path)."""
        B = wav.shape[0]
        k = _KIND.get(feat_kind, feat_kind)
        sh, ga, n = _views(views)
        p = params
        self._check(self.lib.kws_pipeline_host(
            self.h, slot, _hp(wav), _hp(p["time_shift"]), _hp(p["bg_index"]), _hp(p["bg_offset"]),
            _hp(p["bg_volume"]), _hp(p["fg_volume"]), B, k, sh, ga, n, _hp(feat_out), _hp(probs_out),
            _hp(argmax_out)))
        return feat_out, probs_out, argmax_out
