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


def _in(a, dtype, shape, name):
    """Host INPUT array for the C ABI: C-contiguous, exact dtype (converted if needed), checked shape."""
    a = np.ascontiguousarray(a, dtype)
    if a.shape != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {a.shape}")
    return a


def _out(a, dtype, shape, name):
    """Host OUTPUT array: allocated if None; a caller-supplied one must already be writable, C-contiguous, of the
    exact dtype and hold exactly prod(shape) elements (the D2H copy would otherwise overrun or misinterpret it)."""
    if a is None:
        return np.empty(shape, dtype)
    if not isinstance(a, np.ndarray) or a.dtype != np.dtype(dtype) or not a.flags.c_contiguous or not a.flags.writeable:
        raise ValueError(f"{name}: need a writable C-contiguous {np.dtype(dtype).name} ndarray")
    if a.size != int(np.prod(shape)):
        raise ValueError(f"{name}: expected {int(np.prod(shape))} elements for shape {tuple(shape)}, got {a.size}")
    return a


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

    def _dev(self, t, dtype, shape, name):
        """Device tensor handed to the C ABI as a raw pointer: right device, dtype, shape and contiguous."""
        import torch
        if not isinstance(t, torch.Tensor) or not t.is_cuda or t.device.index != self.device:
            raise ValueError(f"{name}: need a torch tensor on cuda:{self.device}")
        if t.dtype != dtype or not t.is_contiguous():
            raise ValueError(f"{name}: need a contiguous {dtype} tensor, got {t.dtype} (contiguous={t.is_contiguous()})")
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t

    def set_host_staging(self, mode="auto"):
        """How the *_host entry points treat caller buffers: 'auto' (pinned -> DMA in place, pageable -> staged
        through the handle's pinned slots), 'always', 'never' (kws_set_host_staging)."""
        self._check(self.lib.kws_set_host_staging(self.h, {"auto": 0, "always": 1, "never": 2}[mode]))

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
        self._dev(wav_t, wav_t.dtype if wav_t.dtype == torch.int16 else torch.float32, (B, SAMPLES), "wav")
        self._dev(out_t, torch.float32, (B, SAMPLES), "out")
        for nm, t, dt in (("time_shift", shift_t, torch.int32), ("bg_index", bg_file_t, torch.int32),
                          ("bg_offset", bg_off_t, torch.int32), ("bg_volume", bg_vol_t, torch.float32),
                          ("fg_volume", fg_vol_t, torch.float32)):
            self._dev(t, dt, (B,), nm)
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

    # -- speed-TTA view --
    def time_stretch(self, pcm_t, rate=0.9, out_t=None):
        """create_tta_set.py:16-22 on the device: int16 PCM [B,16000] -> int16 PCM of the slowed clips."""
        import torch
        B = pcm_t.shape[0]
        self._dev(pcm_t, torch.int16, (B, SAMPLES), "pcm")
        if out_t is None:
            out_t = torch.empty((B, SAMPLES), dtype=torch.int16, device=pcm_t.device)
        self._dev(out_t, torch.int16, (B, SAMPLES), "out")
        self._check(self.lib.kws_time_stretch_pcm16(self.h, self._dp(pcm_t), B, float(rate), self._dp(out_t), self._stream()))
        return out_t

    def time_stretch_host(self, pcm, rate=0.9, out=None):
        pcm = _in(pcm, np.int16, (len(pcm), SAMPLES), "pcm")
        out = _out(out, np.int16, pcm.shape, "out")
        self._check(self.lib.kws_time_stretch_host_pcm16(self.h, _hp(pcm), pcm.shape[0], float(rate), _hp(out)))
        return out

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
        self._dev(wav_t, torch.float32, (B, SAMPLES), "wav")
        self._dev(out_t, torch.float32, None, "out")
        if out_t.numel() != B * fr * d:
            raise ValueError(f"out: expected {B * fr * d} elements, got {out_t.numel()}")
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
        self._dev(wav_t, torch.float32, (B, SAMPLES), "wav")
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
        self._dev(wav_t, torch.float32, (B, SAMPLES), "wav")
        sh, ga, n = _views(views)
        out = torch.empty((B * n, shape[0], shape[1]), dtype=torch.float32, device=wav_t.device)
        self._check(self.lib.kws_debug_activation(self.h, slot, self._dp(wav_t), B, sh, ga, n, int(layer),
                                                  self._dp(out), self._stream()))
        return out

    # -- driver math --
    def convert_classes(self, probs_t, class_map, n_out=12):
        import torch
        B, Cin = probs_t.shape
        self._dev(probs_t, torch.float32, None, "probs")
        cm = (C.c_int32 * Cin)(*[int(c) for c in class_map])
        out = torch.empty((B, n_out), dtype=torch.float32, device=probs_t.device)
        u8 = torch.empty((B, n_out), dtype=torch.uint8, device=probs_t.device)
        self._check(self.lib.kws_convert_classes(self.h, self._dp(probs_t), B, Cin, cm, n_out, self._dp(out),
                                                 self._dp(u8), self._stream()))
        return out, u8

    def select(self, probs_u8_t, thresh: float):
        import torch
        B, Cn = probs_u8_t.shape
        self._dev(probs_u8_t, torch.uint8, None, "probs_u8")
        label = torch.empty((B,), dtype=torch.int32, device=probs_u8_t.device)
        keep = torch.empty((B,), dtype=torch.uint8, device=probs_u8_t.device)
        self._check(self.lib.kws_select(self.h, self._dp(probs_u8_t), B, Cn, float(thresh), self._dp(label),
                                        self._dp(keep), self._stream()))
        return label, keep

    def vote(self, labels_t, min_count=3):
        import torch
        M, B = labels_t.shape
        self._dev(labels_t, torch.int32, None, "labels")
        voted = torch.empty((B,), dtype=torch.int32, device=labels_t.device)
        clear = torch.empty((B,), dtype=torch.uint8, device=labels_t.device)
        self._check(self.lib.kws_vote(self.h, self._dp(labels_t), M, B, int(min_count), self._dp(voted),
                                      self._dp(clear), self._stream()))
        return voted, clear

    # -- host-buffer entry points (NumPy in / NumPy out) --
    @staticmethod
    def _wav_in(wav):
        """[B,16000] float32, or int16 PCM (the WAV wire format) -- anything else is converted to float32."""
        wav = np.asarray(wav)
        dt = np.int16 if wav.dtype == np.int16 else np.float32
        if wav.ndim != 2 or wav.shape[1] != SAMPLES:
            raise ValueError(f"wav: expected shape (B, {SAMPLES}), got {wav.shape}")
        return np.ascontiguousarray(wav, dt)

    def predict_host(self, wav: np.ndarray, views=((0, 1.0),), slot=0, probs_out=None, argmax_out=None,
                     pcm_divisor=32768.0):
        wav = self._wav_in(wav)
        B = wav.shape[0]
        Cn = self.classes(slot)
        sh, ga, n = _views(views)
        probs = _out(probs_out, np.float32, (B, Cn), "probs_out")
        amax = _out(argmax_out, np.int32, (B,), "argmax_out")
        if wav.dtype == np.int16:
            self._check(self.lib.kws_predict_host_pcm16(self.h, slot, _hp(wav), float(pcm_divisor), B, sh, ga, n,
                                                        _hp(probs), _hp(amax)))
        else:
            self._check(self.lib.kws_predict_host(self.h, slot, _hp(wav), B, sh, ga, n, _hp(probs), _hp(amax)))
        return probs, amax

    def get_data_host(self, wav, shift, bg_file, bg_off, bg_vol, fg_vol, kind="raw", out=None):
        wav = np.ascontiguousarray(wav, np.float32)
        if wav.ndim != 2 or wav.shape[1] != SAMPLES:
            raise ValueError(f"wav: expected shape (B, {SAMPLES}), got {wav.shape}")
        B = wav.shape[0]
        k = _KIND.get(kind, kind)
        dim = SAMPLES if k == FEAT_RAW else int(np.prod(self.feature_shape(k)))
        out = _out(out, np.float32, (B, dim), "out")
        args = [_in(shift, np.int32, (B,), "time_shift"), _in(bg_file, np.int32, (B,), "bg_index"),
                _in(bg_off, np.int32, (B,), "bg_offset"), _in(bg_vol, np.float32, (B,), "bg_volume"),
                _in(fg_vol, np.float32, (B,), "fg_volume")]
        self._check(self.lib.kws_get_data_host(self.h, _hp(wav), *[_hp(a) for a in args], B, 0, k, _hp(out)))
        return out

    def features_host(self, wav, kind="mfcc", out=None):
        """STFT / log-mel / MFCC of host waveforms as they are (no augmentation) -> float32 [B, frames*dim]."""
        wav = np.ascontiguousarray(wav, np.float32)
        if wav.ndim != 2 or wav.shape[1] != SAMPLES:
            raise ValueError(f"wav: expected shape (B, {SAMPLES}), got {wav.shape}")
        B = wav.shape[0]
        k = _KIND.get(kind, kind)
        out = _out(out, np.float32, (B, int(np.prod(self.feature_shape(k)))), "out")
        self._check(self.lib.kws_get_data_host(self.h, _hp(wav), None, None, None, None, None, B, 0, k, _hp(out)))
        return out

    def pipeline_host(self, wav, params, feat_kind="logmel", views=((0, 1.0),), slot=0,
                      feat_out=None, probs_out=None, argmax_out=None, want_features=True, pcm_divisor=32768.0):
        """augment -> features -> TTA forward on host buffers (the north-star path).  ``wav`` is float32 or int16
        PCM [B,16000]; ``params`` holds the five pre-drawn parameter arrays (int64 / float64 inputs, NumPy's
        defaults, are converted).  ``want_features=False`` still computes the features on the device but leaves
        them there (the shipped networks consume the raw waveform, make_submission.py:46)."""
        wav = self._wav_in(wav)
        B = wav.shape[0]
        k = _KIND.get(feat_kind, feat_kind)
        sh, ga, n = _views(views)
        p = params
        pa = [_in(p["time_shift"], np.int32, (B,), "time_shift"), _in(p["bg_index"], np.int32, (B,), "bg_index"),
              _in(p["bg_offset"], np.int32, (B,), "bg_offset"), _in(p["bg_volume"], np.float32, (B,), "bg_volume"),
              _in(p["fg_volume"], np.float32, (B,), "fg_volume")]
        feat = None
        if want_features or feat_out is not None:
            dim = SAMPLES if k == FEAT_RAW else int(np.prod(self.feature_shape(k)))
            feat = _out(feat_out, np.float32, (B, dim), "feat_out")
        probs = amax = None
        if n > 0:
            probs = _out(probs_out, np.float32, (B, self.classes(slot)), "probs_out")
            amax = _out(argmax_out, np.int32, (B,), "argmax_out")
        if wav.dtype == np.int16:
            self._check(self.lib.kws_pipeline_host_pcm16(
                self.h, slot, _hp(wav), float(pcm_divisor), *[_hp(a) for a in pa], B, k, sh, ga, n, _hp(feat),
                _hp(probs), _hp(amax)))
        else:
            self._check(self.lib.kws_pipeline_host(
                self.h, slot, _hp(wav), *[_hp(a) for a in pa], B, k, sh, ga, n, _hp(feat), _hp(probs), _hp(amax)))
        return feat, probs, amax
