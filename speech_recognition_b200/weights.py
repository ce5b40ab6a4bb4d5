"""Weight containers for the raw-waveform networks.

``load_weights(path)`` returns ``(arch, {keras_variable_name: float32 array})``.
Formats: ``.npz`` (keys = Keras variable names, optional ``__arch__``), Keras 2.1.2
``.hdf5`` checkpoints and frozen ``.pb`` graphs (see hdf5_reader.py / pb_reader.py).
"""
from __future__ import annotations

import numpy as np

from .arch import weight_shapes


def infer_arch(weights: dict) -> int:
    c0 = weights["conv1d_1/kernel"].shape[-1]
    if weights["conv1d_1/kernel"].shape[0] == 75 and "dense_2/kernel" not in weights:
        return 1663                                  # steffeNet (model.py:1663-1726)
    classes = weights["dense_2/kernel"].shape[-1]
    if c0 == 128 and classes == 12:
        return 195
    if c0 == 64 and classes == 32:
        return 106
    if c0 == 32 and weights["dense_1/kernel"].shape[-1] == 256:
        return 716                                   # conv_1d_time_sliced_model (model.py:716-772)
    raise ValueError(f"unrecognised network: conv1d_1 has {c0} filters, dense_2 has {classes} classes")


def validate(arch: int, weights: dict) -> dict:
    out = {}
    for name, shp in weight_shapes(arch).items():
        if name not in weights:
            raise KeyError(f"missing variable {name!r} for architecture {arch}")
        a = np.ascontiguousarray(weights[name], np.float32)
        if tuple(a.shape) != tuple(shp):          # same element count in another layout would load silently wrong
            raise ValueError(f"{name}: shape {a.shape} does not match {tuple(shp)}")
        out[name] = a
    return out


def save_npz(path: str, arch: int, weights: dict):
    np.savez(path, __arch__=np.int32(arch), **weights)


def load_weights(path: str):
    if path.endswith(".npz"):
        with np.load(path) as z:
            w = {k: z[k] for k in z.files if k != "__arch__"}
            arch = int(z["__arch__"]) if "__arch__" in z.files else infer_arch(w)
        return arch, validate(arch, w)
    if path.endswith(".hdf5") or path.endswith(".h5"):
        from .hdf5_reader import read_keras_weights
        w = read_keras_weights(path)
        arch = infer_arch(w)
        return arch, validate(arch, w)
    if path.endswith(".pb"):
        from .pb_reader import read_frozen_graph_weights
        w = read_frozen_graph_weights(path)
        arch = infer_arch(w)
        return arch, validate(arch, w)
    raise ValueError(f"unknown weight container: {path}")
