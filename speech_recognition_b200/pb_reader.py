"""Minimal reader for frozen TensorFlow GraphDef files (``tf_files/frozen_195.pb``) -- no
TensorFlow / protobuf dependency.

``freeze_graph.py:56-84`` of the reference loads the Keras checkpoint, calls the model on the
``decoded_sample_data`` tensor and runs ``convert_variables_to_constants``: every Keras variable
(``conv1d_1/kernel``, ``batch_normalization_3/moving_mean``, ...) becomes a ``Const`` node of
that name whose ``value`` attribute holds the tensor.  This module walks the protobuf wire
format directly:

    GraphDef.node (1) -> NodeDef{name (1), op (2), attr (5) map<string, AttrValue>}
    AttrValue.tensor (8) -> TensorProto{dtype (1), tensor_shape (2), tensor_content (4),
                                        float_val (5), double_val (6), int_val (7), int64_val (10)}
    TensorShapeProto.dim (2) -> Dim{size (1)}

``read_frozen_graph_weights(path)`` returns the float Const tensors whose names look like Keras
variables, renumbered per layer type from 1 (see hdf5_reader.canonical_names).
"""
from __future__ import annotations

import struct

import numpy as np

_DT = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 19: np.float16}
_VARS = ("kernel", "depthwise_kernel", "bias", "gamma", "beta", "moving_mean", "moving_variance")


class PBError(ValueError):
    pass


def _varint(buf, p):
    x = shift = 0
    while True:
        if p >= len(buf):
            raise PBError("truncated varint")
        b = buf[p]
        p += 1
        x |= (b & 0x7F) << shift
        if not b & 0x80:
            return x, p
        shift += 7
        if shift > 70:
            raise PBError("varint too long")


def _fields(buf):
    """Yield (field number, wire type, value) for one message; value is int for varint / fixed
    and a memoryview for length-delimited fields."""
    p, n = 0, len(buf)
    while p < n:
        key, p = _varint(buf, p)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, p = _varint(buf, p)
        elif wt == 1:
            v = bytes(buf[p:p + 8]); p += 8
        elif wt == 2:
            ln, p = _varint(buf, p)
            if p + ln > n:
                raise PBError("truncated length-delimited field")
            v = buf[p:p + ln]; p += ln
        elif wt == 5:
            v = bytes(buf[p:p + 4]); p += 4
        else:
            raise PBError(f"unsupported wire type {wt}")
        yield fno, wt, v


def _signed64(v):
    return v - (1 << 64) if v >= 1 << 63 else v


def _parse_shape(buf):
    dims = []
    for fno, wt, v in _fields(buf):
        if fno == 2 and wt == 2:
            size = 0
            for f2, w2, v2 in _fields(v):
                if f2 == 1 and w2 == 0:
                    size = _signed64(v2)
            dims.append(size)
    return tuple(dims)


def _parse_tensor(buf):
    dtype, shape, content = 0, (), None
    vals = {5: [], 6: [], 7: [], 10: []}
    for fno, wt, v in _fields(buf):
        if fno == 1 and wt == 0:
            dtype = v
        elif fno == 2 and wt == 2:
            shape = _parse_shape(v)
        elif fno == 4 and wt == 2:
            content = bytes(v)
        elif fno == 5:                                   # float_val: packed or repeated fixed32
            vals[5] += list(struct.unpack(f"<{len(v) // 4}f", bytes(v))) if wt == 2 else [struct.unpack("<f", v)[0]]
        elif fno == 6:
            vals[6] += list(struct.unpack(f"<{len(v) // 8}d", bytes(v))) if wt == 2 else [struct.unpack("<d", v)[0]]
        elif fno in (7, 10):
            if wt == 2:
                p = 0
                while p < len(v):
                    x, p = _varint(v, p)
                    vals[fno].append(_signed64(x))
            else:
                vals[fno].append(_signed64(v))
    if dtype not in _DT:
        return None
    np_dt = _DT[dtype]
    n = int(np.prod(shape)) if shape else 1
    if content is not None and len(content):
        return np.frombuffer(content, dtype=np.dtype(np_dt).newbyteorder("<"), count=n).reshape(shape).astype(np_dt)
    src = vals[{1: 5, 2: 6, 3: 7, 9: 10}.get(dtype, 5)]
    if not src:
        return np.zeros(shape, np_dt)
    arr = np.asarray(src, dtype=np_dt)
    if arr.size == 1 and n != 1:                         # TF stores a splat as one repeated value
        arr = np.full(n, arr[0], np_dt)
    return arr.reshape(shape)


def read_graph_constants(path: str) -> dict:
    """{node name: ndarray} for every Const node with a numeric tensor, plus '__ops__' -> {name: op}."""
    with open(path, "rb") as f:
        buf = memoryview(f.read())
    consts, ops = {}, {}
    seen_node = False
    for fno, wt, v in _fields(buf):
        if fno != 1 or wt != 2:
            continue
        name = op = None
        tensor = None
        for f2, w2, v2 in _fields(v):
            if f2 == 1 and w2 == 2:
                name = bytes(v2).decode("utf-8")
            elif f2 == 2 and w2 == 2:
                op = bytes(v2).decode("utf-8")
            elif f2 == 5 and w2 == 2:
                key = val = None
                for f3, w3, v3 in _fields(v2):
                    if f3 == 1 and w3 == 2:
                        key = bytes(v3)
                    elif f3 == 2 and w3 == 2:
                        val = v3
                if key == b"value" and val is not None:
                    for f4, w4, v4 in _fields(val):
                        if f4 == 8 and w4 == 2:
                            tensor = v4
        if name is None or op is None:
            continue
        seen_node = True
        ops[name] = op
        if op == "Const" and tensor is not None:
            arr = _parse_tensor(tensor)
            if arr is not None:
                consts[name] = arr
    if not seen_node:
        raise PBError(f"{path}: no GraphDef nodes found")
    consts["__ops__"] = ops
    return consts


def read_frozen_graph_weights(path: str) -> dict:
    from .hdf5_reader import canonical_names
    consts = read_graph_constants(path)
    out = {}
    for name, arr in consts.items():
        if name == "__ops__":
            continue
        parts = name.split("/")
        if len(parts) >= 2 and parts[-1] in _VARS and arr.dtype in (np.float32, np.float64, np.float16):
            out["/".join(parts[-2:])] = np.asarray(arr, np.float32)
    if not out:
        raise PBError(f"{path}: no Keras variables among the graph's constants (not a frozen graph?)")
    return canonical_names(out)
