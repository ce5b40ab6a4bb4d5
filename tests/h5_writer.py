"""Minimal HDF5 *writer* used only by the tests: emits the classic on-disk layout that
h5py 2.x / libhdf5 1.8 produce for a Keras 2.1.2 checkpoint (superblock v0, symbol-table groups
with a v1 B-tree + local heap + SNOD nodes, version-1 object headers with continuation-free
message lists, contiguous little-endian datasets, version-1 attribute messages holding
fixed-length string arrays / scalars).  Written from the HDF5 File Format Specification, not
from h5py, so it is an independent statement of the format the reader must understand.

    write_h5(path, {"model_weights": {"__attrs__": {...}, "conv1d_1": {"conv1d_1": {"kernel:0": ndarray}}}})
"""
from __future__ import annotations

import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 4, 16


def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


def _datatype(arr: np.ndarray) -> bytes:
    dt = arr.dtype
    if dt.kind == "f":
        size = dt.itemsize
        if size == 4:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            bits = bytes([0x20, 31, 0])
        elif size == 8:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            bits = bytes([0x20, 63, 0])
        else:
            raise NotImplementedError
        return bytes([0x11]) + bits + struct.pack("<I", size) + props
    if dt.kind in "iu":
        bits = bytes([0x08 if dt.kind == "i" else 0x00, 0, 0])
        return bytes([0x10]) + bits + struct.pack("<I", dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "S":
        return bytes([0x13]) + bytes([0x00, 0, 0]) + struct.pack("<I", dt.itemsize)     # null-terminated ASCII
    raise NotImplementedError(str(dt))


def _dataspace(shape) -> bytes:
    rank = len(shape)
    return bytes([1, rank, 0, 0]) + b"\x00" * 4 + b"".join(struct.pack("<Q", d) for d in shape)


def _message(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _attribute(name: str, value) -> bytes:
    arr = np.asarray(value)
    if arr.dtype.kind == "U":
        arr = np.char.encode(arr, "utf-8")
    nm = name.encode("utf-8") + b"\x00"
    dt, ds = _datatype(arr), _dataspace(arr.shape)
    body = struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + arr.tobytes()
    return _message(0x0C, body)


class _Writer:
    def __init__(self):
        self.buf = bytearray()

    def alloc(self, data: bytes) -> int:
        self.buf += b"\x00" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    def object_header(self, messages: list[bytes]) -> int:
        body = b"".join(messages)
        hdr = struct.pack("<BxHII4x", 1, len(messages), 1, len(body))
        return self.alloc(hdr + body)

    def dataset(self, arr: np.ndarray, attrs: dict) -> int:
        arr = np.ascontiguousarray(arr)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        data_addr = self.alloc(arr.tobytes()) if arr.size else UNDEF
        layout = struct.pack("<BBQQ", 3, 1, data_addr, arr.nbytes)
        msgs = [_message(0x01, _dataspace(arr.shape)), _message(0x03, _datatype(arr), flags=1),
                _message(0x05, bytes([2, 2, 2, 0])),                       # fill value v2: never written, undefined
                _message(0x08, layout)]
        msgs += [_attribute(k, v) for k, v in attrs.items()]
        return self.object_header(msgs)

    def dataset_chunked(self, arr: np.ndarray, chunks: tuple, deflate: bool, attrs: dict) -> int:
        """Chunked layout (v3 class 2) indexed by a single-node v1 B-tree, optional deflate filter."""
        import itertools
        import zlib
        arr = np.ascontiguousarray(arr)
        rank = arr.ndim
        entries = []
        grid = [range(0, s, c) for s, c in zip(arr.shape, chunks)]
        for offs in itertools.product(*grid):
            chunk = np.zeros(chunks, arr.dtype)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunks, arr.shape))
            chunk[tuple(slice(0, x.stop - x.start) for x in sl)] = arr[sl]
            raw = chunk.tobytes()
            if deflate:
                raw = zlib.compress(raw, 4)
            entries.append((offs, len(raw), self.alloc(raw)))
        assert len(entries) <= 64
        node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(entries), UNDEF, UNDEF)
        for offs, n, addr in entries:
            node += struct.pack("<II", n, 0) + b"".join(struct.pack("<Q", o) for o in offs) + struct.pack("<Q", 0)
            node += struct.pack("<Q", addr)
        node += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", s) for s in arr.shape) + struct.pack("<Q", 0)
        btree = self.alloc(node)
        layout = struct.pack("<BBBQ", 3, 2, rank + 1, btree) + b"".join(struct.pack("<I", c) for c in chunks)
        layout += struct.pack("<I", arr.dtype.itemsize)
        msgs = [_message(0x01, _dataspace(arr.shape)), _message(0x03, _datatype(arr), flags=1),
                _message(0x05, bytes([2, 2, 2, 0]))]
        if deflate:
            msgs.append(_message(0x0B, struct.pack("<BB6x", 1, 1) + struct.pack("<HHHH", 1, 0, 0, 1) + struct.pack("<II", 4, 0)))
        msgs.append(_message(0x08, layout))
        msgs += [_attribute(k, v) for k, v in attrs.items()]
        return self.object_header(msgs)

    def group(self, members: dict, attrs: dict) -> int:
        """members: name -> object header address."""
        names = sorted(members)
        heap_data = bytearray(b"\x00" * 8)                                  # offset 0 = empty string
        offs = {}
        for n in names:
            offs[n] = len(heap_data)
            heap_data += _pad8(n.encode("utf-8") + b"\x00")
        free_off = len(heap_data)
        heap_data += struct.pack("<QQ", 1, 16)                              # one free block: next = 1 (none), size 16
        heap_data_addr = self.alloc(bytes(heap_data))
        heap_addr = self.alloc(b"HEAP" + bytes([0, 0, 0, 0]) + struct.pack("<QQQ", len(heap_data), free_off, heap_data_addr))
        # symbol table nodes of at most 2 * LEAF_K entries each
        snods, keys = [], [0]
        for i in range(0, max(len(names), 1), 2 * LEAF_K):
            part = names[i:i + 2 * LEAF_K]
            ent = b"".join(struct.pack("<QQII16x", offs[n], members[n], 0, 0) for n in part)
            ent += b"\x00" * (40 * (2 * LEAF_K - len(part)))
            snods.append(self.alloc(b"SNOD" + struct.pack("<BxH", 1, len(part)) + ent))
            keys.append(offs[part[-1]] if part else 0)
        if len(snods) > 2 * INTERNAL_K:
            raise NotImplementedError("group too large for a single-level B-tree in this test writer")
        node = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF)
        for i, a in enumerate(snods):
            node += struct.pack("<QQ", keys[i], a)
        node += struct.pack("<Q", keys[len(snods)])
        node += b"\x00" * (16 * (2 * INTERNAL_K - len(snods)))
        btree_addr = self.alloc(node)
        msgs = [_message(0x11, struct.pack("<QQ", btree_addr, heap_addr))]
        msgs += [_attribute(k, v) for k, v in attrs.items()]
        return self.object_header(msgs), btree_addr, heap_addr

    def tree(self, node: dict):
        attrs = node.get("__attrs__", {})
        members = {}
        for k, v in node.items():
            if k == "__attrs__":
                continue
            if isinstance(v, dict):
                members[k] = self.tree(v)[0]
            elif isinstance(v, tuple) and len(v) == 4:                      # (array, attrs, chunks, deflate)
                members[k] = self.dataset_chunked(np.asarray(v[0]), v[2], v[3], v[1])
            elif isinstance(v, tuple):                                      # (array, attrs)
                members[k] = self.dataset(v[0], v[1])
            else:
                members[k] = self.dataset(np.asarray(v), {})
        return self.group(members, attrs)


def write_h5(path: str, tree: dict):
    w = _Writer()
    w.buf += b"\x00" * 96                                                   # superblock v0 (56 B) + root entry (40 B)
    root_addr, btree, heap = w.tree(tree)
    eof = len(w.buf)
    sb = b"\x89HDF\r\n\x1a\n" + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", LEAF_K, INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", btree, heap)   # cached symbol-table scratch
    assert len(sb) == 96
    w.buf[0:96] = sb
    with open(path, "wb") as f:
        f.write(bytes(w.buf))


def keras_checkpoint_tree(weights: dict, model_config: str = "{}", wrap_model_weights: bool = True) -> dict:
    """The group/attribute layout of keras.engine.topology.save_weights_to_hdf5_group (Keras 2.1.2)
    for weights named '<layer>/<var>' -> ndarray."""
    layers: dict = {}
    for k in weights:
        layers.setdefault(k.split("/")[0], []).append(k)
    g: dict = {"__attrs__": {"layer_names": np.array([n.encode() for n in layers], dtype="S"),
                             "backend": np.bytes_(b"tensorflow"), "keras_version": np.bytes_(b"2.1.2")}}
    for layer, keys in layers.items():
        wn = [f"{k}:0".encode() for k in keys]
        sub: dict = {}
        for k in keys:
            sub[k.split("/", 1)[1] + ":0"] = np.asarray(weights[k], np.float32)
        g[layer] = {"__attrs__": {"weight_names": np.array(wn, dtype="S")}, layer: sub}
    if not wrap_model_weights:
        return g
    return {"__attrs__": {"keras_version": np.bytes_(b"2.1.2"), "backend": np.bytes_(b"tensorflow"),
                          "model_config": np.bytes_(model_config.encode())},
            "model_weights": g,
            "optimizer_weights": {"__attrs__": {"weight_names": np.array([b"iterations:0"], dtype="S")},
                                  "iterations:0": np.asarray(12345, np.int64)}}
