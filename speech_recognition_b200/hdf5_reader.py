"""Minimal pure-Python HDF5 reader -- enough for Keras 2.1.2 checkpoints (h5py is not a dependency).

The reference loads ``checkpoints_*/ep-*.hdf5`` with ``keras.models.load_model``
(make_submission.py:64, freeze_graph.py:56).  Keras 2.1.2 writes those files through h5py 2.x /
libhdf5 1.8-1.10 with default settings, i.e. the "classic" on-disk format this module implements
from the HDF5 File Format Specification (version 2.0 / 3.0):

* superblock version 0/1 (and 2/3), 8-byte offsets and lengths
* groups as symbol tables (B-tree v1 + local heap + SNOD nodes) and as compact link messages
* object headers version 1 (with continuation blocks) and version 2 ("OHDR"/"OCHK")
* datasets: contiguous, compact and chunked (B-tree v1; deflate / shuffle / fletcher32 filters)
* datatypes: fixed-point, IEEE float, fixed-length strings, variable-length strings (global heap)
* attributes: message versions 1-3 stored compactly in the object header

Not implemented (a clear ``NotImplementedError`` is raised): dense attribute / link storage in
fractal heaps, layout message version 4, compound / array / reference datatypes.

``read_keras_weights(path)`` returns ``{"conv1d_1/kernel": ndarray, ...}`` following Keras'
``save_weights_to_hdf5_group`` layout: group ``model_weights`` (or the root for weights-only
files) with attribute ``layer_names``; one group per layer with attribute ``weight_names``
(e.g. ``b'conv1d_1/kernel:0'``) naming datasets below it.
"""
from __future__ import annotations

import mmap
import struct
import zlib

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class HDF5Error(ValueError):
    pass


def _pad8(n: int) -> int:
    return (n + 7) & ~7


class _Reader:
    def __init__(self, buf):
        self.buf = buf

    def u(self, off, size):
        return int.from_bytes(self.buf[off:off + size], "little")

    def bytes(self, off, size):
        return bytes(self.buf[off:off + size])


class File:
    """Read-only view of an HDF5 file; behaves like its root group."""

    def __init__(self, path):
        self._fh = open(path, "rb")
        try:
            self._mm = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
        except ValueError:
            self._fh.close()
            raise HDF5Error(f"{path}: empty file")
        self.r = _Reader(self._mm)
        self.path = path
        self._parse_superblock()
        self.root = Object(self, self.root_addr, "/")

    # -- container protocol delegates to the root group --
    def __getitem__(self, name):
        return self.root[name]

    def __contains__(self, name):
        return name in self.root

    def keys(self):
        return self.root.keys()

    @property
    def attrs(self):
        return self.root.attrs

    def close(self):
        try:
            self._mm.close()
        finally:
            self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _parse_superblock(self):
        r = self.r
        base = None
        for off in [0] + [512 << i for i in range(12)]:         # the superblock may follow a user block
            if off + 8 <= len(r.buf) and r.bytes(off, 8) == _SIG:
                base = off
                break
        if base is None:
            raise HDF5Error(f"{self.path}: not an HDF5 file (signature not found)")
        ver = r.u(base + 8, 1)
        if ver in (0, 1):
            self.O, self.L = r.u(base + 13, 1), r.u(base + 14, 1)
            p = base + 24 + (4 if ver == 1 else 0)
            self.base_addr = r.u(p, self.O)
            p += 4 * self.O                                       # base, free-space, EOF, driver info
            # root group symbol table entry: name offset, header address, cache type, reserved, scratch
            self.root_addr = r.u(p + self.O, self.O)
        elif ver in (2, 3):
            self.O, self.L = r.u(base + 9, 1), r.u(base + 10, 1)
            p = base + 12
            self.base_addr = r.u(p, self.O)
            self.root_addr = r.u(p + 3 * self.O, self.O)
        else:
            raise HDF5Error(f"unsupported superblock version {ver}")
        if self.O != 8 or self.L != 8:
            raise NotImplementedError("only 8-byte offsets / lengths are supported")
        self.root_addr += self.base_addr

    # -- shared structures --
    def local_heap_string(self, heap_addr, offset):
        r = self.r
        if r.bytes(heap_addr, 4) != b"HEAP":
            raise HDF5Error("bad local heap signature")
        data_addr = r.u(heap_addr + 8 + 2 * self.L, self.O) + self.base_addr
        start = data_addr + offset
        end = self._mm.find(b"\x00", start)
        return r.bytes(start, end - start).decode("utf-8")

    def global_heap_object(self, coll_addr, index):
        r = self.r
        coll_addr += self.base_addr
        if r.bytes(coll_addr, 4) != b"GCOL":
            raise HDF5Error("bad global heap signature")
        size = r.u(coll_addr + 8, self.L)
        p, end = coll_addr + 8 + self.L, coll_addr + size
        while p + 8 + self.L <= end:
            idx = r.u(p, 2)
            osize = r.u(p + 8, self.L)
            if idx == index:
                return r.bytes(p + 8 + self.L, osize)
            if idx == 0:
                break
            p += 8 + self.L + _pad8(osize)
        raise HDF5Error(f"global heap object {index} not found")


class Object:
    """A group or a dataset (decided by the header messages present)."""

    def __init__(self, f: File, addr: int, name: str):
        self.f, self.addr, self.name = f, addr, name
        self.msgs = self._read_header()
        self._attrs = None
        self._links = None

    # ---------------- object header ----------------
    def _read_header(self):
        r, f = self.f.r, self.f
        msgs = []
        a = self.addr
        if r.bytes(a, 4) == b"OHDR":
            if r.u(a + 4, 1) != 2:
                raise HDF5Error("bad object header version")
            flags = r.u(a + 5, 1)
            p = a + 6
            if flags & 0x20:
                p += 16
            if flags & 0x10:
                p += 4
            csize_len = 1 << (flags & 3)
            chunk_size = r.u(p, csize_len)
            p += csize_len
            blocks = [(p, p + chunk_size)]
            track_order = bool(flags & 0x04)
            while blocks:
                p, end = blocks.pop(0)
                while p + 4 <= end:
                    mtype, msize, _mflags = r.u(p, 1), r.u(p + 1, 2), r.u(p + 3, 1)
                    p += 4 + (2 if track_order else 0)
                    if p + msize > end:
                        break
                    if mtype == 0x10:
                        coff, clen = r.u(p, f.O) + f.base_addr, r.u(p + f.O, f.L)
                        if r.bytes(coff, 4) != b"OCHK":
                            raise HDF5Error("bad continuation chunk signature")
                        blocks.append((coff + 4, coff + clen - 4))
                    elif mtype != 0:
                        msgs.append((mtype, p, msize))
                    p += msize
            return msgs
        ver = r.u(a, 1)
        if ver != 1:
            raise HDF5Error(f"unsupported object header version {ver} at {a:#x}")
        nmsgs = r.u(a + 2, 2)
        hsize = r.u(a + 8, 4)
        blocks = [(a + 16, a + 16 + hsize)]
        while blocks and len(msgs) < nmsgs + 64:
            p, end = blocks.pop(0)
            while p + 8 <= end and nmsgs > 0:
                mtype, msize = r.u(p, 2), r.u(p + 2, 2)
                p += 8
                nmsgs -= 1
                if mtype == 0x10:
                    blocks.append((r.u(p, f.O) + f.base_addr, r.u(p, f.O) + f.base_addr + r.u(p + f.O, f.L)))
                elif mtype != 0:
                    msgs.append((mtype, p, msize))
                p += msize
        return msgs

    def _msg(self, mtype):
        for t, p, n in self.msgs:
            if t == mtype:
                return p, n
        return None

    @property
    def is_dataset(self):
        return self._msg(0x08) is not None and self._msg(0x01) is not None

    @property
    def is_group(self):
        return not self.is_dataset

    # ---------------- datatype / dataspace ----------------
    def _parse_datatype(self, p):
        """-> (kind, numpy dtype or None, size, consumed bytes)."""
        r = self.f.r
        cv = r.u(p, 1)
        cls, _ver = cv & 0x0F, cv >> 4
        bits = r.u(p + 1, 3)
        size = r.u(p + 4, 4)
        if cls == 0:
            order = ">" if bits & 1 else "<"
            kind = "i" if bits & 0x08 else "u"
            return "num", np.dtype(f"{order}{kind}{size}"), size, 8 + 4
        if cls == 1:
            order = ">" if bits & 1 else "<"
            if size not in (2, 4, 8):
                raise NotImplementedError(f"float of {size} bytes")
            return "num", np.dtype(f"{order}f{size}"), size, 8 + 12
        if cls == 3:
            return "str", np.dtype(f"S{size}"), size, 8
        if cls == 9:
            if (bits & 0x0F) != 1:
                raise NotImplementedError("variable-length sequences")
            _, _, _, used = self._parse_datatype(p + 8)
            return "vlen_str", None, size, 8 + used
        raise NotImplementedError(f"HDF5 datatype class {cls}")

    def _parse_dataspace(self, p):
        r, f = self.f.r, self.f
        ver, rank, flags = r.u(p, 1), r.u(p + 1, 1), r.u(p + 2, 1)
        if ver == 1:
            q = p + 8
        elif ver == 2:
            if r.u(p + 3, 1) == 2:
                return None                                       # null dataspace
            q = p + 4
        else:
            raise HDF5Error(f"dataspace version {ver}")
        return tuple(r.u(q + i * f.L, f.L) for i in range(rank))

    def _decode(self, kind, dt, size, shape, raw):
        n = int(np.prod(shape)) if shape else 1
        if kind in ("num", "str"):
            arr = np.frombuffer(raw, dtype=dt, count=n).reshape(shape)
            return arr.copy()
        out = []
        for i in range(n):                                        # variable-length string: length, heap address, index
            rec = raw[i * size:(i + 1) * size]
            ln = int.from_bytes(rec[0:4], "little")
            addr = int.from_bytes(rec[4:4 + self.f.O], "little")
            idx = int.from_bytes(rec[4 + self.f.O:8 + self.f.O], "little")
            out.append(self.f.global_heap_object(addr, idx)[:ln] if ln else b"")
        arr = np.array(out, dtype=object).reshape(shape)
        return arr

    # ---------------- attributes ----------------
    @property
    def attrs(self):
        if self._attrs is None:
            r = self.f.r
            out = {}
            if self._msg(0x15) is not None:
                p, _ = self._msg(0x15)
                flags = r.u(p + 1, 1)
                q = p + 2 + (2 if flags & 1 else 0)
                if r.u(q, self.f.O) != _UNDEF:
                    raise NotImplementedError("dense attribute storage (fractal heap) is not supported")
            for t, p, n in self.msgs:
                if t != 0x0C:
                    continue
                ver = r.u(p, 1)
                nsz, dsz, ssz = r.u(p + 2, 2), r.u(p + 4, 2), r.u(p + 6, 2)
                q = p + 8 + (1 if ver == 3 else 0)
                pad = _pad8 if ver == 1 else (lambda v: v)
                name = r.bytes(q, nsz).split(b"\x00")[0].decode("utf-8")
                q += pad(nsz)
                kind, dt, size, _ = self._parse_datatype(q)
                q += pad(dsz)
                shape = self._parse_dataspace(q)
                q += pad(ssz)
                if shape is None:
                    out[name] = None
                    continue
                cnt = int(np.prod(shape)) if shape else 1
                val = self._decode(kind, dt, size, shape, r.bytes(q, cnt * size))
                out[name] = val[()] if shape == () else val
            self._attrs = out
        return self._attrs

    # ---------------- group ----------------
    def _load_links(self):
        if self._links is not None:
            return self._links
        r, f = self.f.r, self.f
        links = {}
        st = self._msg(0x11)
        if st is not None:
            btree = r.u(st[0], f.O) + f.base_addr
            heap = r.u(st[0] + f.O, f.O) + f.base_addr
            self._walk_group_btree(btree, heap, links)
        li = self._msg(0x02)
        if li is not None:
            p = li[0]
            flags = r.u(p + 1, 1)
            q = p + 2 + (8 if flags & 1 else 0)
            if r.u(q, f.O) != _UNDEF:
                raise NotImplementedError("dense link storage (fractal heap) is not supported")
        for t, p, n in self.msgs:
            if t != 0x06:
                continue
            flags = r.u(p + 1, 1)
            q = p + 2
            ltype = 0
            if flags & 0x08:
                ltype = r.u(q, 1); q += 1
            if flags & 0x04:
                q += 8
            if flags & 0x10:
                q += 1
            lsz = 1 << (flags & 3)
            nlen = r.u(q, lsz); q += lsz
            name = r.bytes(q, nlen).decode("utf-8"); q += nlen
            if ltype == 0:
                links[name] = r.u(q, f.O) + f.base_addr
        self._links = links
        return links

    def _walk_group_btree(self, addr, heap, links):
        r, f = self.f.r, self.f
        sig = r.bytes(addr, 4)
        if sig == b"SNOD":
            n = r.u(addr + 6, 2)
            p = addr + 8
            esz = 2 * f.O + 8 + 16
            for i in range(n):
                e = p + i * esz
                name = f.local_heap_string(heap, r.u(e, f.O))
                links[name] = r.u(e + f.O, f.O) + f.base_addr
            return
        if sig != b"TREE":
            raise HDF5Error(f"bad group B-tree node at {addr:#x}")
        if r.u(addr + 4, 1) != 0:
            raise HDF5Error("expected a group B-tree node")
        n = r.u(addr + 6, 2)
        p = addr + 8 + 2 * f.O
        for i in range(n):
            child = r.u(p + f.L + i * (f.L + f.O), f.O) + f.base_addr
            self._walk_group_btree(child, heap, links)

    def keys(self):
        return list(self._load_links().keys())

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __getitem__(self, name):
        if name is Ellipsis or name == ():
            return self.read()
        obj = self
        for part in [s for s in name.split("/") if s]:
            links = obj._load_links()
            if part not in links:
                raise KeyError(f"{name!r}: no member {part!r} in {obj.name!r}")
            obj = Object(self.f, links[part], obj.name.rstrip("/") + "/" + part)
        return obj

    # ---------------- dataset ----------------
    @property
    def shape(self):
        return self._parse_dataspace(self._msg(0x01)[0])

    @property
    def dtype(self):
        return self._parse_datatype(self._msg(0x03)[0])[1]

    def read(self):
        if not self.is_dataset:
            raise TypeError(f"{self.name} is a group")
        r, f = self.f.r, self.f
        shape = self.shape
        kind, dt, size, _ = self._parse_datatype(self._msg(0x03)[0])
        if shape is None:
            return None
        n = int(np.prod(shape)) if shape else 1
        p, _ = self._msg(0x08)
        ver = r.u(p, 1)
        if ver == 3:
            cls = r.u(p + 1, 1)
            if cls == 0:
                raw = r.bytes(p + 4, r.u(p + 2, 2))
            elif cls == 1:
                addr, nbytes = r.u(p + 2, f.O), r.u(p + 2 + f.O, f.L)
                raw = b"\x00" * (n * size) if addr == _UNDEF else r.bytes(addr + f.base_addr, nbytes)
            elif cls == 2:
                rank = r.u(p + 2, 1)
                btree = r.u(p + 3, f.O)
                cdims = [r.u(p + 3 + f.O + 4 * i, 4) for i in range(rank)]
                return self._read_chunked(btree, cdims[:-1], shape, kind, dt, size)
            else:
                raise NotImplementedError(f"layout class {cls}")
        elif ver in (1, 2):
            rank, cls = r.u(p + 1, 1), r.u(p + 2, 1)
            q = p + 8
            addr = None
            if cls != 0:
                addr = r.u(q, f.O); q += f.O
            dims = [r.u(q + 4 * i, 4) for i in range(rank)]
            q += 4 * rank
            if cls == 2:
                return self._read_chunked(addr, dims, shape, kind, dt, size)
            if cls == 0:
                raw = r.bytes(q + 4, r.u(q, 4))
            else:
                raw = r.bytes(addr + f.base_addr, n * size)
        else:
            raise NotImplementedError(f"data layout message version {ver} (file written with libver='latest')")
        return self._decode(kind, dt, size, shape, raw[: n * size])

    def __array__(self, dtype=None, copy=None):
        a = self.read()
        return a.astype(dtype) if dtype is not None else a

    def _filters(self):
        m = self._msg(0x0B)
        if m is None:
            return []
        r = self.f.r
        p = m[0]
        ver, nf = r.u(p, 1), r.u(p + 1, 1)
        q = p + (8 if ver == 1 else 2)
        out = []
        for _ in range(nf):
            fid = r.u(q, 2); q += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                nlen = r.u(q, 2); q += 2
            q += 2                                                # flags
            ncd = r.u(q, 2); q += 2
            q += _pad8(nlen) if ver == 1 else nlen
            cd = [r.u(q + 4 * i, 4) for i in range(ncd)]
            q += 4 * ncd
            if ver == 1 and ncd % 2:
                q += 4
            out.append((fid, cd))
        return out

    def _read_chunked(self, btree, cdims, shape, kind, dt, size):
        if kind == "vlen_str":
            raise NotImplementedError("chunked variable-length strings")
        r, f = self.f.r, self.f
        filters = self._filters()
        out = np.zeros(shape, dtype=dt)
        rank = len(shape)
        if btree == _UNDEF:
            return out

        def walk(addr):
            addr += f.base_addr
            if r.bytes(addr, 4) != b"TREE" or r.u(addr + 4, 1) != 1:
                raise HDF5Error("bad chunk B-tree node")
            level, n = r.u(addr + 5, 1), r.u(addr + 6, 2)
            ksz = 8 + 8 * (rank + 1)
            p = addr + 8 + 2 * f.O
            for i in range(n):
                k = p + i * (ksz + f.O)
                nbytes, mask = r.u(k, 4), r.u(k + 4, 4)
                offs = [r.u(k + 8 + 8 * d, 8) for d in range(rank)]
                child = r.u(k + ksz, f.O)
                if level > 0:
                    walk(child)
                    continue
                raw = r.bytes(child + f.base_addr, nbytes)
                for j, (fid, cd) in reversed(list(enumerate(filters))):
                    if mask & (1 << j):
                        continue
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:
                        es = cd[0] if cd else size
                        a = np.frombuffer(raw, np.uint8)
                        m = len(a) // es
                        raw = a[: m * es].reshape(es, m).T.tobytes() + a[m * es:].tobytes()
                    elif fid == 3:
                        raw = raw[:-4]
                    else:
                        raise NotImplementedError(f"HDF5 filter {fid}")
                chunk = np.frombuffer(raw, dtype=dt, count=int(np.prod(cdims))).reshape(cdims)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
                out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
        walk(btree)
        return out


# ------------------------------------------------------------------------------------------------
# Keras 2.1.2 checkpoint layout
# ------------------------------------------------------------------------------------------------
def _as_str_list(v):
    if v is None:
        return []
    arr = np.atleast_1d(v)
    return [x.decode("utf-8") if isinstance(x, (bytes, np.bytes_)) else str(x) for x in arr.tolist()]


def read_keras_weights(path: str) -> dict:
    """{variable name without ':0' -> float32 ndarray} from a Keras HDF5 checkpoint
    (``model.save`` file with a ``model_weights`` group, or a ``save_weights`` file)."""
    out = {}
    with File(path) as f:
        g = f["model_weights"] if "model_weights" in f else f.root
        layer_names = _as_str_list(g.attrs.get("layer_names"))
        if not layer_names:
            raise HDF5Error(f"{path}: no 'layer_names' attribute -- not a Keras weight file")
        for ln in layer_names:
            lg = g[ln]
            for wn in _as_str_list(lg.attrs.get("weight_names")):
                arr = np.asarray(lg[wn].read(), dtype=np.float32)
                key = wn.split(":")[0]
                parts = key.split("/")
                out["/".join(parts[-2:])] = arr                   # drop wrapper-model scopes
    return canonical_names(out)


def canonical_names(weights: dict) -> dict:
    """Keras numbers layers per process (conv1d_13, ...); renumber every layer type from 1 in
    order of its numeric suffix so the names match the architecture tables (arch.py)."""
    import re
    by_type = {}
    for k in weights:
        layer = k.split("/")[0]
        m = re.match(r"^(.*?)(?:_(\d+))?$", layer)
        by_type.setdefault(m.group(1), set()).add(int(m.group(2) or 0))
    remap = {}
    for t, idx in by_type.items():
        for new, old in enumerate(sorted(idx), start=1):
            remap[(t, old)] = new
    out = {}
    for k, v in weights.items():
        layer, var = k.split("/", 1)
        m = re.match(r"^(.*?)(?:_(\d+))?$", layer)
        out[f"{m.group(1)}_{remap[(m.group(1), int(m.group(2) or 0))]}/{var}"] = v
    return out
