"""Evaluate the reference's OWN serialized TensorFlow graphs in NumPy (no TensorFlow needed).

Keras' TensorBoard callback wrote the full `GraphDef` of each experiment into
`/root/reference/logs_*/events.out.tfevents.*` (train.py:64).  That graph holds both halves of the hot path
exactly as the reference built them: the `AudioProcessor` processing graph (input_data.py:311-381: tf_roll,
mix, stft, mel, log, mfcc) and the Keras network (model.py:775-838 for exp 195 / 206, the older exp-106
variant).  This module is a small lazy interpreter for the TF 1.4 ops on those forward paths.  Feeding
synthetic inputs / weights and fetching `dense_2/Softmax`, `Abs` (spectrogram_), `Log`, `strided_slice`
(mfcc_) gives known answers that come from the reference's graph structure, constants, paddings and
strides -- not from this repository's reading of the Python source.

`python tests/golden/make_golden.py graph` runs it in the build container (needs /root/reference and
tensorboard's protobuf classes) and writes `graph_*.npz`; `tests/test_graphdef_cpu.py` checks the oracle
against those files on any machine.

Arithmetic is float64 unless the graph casts (a structural pin; TF kernel rounding is not reproduced).
"""
from __future__ import annotations

import glob
import os
import sys

import numpy as np

REF = "/root/reference"
DEAD = object()          # value of the untaken output of a Switch


def load_graph(log_dir):
    from tensorboard.backend.event_processing.event_file_loader import RawEventFileLoader
    from tensorboard.compat.proto import event_pb2, graph_pb2
    f = sorted(glob.glob(os.path.join(REF, log_dir, "events.out.tfevents.*")))[0]
    for raw in RawEventFileLoader(f).Load():
        ev = event_pb2.Event.FromString(raw)
        if ev.graph_def:
            return graph_pb2.GraphDef.FromString(ev.graph_def)
    raise RuntimeError("no graph_def in " + f)


def _const(node):
    from tensorboard.util import tensor_util
    return tensor_util.make_ndarray(node.attr["value"].tensor)


def _same_pads(size, k, s):
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return out, total // 2, total - total // 2


def _conv2d(x, w, strides, padding, depthwise=False):
    """NHWC x, HWIO w (depthwise: HWC1 multiplier 1); cross-correlation like TF."""
    n, h, wd, c = x.shape
    kh, kw = w.shape[:2]
    sh, sw = strides[1], strides[2]
    if padding == "SAME":
        oh, pt, pb = _same_pads(h, kh, sh)
        ow, pl, pr = _same_pads(wd, kw, sw)
        x = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    else:
        oh, ow = (h - kh) // sh + 1, (wd - kw) // sw + 1
    out = None
    for i in range(kh):
        for j in range(kw):
            patch = x[:, i:i + (oh - 1) * sh + 1:sh, j:j + (ow - 1) * sw + 1:sw, :]
            term = patch * w[i, j, :, 0] if depthwise else patch @ w[i, j]
            out = term if out is None else out + term
    return out


def _extract_patches(x, ksizes, strides, rates, padding):
    n, h, wd, c = x.shape
    assert h == 1 and c == 1 and ksizes[1] == 1 and rates[2] == 1
    k, s = ksizes[2], strides[2]
    if padding == "SAME":
        ow, pl, pr = _same_pads(wd, k, s)
        x = np.pad(x, ((0, 0), (0, 0), (pl, pr), (0, 0)))
    else:
        ow = (wd - k) // s + 1
    idx = s * np.arange(ow)[:, None] + np.arange(k)[None, :]
    return x[:, 0, :, 0][:, idx][:, None, :, :]               # [n, 1, ow, k]


def _strided_slice(x, begin, end, strides, a):
    bm, em, sm, nm, el = (a["begin_mask"].i, a["end_mask"].i, a["shrink_axis_mask"].i, a["new_axis_mask"].i,
                          a["ellipsis_mask"].i)
    idx = []
    for d in range(len(begin)):
        if el & (1 << d):
            idx.append(Ellipsis)
            continue
        if nm & (1 << d):
            idx.append(np.newaxis)
            continue
        if sm & (1 << d):
            idx.append(int(begin[d]))
            continue
        b = None if bm & (1 << d) else int(begin[d])
        e = None if em & (1 << d) else int(end[d])
        idx.append(slice(b, e, int(strides[d])))
    return x[tuple(idx)]


class Evaluator:
    def __init__(self, graph, feeds, variables, float_dtype=np.float64):
        self.nodes = {n.name: n for n in graph.node}
        self.feeds = feeds
        self.vars = variables
        self.fd = float_dtype
        self.cache = {}
        self.ops_used = set()

    def get(self, ref):
        ref = ref.lstrip("^")
        name, _, port = ref.partition(":")
        port = int(port) if port else 0
        n = self.nodes.get(name)
        if n is not None and n.op == "Switch" and name not in self.feeds:
            # lazy: the data input of the untaken side is never evaluated (e.g. BatchNorm batch moments)
            data_in = [i for i in n.input if not i.startswith("^")]
            pred = bool(self.get(data_in[1]))
            return self.get(data_in[0]) if port == (1 if pred else 0) else DEAD
        out = self.node_outputs(name)
        return out[port] if isinstance(out, tuple) else out

    def node_outputs(self, name):
        if name in self.cache:
            return self.cache[name]
        if name in self.feeds:
            v = self.feeds[name]
        else:
            v = self._eval(self.nodes[name])
        self.cache[name] = v
        return v

    def _f(self, v):
        return np.asarray(v, self.fd) if np.issubdtype(np.asarray(v).dtype, np.floating) else np.asarray(v)

    def _eval(self, n):  # noqa: C901
        op = n.op
        self.ops_used.add(op)
        a = n.attr
        data_in = [i for i in n.input if not i.startswith("^")]
        if op == "Const":
            return self._f(_const(n))
        if op in ("VariableV2", "Variable"):
            if n.name not in self.vars:
                raise KeyError("no value for variable " + n.name)
            return self._f(self.vars[n.name])
        if op == "Placeholder":
            raise KeyError("placeholder %s needs a feed" % n.name)
        if op == "Switch":
            data, pred = self.get(data_in[0]), bool(self.get(data_in[1]))
            return (DEAD, data) if pred else (data, DEAD)
        if op == "Merge":
            vals = [self.get(i) for i in data_in]
            live = [v for v in vals if v is not DEAD]
            assert len(live) == 1, n.name
            return (live[0], np.int32(0))
        x = []
        for i in data_in:
            v = self.get(i)
            if v is DEAD:
                return DEAD
            x.append(v)
        if op in ("Identity", "StopGradient", "PlaceholderWithDefault"):
            return x[0]
        if op == "Mul":
            return x[0] * x[1]
        if op in ("Add", "BiasAdd"):
            return x[0] + x[1]
        if op == "Sub":
            return x[0] - x[1]
        if op == "RealDiv":
            return x[0] / x[1]
        if op == "FloorDiv":
            return np.floor_divide(x[0], x[1])
        if op == "FloorMod":
            return np.mod(x[0], x[1])
        if op == "Neg":
            return -x[0]
        if op == "Rsqrt":
            return 1.0 / np.sqrt(x[0])
        if op == "Sqrt":
            return np.sqrt(x[0])
        if op == "SquaredDifference":
            return (x[0] - x[1]) ** 2
        if op == "Square":
            return x[0] * x[0]
        if op == "Log":
            return np.log(x[0])
        if op == "Exp":
            return np.exp(x[0])
        if op == "Cos":
            return np.cos(x[0])
        if op == "Relu":
            return np.maximum(x[0], 0)
        if op == "Maximum":
            return np.maximum(x[0], x[1])
        if op == "Minimum":
            return np.minimum(x[0], x[1])
        if op in ("ComplexAbs", "Abs"):
            return np.abs(x[0])
        if op == "Real":
            return np.real(x[0])
        if op == "Complex":
            return np.asarray(x[0]) + 1j * np.asarray(x[1])
        if op == "Cast":
            dst = a["DstT"].type
            m = {1: self.fd, 2: np.float64, 3: np.int32, 9: np.int64, 10: np.bool_, 8: np.complex128, 18: np.complex128}
            if dst in (8, 18):
                return np.asarray(x[0]).astype(np.complex128)
            return np.asarray(x[0]).astype(m[dst])
        if op == "Shape":
            return np.array(np.shape(x[0]), np.int32)
        if op == "Size":
            return np.int32(np.size(x[0]))
        if op == "Rank":
            return np.int32(np.ndim(x[0]))
        if op == "Reshape":
            return np.reshape(x[0], [int(v) for v in x[1]])
        if op == "Squeeze":
            dims = tuple(a["squeeze_dims"].list.i)
            return np.squeeze(x[0], axis=dims if dims else None)
        if op == "ExpandDims":
            return np.expand_dims(x[0], int(x[1]))
        if op == "ConcatV2":
            return np.concatenate([np.atleast_1d(v) for v in x[:-1]], axis=int(x[-1]))
        if op == "Pack":
            return np.stack(x, axis=a["axis"].i)
        if op == "Tile":
            return np.tile(x[0], [int(v) for v in x[1]])
        if op == "Fill":
            return np.full([int(v) for v in x[0]], x[1])
        if op == "Range":
            return np.arange(x[0], x[1], x[2])
        if op == "LinSpace":
            return np.linspace(x[0], x[1], int(x[2]))
        if op == "Pad":
            return np.pad(x[0], [(int(p[0]), int(p[1])) for p in x[1]])
        if op == "Transpose":
            return np.transpose(x[0], [int(v) for v in x[1]])
        if op in ("Gather", "GatherV2"):
            axis = int(x[2]) if op == "GatherV2" else 0
            return np.take(x[0], np.asarray(x[1], np.int64), axis=axis)
        if op == "ListDiff":
            keep = [v for v in np.asarray(x[0]).tolist() if v not in set(np.asarray(x[1]).tolist())]
            return (np.array(keep, np.int32), np.arange(len(keep), dtype=np.int32))
        if op == "Prod":
            return np.prod(x[0], axis=tuple(np.atleast_1d(x[1]).tolist()), keepdims=a["keep_dims"].b).astype(np.int32)
        if op in ("Max", "Mean", "Sum"):
            f = {"Max": np.max, "Mean": np.mean, "Sum": np.sum}[op]
            return f(x[0], axis=tuple(np.atleast_1d(x[1]).tolist()), keepdims=a["keep_dims"].b)
        if op in ("GreaterEqual", "LessEqual", "Less", "Greater", "Equal"):
            f = {"GreaterEqual": np.greater_equal, "LessEqual": np.less_equal, "Less": np.less, "Greater": np.greater,
                 "Equal": np.equal}[op]
            return f(x[0], x[1])
        if op == "Select":
            return np.where(x[0], x[1], x[2])
        if op == "Slice":
            b, s = [int(v) for v in x[1]], [int(v) for v in x[2]]
            return x[0][tuple(slice(bi, None if si == -1 else bi + si) for bi, si in zip(b, s))]
        if op == "StridedSlice":
            return _strided_slice(x[0], x[1], x[2], x[3], a)
        if op == "SplitV":
            sizes = [int(v) for v in x[1]]
            cuts = np.cumsum(sizes)[:-1]
            return tuple(np.split(x[0], cuts, axis=int(x[2])))
        if op == "Split":                                          # inputs: (axis, value)
            return tuple(np.split(x[1], a["num_split"].i, axis=int(x[0])))
        if op == "Sin":
            return np.sin(x[0])
        if op == "Floor":
            return np.floor(x[0])
        if op == "Reciprocal":
            return 1.0 / x[0]
        if op == "ZerosLike":
            return np.zeros_like(x[0])
        if op == "OnesLike":
            return np.ones_like(x[0])
        if op == "MatMul":
            l = x[0].T if a["transpose_a"].b else x[0]
            r = x[1].T if a["transpose_b"].b else x[1]
            return l @ r
        if op == "Softmax":
            e = np.exp(x[0] - x[0].max(axis=-1, keepdims=True))
            return e / e.sum(axis=-1, keepdims=True)
        if op == "RFFT":
            return np.fft.rfft(x[0], n=int(np.atleast_1d(x[1])[0]), axis=-1)
        if op == "Conv2D":
            assert a["data_format"].s in (b"NHWC", b"")
            return _conv2d(x[0], x[1], list(a["strides"].list.i), a["padding"].s.decode())
        if op == "DepthwiseConv2dNative":
            assert x[1].shape[-1] == 1
            return _conv2d(x[0], x[1], list(a["strides"].list.i), a["padding"].s.decode(), depthwise=True)
        if op == "ExtractImagePatches":
            return _extract_patches(x[0], list(a["ksizes"].list.i), list(a["strides"].list.i), list(a["rates"].list.i),
                                    a["padding"].s.decode())
        raise NotImplementedError("op %s (node %s)" % (op, n.name))


# ----------------------------------------------------------------------------------------------------------
def keras_variables(graph, weights):
    """Map the graph's VariableV2 nodes (Keras names, e.g. 'conv1d_1/kernel') to our weight dict."""
    out = {}
    for n in graph.node:
        if n.op == "VariableV2" and n.name in weights:
            shp = [d.size for d in n.attr["shape"].shape.dim]
            out[n.name] = np.asarray(weights[n.name], np.float64).reshape(shp)
    return out


def run_network(log_dir, x, weights):
    """dense_2/Softmax (and the 12 block activations) of the reference graph for waveforms x [B,16000]."""
    g = load_graph(log_dir)
    names = {n.name for n in g.node}
    lp = [n.name for n in g.node if n.op == "Placeholder" and n.name.endswith("keras_learning_phase")][0]
    feeds = {"input_1": np.asarray(x, np.float64), lp: np.bool_(False)}
    ev = Evaluator(g, feeds, keras_variables(g, weights))
    acts = []
    for i in range(1, 13):
        acts.append(np.asarray(ev.get("activation_%d/clip_by_value" % i)))
    probs = np.asarray(ev.get("dense_2/Softmax"))
    assert "dense_2/Softmax" in names
    return probs, acts, sorted(ev.ops_used)


def run_frontend(log_dir, wav, fg_volume, time_shift, bg, bg_volume):
    """background_clamp_ ('Reshape'), spectrogram_ ('Abs'), log-mel ('Log') and mfcc_ ('strided_slice') of the
    reference's processing graph (input_data.py:332-381) for ONE clip -- the graph is built for batch 1.
    Placeholders: filename -> ReadFile -> DecodeWav (fed directly with the decoded samples), foreground_volme
    [sic], timeshift, background_data, background_volume."""
    g = load_graph(log_dir)
    feeds = {
        "DecodeWav": (np.asarray(wav, np.float64).reshape(16000, 1), np.int32(16000)),
        "foreground_volme": np.float64(fg_volume),
        "timeshift": np.int32(time_shift),
        "background_data": np.asarray(bg, np.float64).reshape(16000, 1),
        "background_volume": np.float64(bg_volume),
    }
    ev = Evaluator(g, feeds, {})
    out = {"background_clamp": np.asarray(ev.get("Reshape")),
           "spectrogram": np.asarray(ev.get("Abs")),
           "logmel": np.asarray(ev.get("Log")),
           "mfcc": np.asarray(ev.get("strided_slice")),
           "mel_matrix": np.asarray(ev.get("linear_to_mel_weight_matrix")),
           "hann": np.asarray(ev.get("stft/hann_window/sub_2")) if "stft/hann_window/sub_2" in ev.nodes else np.zeros(0)}
    return out, sorted(ev.ops_used)


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(os.path.dirname(here)))
    from speech_recognition_b200 import synth
    # ---- networks ----
    for arch, log_dir in ((195, "logs_195"), (206, "logs_206"), (106, "logs_106")):
        w = synth.synthetic_weights(195 if arch == 206 else arch)
        x = synth.make_clips(2, seed=900 + arch)
        probs, acts, ops = run_network(log_dir, x, w)
        keep = {"x": x.astype(np.float32), "probs": probs.astype(np.float64)}
        for i in (0, 1, 2, 10, 11):
            keep["act_%d" % i] = acts[i].astype(np.float32)
        np.savez_compressed(os.path.join(here, "graph_net_%d.npz" % arch), **keep)
        print("graph_net_%d" % arch, probs.shape, "ops:", ops)
    # ---- front end (logs_195 holds the HEAD processing graph: 80 mel, keep 60) ----
    rs = np.random.RandomState(7)
    clips = synth.make_clips(3, seed=901)
    bank, offs = synth.make_noise_bank(seconds=2)
    res = {}
    for i, (shift, fv, bv) in enumerate(((0, 1.0, 0.0), (-317, 0.9, 0.1), (211, -1.1, 0.05))):
        bg = bank[offs[i % (len(offs) - 1)] + 100: offs[i % (len(offs) - 1)] + 100 + 16000]
        out, ops = run_frontend("logs_195", clips[i], fv, shift, bg, bv)
        for k, v in out.items():
            res["%s_%d" % (k, i)] = v.astype(np.float64)
        res["wav_%d" % i] = clips[i]; res["bg_%d" % i] = bg.astype(np.float32)
        res["params_%d" % i] = np.array([shift, fv, bv], np.float64)
    np.savez_compressed(os.path.join(here, "graph_frontend_195.npz"), **res)
    print("graph_frontend_195", sorted(res)[:8], "ops:", ops)


if __name__ == "__main__":
    main()
