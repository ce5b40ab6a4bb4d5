"""Weight containers of the drop-in surface (make_submission.py:64 load_model on a Keras 2.1.2
.hdf5; make_submission_on_rpi.py:38-61 on a frozen .pb): the pure-Python HDF5 and GraphDef
readers against files written by independent writers (tests/h5_writer.py from the HDF5 spec;
tensorboard's protobuf classes for the GraphDef)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import h5_writer  # noqa: E402
from speech_recognition_b200 import hdf5_reader, synth, weights as W  # noqa: E402


def test_hdf5_roundtrip_types_and_groups(tmp_path):
    rng = np.random.RandomState(0)
    a = rng.randn(3, 40, 128).astype(np.float32)
    b = rng.randn(17).astype(np.float64)
    c = np.arange(24, dtype=np.int32).reshape(2, 3, 4)
    many = {f"member_{i:02d}": np.full((2,), i, np.float32) for i in range(41)}   # > 1 SNOD, names out of order
    tree = {"__attrs__": {"title": np.bytes_(b"kws"), "names": np.array([b"alpha", b"be", b"gamma_long"], dtype="S"),
                          "scale": np.float64(2.5), "count": np.int64(7)},
            "a": (a, {"unit": np.bytes_(b"none")}), "grp": {"b": b, "deep": {"c": c}}, "many": many,
            "empty": np.zeros((0, 4), np.float32), "scalar": np.asarray(3.25, np.float32)}
    path = str(tmp_path / "t.h5")
    h5_writer.write_h5(path, tree)
    with hdf5_reader.File(path) as f:
        assert sorted(f.keys()) == ["a", "empty", "grp", "many", "scalar"]
        assert f.attrs["title"] == b"kws" and f.attrs["scale"] == 2.5 and f.attrs["count"] == 7
        assert [x.decode() for x in f.attrs["names"]] == ["alpha", "be", "gamma_long"]
        assert np.array_equal(f["a"].read(), a) and f["a"].shape == (3, 40, 128) and f["a"].dtype == np.float32
        assert f["a"].attrs["unit"] == b"none"
        assert np.array_equal(f["grp/b"].read(), b)
        assert np.array_equal(f["grp"]["deep"]["c"].read(), c)
        assert sorted(f["many"].keys()) == sorted(many)
        for k, v in many.items():
            assert np.array_equal(f["many"][k].read(), v)
        assert f["empty"].read().shape == (0, 4)
        assert f["scalar"].read() == np.float32(3.25)
        assert "nope" not in f
        with pytest.raises(KeyError):
            f["grp/missing"]


@pytest.mark.parametrize("deflate", [False, True])
def test_hdf5_chunked(tmp_path, deflate):
    rng = np.random.RandomState(1)
    a = rng.randn(37, 10).astype(np.float32)
    path = str(tmp_path / "c.h5")
    h5_writer.write_h5(path, {"x": (a, {}, (8, 4), deflate)})
    with hdf5_reader.File(path) as f:
        assert np.array_equal(f["x"].read(), a)


def test_hdf5_rejects_other_files(tmp_path):
    p = tmp_path / "not.h5"
    p.write_bytes(b"PK\x03\x04" + b"\x00" * 100)
    with pytest.raises(hdf5_reader.HDF5Error):
        hdf5_reader.File(str(p))


def test_hdf5_reader_on_a_file_written_by_libhdf5():
    """The one file in this image that a REAL libhdf5 wrote: scipy ships MATLAB's v7.3 test file, which is an HDF5 file
    behind a 512-byte user block (superblock search, symbol-table group, v1 object header with attribute, contiguous
    float64 dataset -- the same structures h5py 2.x / libhdf5 1.8 emit for a Keras checkpoint).  scipy's own test suite
    documents its content: testdouble = 0 : pi/4 : 2*pi.  (h5py is not installed and the reference's checkpoints are
    not in the mount, so this is the independent pin there is; tests/h5_writer.py covers the Keras layout itself.)"""
    import scipy.io
    path = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(path):
        pytest.skip("scipy's MATLAB v7.3 test file is not installed")
    with hdf5_reader.File(path) as f:
        assert list(f.keys()) == ["testdouble"]
        d = f["testdouble"]
        assert d.shape == (9, 1) and d.dtype == np.float64 and d.attrs["MATLAB_class"] == b"double"
        np.testing.assert_allclose(d.read()[:, 0], np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)


@pytest.mark.parametrize("arch", [195, 106])
def test_keras_checkpoint_to_weights(tmp_path, arch):
    w = synth.synthetic_weights(arch)
    path = str(tmp_path / f"ep-001-vl-0.1.hdf5")
    h5_writer.write_h5(path, h5_writer.keras_checkpoint_tree(w, model_config='{"class_name": "Model"}'))
    got_arch, got = W.load_weights(path)
    assert got_arch in ((195, 206) if arch == 195 else (106,))
    assert set(got) == set(w)
    for k in w:
        assert np.array_equal(got[k], np.asarray(w[k], np.float32).reshape(got[k].shape)), k
    # weights-only file (model.save_weights) and per-process layer numbering (conv1d_13, ...)
    shifted = {}
    for k, v in w.items():
        layer, var = k.split("/")
        base, idx = layer.rsplit("_", 1)
        shifted[f"{base}_{int(idx) + 24}/{var}"] = v
    path2 = str(tmp_path / "weights_only.h5")
    h5_writer.write_h5(path2, h5_writer.keras_checkpoint_tree(shifted, wrap_model_weights=False))
    _, got2 = W.load_weights(path2)
    for k in w:
        assert np.array_equal(got2[k], got[k]), k


def _frozen_graph_bytes(w):
    """A frozen-graph-shaped GraphDef built with tensorboard's protobuf classes (the layout
    graph_util.convert_variables_to_constants produces: one Const per variable, plus op nodes)."""
    from tensorboard.compat.proto import graph_pb2, types_pb2
    g = graph_pb2.GraphDef()
    n = g.node.add(); n.name = "wav_fn"; n.op = "Placeholder"
    n = g.node.add(); n.name = "decoded_sample_data"; n.op = "DecodeWav"; n.input.append("ReadFile")
    n.attr["desired_samples"].i = 16000
    for k, v in w.items():
        v = np.asarray(v, np.float32)
        n = g.node.add(); n.name = k; n.op = "Const"
        n.attr["dtype"].type = types_pb2.DT_FLOAT
        t = n.attr["value"].tensor
        t.dtype = types_pb2.DT_FLOAT
        for d in v.shape:
            t.tensor_shape.dim.add().size = d
        if v.size <= 9:
            t.float_val.extend(v.ravel().tolist())            # small tensors are stored as repeated values
        else:
            t.tensor_content = v.tobytes()
        r = g.node.add(); r.name = k + "/read"; r.op = "Identity"; r.input.append(k)
    n = g.node.add(); n.name = "model_1/batch_normalization_1/cond/batchnorm/add/y"; n.op = "Const"
    n.attr["value"].tensor.dtype = types_pb2.DT_FLOAT
    n.attr["value"].tensor.float_val.append(1e-3)
    n = g.node.add(); n.name = "stft/frame_length"; n.op = "Const"
    n.attr["value"].tensor.dtype = types_pb2.DT_INT32
    n.attr["value"].tensor.int_val.append(480)
    n = g.node.add(); n.name = "labels_softmax"; n.op = "Reshape"
    return g.SerializeToString()


def test_frozen_pb_to_weights(tmp_path):
    from speech_recognition_b200 import pb_reader
    w = synth.synthetic_weights(195)
    path = str(tmp_path / "frozen_195.pb")
    with open(path, "wb") as f:
        f.write(_frozen_graph_bytes(w))
    consts = pb_reader.read_graph_constants(path)
    assert consts["__ops__"]["labels_softmax"] == "Reshape" and consts["stft/frame_length"] == 480
    assert np.isclose(consts["model_1/batch_normalization_1/cond/batchnorm/add/y"], 1e-3)
    arch, got = W.load_weights(path)
    assert arch == 195 and set(got) == set(w)
    for k in w:
        assert np.array_equal(got[k], np.asarray(w[k], np.float32).reshape(got[k].shape)), k
    with open(path, "wb") as f:
        f.write(b"\x0a\x03abc")
    with pytest.raises(pb_reader.PBError):
        pb_reader.read_frozen_graph_weights(path)


@pytest.mark.skipif(not os.path.isdir("/root/reference/logs_195"), reason="reference mount not present")
@pytest.mark.parametrize("log", ["logs_106", "logs_195", "logs_206"])
def test_pb_reader_on_reference_graphdefs(tmp_path, log):
    """The wire-format reader against tensorboard's own protobuf parser on the reference's REAL
    serialized graphs (train.py:64 TensorBoard callback): every numeric Const and every op name."""
    import glob
    from tensorboard.backend.event_processing.event_file_loader import RawEventFileLoader
    from tensorboard.compat.proto import event_pb2, graph_pb2
    from tensorboard.util import tensor_util
    from speech_recognition_b200 import pb_reader
    f = sorted(glob.glob(f"/root/reference/{log}/events.out.tfevents.*"))[0]
    gd = None
    for raw in RawEventFileLoader(f).Load():
        ev = event_pb2.Event.FromString(raw)
        if ev.graph_def:
            gd = ev.graph_def
            break
    path = str(tmp_path / "g.pb")
    with open(path, "wb") as fh:
        fh.write(gd)
    mine = pb_reader.read_graph_constants(path)
    g = graph_pb2.GraphDef.FromString(gd)
    checked = 0
    for node in g.node:
        assert mine["__ops__"][node.name] == node.op
        if node.op != "Const":
            continue
        try:
            ref = tensor_util.make_ndarray(node.attr["value"].tensor)
        except Exception:
            continue
        if ref.dtype.kind not in "fi":
            continue
        assert mine[node.name].shape == ref.shape and np.array_equal(mine[node.name], ref), node.name
        checked += 1
    assert checked > 1000
