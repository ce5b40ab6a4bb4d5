"""Host-side logic that needs no GPU: the buffer checks Engine applies before a pointer crosses the C ABI, the trained
synthetic checkpoints and the keyword clip generator."""
import numpy as np
import pytest
import torch

from speech_recognition_b200 import engine as E, synth
from speech_recognition_b200.arch import weight_shapes, layer_lengths
from oracle import network


def test_host_buffer_checks():
    a = E._in(np.arange(6, dtype=np.int64), np.int32, (6,), "time_shift")          # NumPy's default int is converted
    assert a.dtype == np.int32 and a.flags.c_contiguous
    assert E._in(np.ones((4, 2), np.float64)[:, 0], np.float32, (4,), "v").flags.c_contiguous   # strided view -> contiguous copy
    with pytest.raises(ValueError, match="expected shape"):
        E._in(np.zeros(5), np.float32, (6,), "bg_volume")
    out = E._out(None, np.float32, (3, 12), "probs_out")
    assert out.shape == (3, 12) and out.dtype == np.float32
    ok = np.empty((3, 12), np.float32)
    assert E._out(ok, np.float32, (3, 12), "probs_out") is ok                      # used in place: the D2H copy lands in it
    for bad in (np.empty((3, 12), np.float64), np.empty((3, 11), np.float32), np.empty((12, 3), np.float32).T,
                [[0.0] * 12] * 3):
        with pytest.raises(ValueError):
            E._out(bad, np.float32, (3, 12), "probs_out")
    ro = np.empty((3, 12), np.float32); ro.setflags(write=False)
    with pytest.raises(ValueError):
        E._out(ro, np.float32, (3, 12), "probs_out")
    with pytest.raises(ValueError, match="expected shape"):
        E.Engine._wav_in(np.zeros((2, 15999), np.float32))
    assert E.Engine._wav_in(np.zeros((2, 16000), np.int16)).dtype == np.int16       # PCM stays PCM
    assert E.Engine._wav_in(np.zeros((2, 16000), np.float64)).dtype == np.float32


@pytest.mark.parametrize("arch", [195, 106, 206])
def test_trained_checkpoints(arch):
    """data/trained_<arch>.npz: every Keras tensor of the architecture, exactly representable in fp16, and a confident,
    accurate classifier of its own task through the float64 oracle (what the label-agreement tests rely on)."""
    w = synth.trained_weights(arch)
    shapes = weight_shapes(arch)
    assert list(w) == list(shapes) and all(w[k].shape == tuple(shapes[k]) and w[k].dtype == np.float32 for k in shapes)
    assert all(np.array_equal(v, v.astype(np.float16).astype(np.float32)) for v in w.values())
    C = network.ARCHS[arch]["classes"]
    x, y = synth.make_word_clips(96, C, seed=123 + arch, return_labels=True)
    p = network.forward(x.numpy(), w, arch, dtype=torch.float64)
    assert (p.argmax(1) == y.numpy()).mean() > 0.9 and np.median(p.max(1)) > 0.8


def test_word_clips_are_pcm_exact_and_seeded():
    a, la = synth.make_word_clips(16, 12, seed=5, return_labels=True)
    b, lb = synth.make_word_clips(16, 12, seed=5, return_labels=True)
    assert torch.equal(a, b) and torch.equal(la, lb) and a.shape == (16, 16000) and a.dtype == torch.float32
    pcm = a.numpy() * 32768.0
    assert np.array_equal(pcm, np.round(pcm)) and np.abs(pcm).max() <= 32768     # int16-quantised like decoded WAV data
    assert not torch.equal(a, synth.make_word_clips(16, 12, seed=6))
    assert layer_lengths(716)[-1] == 3 and layer_lengths(195)[-1] == 9
