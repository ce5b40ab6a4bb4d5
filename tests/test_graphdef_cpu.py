"""The oracle against the reference's OWN serialized TensorFlow graphs.

`tests/golden/graph_*.npz` hold the outputs of the GraphDefs stored in /root/reference/logs_{106,195,206}
(the exact graphs the reference trained and predicted with), evaluated node by node in NumPy by
`tests/golden/graphdef_eval.py` on synthetic inputs / weights.  These tests check that `oracle/` -- the
restatement every GPU parity test is judged against -- reproduces them, i.e. that its reading of the
paddings, strides, constants, tf_roll branches, BatchNorm form, head wiring and op order is the graph's.
"""
import os

import numpy as np
import pytest
import torch

from oracle import augment, frontend, network
from speech_recognition_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize("arch", [195, 206, 106])
def test_network_matches_reference_graph(arch):
    g = _load("graph_net_%d.npz" % arch)
    w = synth.synthetic_weights(195 if arch == 206 else arch)
    probs, _, acts = network.forward(g["x"], w, arch, dtype=torch.float64, return_activations=True)
    assert probs.shape == g["probs"].shape
    np.testing.assert_allclose(probs, g["probs"], rtol=1e-5, atol=1e-8)     # oracle returns fp32 probabilities
    assert np.array_equal(probs.argmax(-1), g["probs"].argmax(-1))
    for i in (0, 1, 2, 10, 11):                    # conv1d_1, first blocks, the (0,1)-padded block, the last
        ref = g["act_%d" % i]
        got = np.asarray(acts[i], np.float64).reshape(ref.shape)
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6)


def test_frontend_matches_reference_graph():
    g = _load("graph_frontend_195.npz")
    for i in range(3):
        shift, fv, bv = g["params_%d" % i]
        wav, bg = g["wav_%d" % i][None], g["bg_%d" % i][None]
        # stage 1a in float64 (the graph was evaluated in float64): same tf_roll branch, same mix
        mix = augment.tf_roll(wav[0].astype(np.float64), int(shift)) * fv + bg[0].astype(np.float64) * bv
        np.testing.assert_allclose(mix[None], g["background_clamp_%d" % i], rtol=0, atol=1e-15)
        x = g["background_clamp_%d" % i].astype(np.float32)
        spec = frontend.features(x, kind="spec")
        np.testing.assert_allclose(spec, g["spectrogram_%d" % i], rtol=2e-4, atol=2e-5)
        # exp 195 front end: 80 mel bins, keep 60 (train.py:38-39)
        lm = frontend.features(x, kind="logmel", dct_coefficient_count=80)
        mf = frontend.features(x, kind="mfcc", dct_coefficient_count=80, num_log_mel_features=60)
        assert lm.shape == g["logmel_%d" % i].shape and mf.shape == g["mfcc_%d" % i].shape
        np.testing.assert_allclose(lm, g["logmel_%d" % i], rtol=1e-4, atol=2e-3)
        np.testing.assert_allclose(mf, g["mfcc_%d" % i], rtol=1e-4, atol=5e-3)


def test_mel_matrix_and_window_match_reference_graph():
    g = _load("graph_frontend_195.npz")
    W = frontend.linear_to_mel_weight_matrix(80)
    np.testing.assert_allclose(W, g["mel_matrix_0"], rtol=1e-6, atol=1e-7)
    assert int((np.asarray(g["mel_matrix_0"]) != 0).sum()) == int((np.asarray(W) != 0).sum())
