"""CPU tests of the oracle (checker) against independent implementations and of the
host-side logic that mirrors the reference's parameter draw."""
import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import augment, frontend, network, driver
from speech_recognition_b200 import arch as parch, synth
from speech_recognition_b200.audio_processor import draw_augmentation_params
from speech_recognition_b200.model_settings import prepare_model_settings


@given(st.integers(-40000, 40000))
@settings(max_examples=200, deadline=None)
def test_tf_roll_is_np_roll(shift):
    # input_data.py:345 "TODO(see--): Write test with np.roll"
    x = np.arange(16000, dtype=np.float32)
    assert np.array_equal(augment.tf_roll(x, shift), np.roll(x, shift))


def test_augment_mix_separate_roundings():
    rs = np.random.RandomState(0)
    wav = rs.randn(4, 16000).astype(np.float32)
    bg = rs.randn(4, 16000).astype(np.float32)
    fv = np.array([1.0, 0.9, 1.13, 0.0], np.float32)
    bv = np.array([0.0, 0.1, 0.07, 0.3], np.float32)
    sh = np.array([0, -500, 77, 16000], np.int32)
    out = augment.augment_mix(wav, sh, bg, bv, fv)
    for b in range(4):
        ref = np.float32(bg[b] * bv[b]) + np.roll(np.float32(wav[b] * fv[b]), sh[b])
        assert np.array_equal(out[b], ref.astype(np.float32))
    assert out.dtype == np.float32


def test_same_pad_rules():
    # (1,1) for T=397,197,97,47 and (0,1) for T=22; patches pad (10,10)
    for T in (397, 197, 97, 47):
        assert network.same_pad(T, 3, 2)[1:] == (1, 1)
    assert network.same_pad(22, 3, 2) == (11, 0, 1)
    assert network.same_pad(16000, 40, 20) == (800, 10, 10)
    assert network.layer_lengths(195) == [800, 399, 397, 199, 197, 99, 97, 49, 47, 24, 22, 11, 9]
    assert parch.layer_lengths(106) == network.layer_lengths(106)
    assert parch.weight_shapes(195) == network.weight_shapes(195)
    assert sum(int(np.prod(s)) for s in network.weight_shapes(195).values()) == 1198601
    assert sum(int(np.prod(s)) for s in network.weight_shapes(106).values()) == 1092416


def test_spectrogram_matches_torch_fft():
    x = synth.make_clips(3, seed=11)
    ours = frontend.spectrogram(x)
    w = torch.tensor(frontend.hann_window_periodic(480))
    fr = torch.tensor(x).unfold(1, 480, 160) * w
    ref = torch.fft.rfft(torch.nn.functional.pad(fr, (0, 32)).double(), dim=-1).abs().float().numpy()
    assert ours.shape == (3, 98, 257)
    np.testing.assert_allclose(ours, ref, rtol=1e-6, atol=1e-6)
    # hann: torch's periodic window is the same function
    np.testing.assert_allclose(frontend.hann_window_periodic(480), torch.hann_window(480, periodic=True).numpy(),
                               atol=5e-7)


def test_mel_matrix_properties():
    for M in (40, 80):
        W = frontend.linear_to_mel_weight_matrix(M)
        assert W.shape == (257, M) and W.dtype == np.float32
        assert (W[0] == 0).all()                      # DC row is the zero pad
        assert ((W > 0).sum(axis=1) <= 2).all()       # triangular filters overlap pairwise
    assert (frontend.linear_to_mel_weight_matrix(40) > 0).sum() == 465
    assert (frontend.linear_to_mel_weight_matrix(80) > 0).sum() == 473


def test_mfcc_is_scaled_dct2():
    import scipy.fft
    rs = np.random.RandomState(1)
    lm = rs.randn(2, 98, 40).astype(np.float32)
    ours = frontend.mfcc_from_log_mel(lm)
    ref = scipy.fft.dct(lm.astype(np.float64), type=2, axis=-1) / np.sqrt(2 * 40)
    np.testing.assert_allclose(ours, ref, rtol=1e-5, atol=1e-5)


def test_network_matches_naive_loops():
    """The torch-op restatement against literal loops for the first layers and the head."""
    w = synth.synthetic_weights(195)
    x = synth.make_clips(2, seed=3)
    probs, logits, acts = network.forward(x, w, 195, dtype=torch.float64, return_activations=True)
    # conv1d_1 at a few positions: y[j,co] = sum_{f,i} P[2j+f,i] W[f,i,co], P[j,i] = x[20j-10+i]
    k = w["conv1d_1/kernel"].astype(np.float64)
    xp = np.pad(x[0].astype(np.float64), (10, 10))
    s = w["batch_normalization_1/gamma"] / np.sqrt(w["batch_normalization_1/moving_variance"].astype(np.float64) + 1e-3)
    sh = w["batch_normalization_1/beta"] - w["batch_normalization_1/moving_mean"] * s
    for j in (0, 1, 200, 398):
        acc = np.zeros(128)
        for f in range(3):
            patch = xp[20 * (2 * j + f): 20 * (2 * j + f) + 40]
            acc += patch @ k[f]
        ref = np.clip(acc * s + sh, 0, 6)
        np.testing.assert_allclose(acts[0][0, j], ref, rtol=1e-9, atol=1e-9)
    # block 2 (stride 2, SAME pad (1,1)) at the edges
    a_in = acts[1][0]                                   # [397,128]
    dk = w["depthwise_conv2d_2/depthwise_kernel"][0, :, :, 0].astype(np.float64)
    pk = w["conv1d_3/kernel"][0].astype(np.float64)
    s = w["batch_normalization_3/gamma"] / np.sqrt(w["batch_normalization_3/moving_variance"].astype(np.float64) + 1e-3)
    sh = w["batch_normalization_3/beta"] - w["batch_normalization_3/moving_mean"] * s
    for t in (0, 1, 198):
        d = np.zeros(128)
        for j in range(3):
            ti = 2 * t + j - 1
            if 0 <= ti < 397:
                d += dk[j] * a_in[ti]
        ref = np.clip((d @ pk) * s + sh, 0, 6)
        np.testing.assert_allclose(acts[2][0, t], ref, rtol=1e-9, atol=1e-9)
    # head
    xl = acts[-1][0]                                    # [9,512]
    att = xl.reshape(-1) @ w["dense_1/kernel"].astype(np.float64) + w["dense_1/bias"]
    att = np.exp(att - att.max()); att /= att.sum()
    z = np.concatenate([(xl * att[:, None]).max(0), xl.mean(0)])
    lg = z @ w["dense_2/kernel"].astype(np.float64)
    np.testing.assert_allclose(logits[0], lg, rtol=1e-9, atol=1e-9)
    assert probs.shape == (2, 12)


def test_network_106_shapes():
    w = synth.synthetic_weights(106)
    p = network.forward(synth.make_clips(2, seed=4), w, 106)
    assert p.shape == (2, 32)
    np.testing.assert_allclose(p.sum(1), 1.0, atol=1e-5)


def test_draw_order_matches_product_host_code():
    """oracle.draw_params (checker) and AudioProcessor's draw (product host logic) are two
    restatements of input_data.py:457-514: same seed -> same parameters, bit for bit."""
    ms = prepare_model_settings(12, 16000, 1000, 30.0, 10.0, 40, 40, 'raw')
    n = 200
    labels = np.arange(n) % 12                      # label 0 = silence
    pseudo_labels = (np.arange(50) * 7) % 12
    clips = np.zeros((n, 1), np.float32)
    bgs = [np.zeros(40000, np.float32), np.zeros(70000, np.float32), np.zeros(16500, np.float32)]
    kw = dict(background_frequency=0.3, background_volume_range=0.15, foreground_frequency=0.3,
              foreground_volume_range=0.15, time_shift_frequency=0.3, time_shift_range=[-500, 0],
              pseudo_frequency=0.33, flip_frequency=0.2, silence_volume_range=0.3)
    for mode in ("training", "validation"):
        rs = np.random.RandomState(123)
        ref = augment.draw_params(rs, how_many=64, offset=8, n_candidates=n, candidate_is_silence=labels == 0,
                                  mode=mode, background_lengths=[len(b) for b in bgs], n_pseudo=50,
                                  pseudo_is_silence=pseudo_labels == 0, **kw)
        np.random.seed(123)
        idx, from_pseudo, lab, p = draw_augmentation_params(
            {mode: (clips, labels), 'pseudo': (np.zeros((50, 1), np.float32), pseudo_labels)}, bgs, ms,
            64, 8, kw['background_frequency'], kw['background_volume_range'], kw['foreground_frequency'],
            kw['foreground_volume_range'], kw['time_shift_frequency'], kw['time_shift_range'], mode,
            kw['pseudo_frequency'], kw['flip_frequency'], kw['silence_volume_range'])
        assert np.array_equal(idx, ref['sample_index'])
        assert np.array_equal(from_pseudo, ref['from_pseudo'])
        for k_ref, k_p in (('time_shift', 'time_shift'), ('bg_index', 'bg_index'), ('bg_offset', 'bg_offset'),
                           ('bg_volume', 'bg_volume'), ('fg_volume', 'fg_volume')):
            assert np.array_equal(ref[k_ref], p[k_p]), (mode, k_ref)
        if mode == "training":
            assert (ref['time_shift'] <= 0).all() and (ref['time_shift'] >= -500).all()
            assert (p['fg_volume'][lab == 0] == 0).all()


def test_model_settings():
    ms = prepare_model_settings(12, 16000, 1000, 30.0, 10.0, 40, 40, 'mfcc')
    assert ms['spectrogram_length'] == 98 and ms['fingerprint_size'] == 3920     # settings.py:1-11
    ms = prepare_model_settings(32, 16000, 1000, 25.0, 15.0, 80, 60, 'raw')      # make_submission.py:53-57
    assert ms['window_size_samples'] == 400 and ms['window_stride_samples'] == 240
    assert ms['fingerprint_size'] == 16000


def test_tta_predict_order():
    calls = []

    def predict(x):
        calls.append(x.copy())
        return np.tile(np.arange(3, dtype=np.float32)[None] * (len(calls)), (len(x), 1))
    x = np.random.RandomState(0).randn(2, 16000).astype(np.float32)
    probs, pred = driver.tta_predict(predict, x, driver.TTA_SHIPPED)
    assert np.array_equal(calls[0], x)
    assert np.array_equal(calls[1], (np.float32(1.2) * x).astype(np.float32))     # loud
    assert np.array_equal(calls[2], np.roll(x, -1500, axis=1))                    # left
    assert np.array_equal(pred, [2, 2])


def test_contrib_audio_restatement_properties():
    """Native contrib_audio flavour (audio.py:15-23), parity unpinned: internal consistency checks of the
    restatement -- the filterbank splits every in-band bin between two adjacent channels (weights sum to
    one), the DCT equals scipy's DCT-II up to the sqrt(2/N)/2 normalisation of mfcc_dct.cc, and the power
    spectrogram equals the squared tf.signal-style magnitude."""
    from scipy.fft import dct
    from oracle import frontend as fe
    from speech_recognition_b200 import synth
    band, w, start, end = fe.contrib_mel_filterbank(257, 16000.0, 40, 20.0, 4000.0)
    assert (start, end) == (2, 128)
    assert (band[:start] == -2).all() and (band[end + 1:] == -2).all()
    assert (np.diff(band[start:end + 1]) >= 0).all() and band[start:end + 1].max() == 39
    assert ((w[start:end + 1] >= 0) & (w[start:end + 1] <= 1)).all()
    x = synth.make_clips(2, seed=5)
    p = fe.contrib_audio_spectrogram(x)
    m = fe.contrib_audio_spectrogram(x, magnitude_squared=False)
    assert np.allclose(p, m.astype(np.float64) ** 2, rtol=1e-5, atol=1e-12)
    # same frames / window as the tf.signal chain (window in double vs fp32: 1e-6 apart)
    assert np.allclose(m, fe.spectrogram(x), rtol=1e-4, atol=1e-5)
    lm = fe.contrib_mfcc(p, return_log_mel=True).astype(np.float64)
    mf = fe.contrib_mfcc(p)
    assert np.abs(dct(lm, type=2, axis=-1) * np.sqrt(2.0 / 40) / 2 - mf).max() < 1e-4
    # a unit-magnitude flat spectrum puts (sum of weights) into each channel
    flat = np.ones((1, 257), np.float32)
    e = np.exp(fe.contrib_mfcc(flat, return_log_mel=True).astype(np.float64))[0]
    dense = np.zeros((257, 40))
    for i in range(start, end + 1):
        if band[i] >= 0:
            dense[i, band[i]] += w[i]
        if band[i] + 1 < 40:
            dense[i, band[i] + 1] += 1 - w[i]
    assert np.allclose(e, dense.sum(0), rtol=1e-6)


def test_time_sliced_oracle_against_independent_torch_module():
    """conv_1d_time_sliced_model (model.py:716-772) restated twice: oracle.network.forward(arch=716) vs a plain
    torch.nn stack built here from the reference's builder, on the same Keras-named weights."""
    import torch
    import torch.nn.functional as F
    from oracle import network
    from speech_recognition_b200 import synth
    w = synth.synthetic_weights(716)
    x = torch.from_numpy(synth.make_clips(3, seed=4)).double()
    # overlapping_time_slice_stack(x, 40, 20): SAME patches, then Conv1D(32, 3, strides=2) over the patch axis
    p = F.pad(x, (10, 10)).unfold(1, 40, 20)                                   # [B,800,40]
    k = torch.as_tensor(w["conv1d_1/kernel"]).double()                          # [3,40,32]
    y = torch.stack([sum(p[:, 2 * j + f] @ k[f] for f in range(3)) for j in range(399)], 1)   # [B,399,32]

    def bn_relu6(y, i):
        g, b, m, v = (torch.as_tensor(w[f"batch_normalization_{i}/{n}"]).double() for n in ("gamma", "beta", "moving_mean", "moving_variance"))
        return torch.clamp((y - m) / torch.sqrt(v + 1e-3) * g + b, 0, 6)
    y = bn_relu6(y, 1)
    for i, (co, s) in enumerate(network.ARCHS[716]["blocks"], start=1):
        d = torch.as_tensor(w[f"depthwise_conv2d_{i}/depthwise_kernel"]).double()[0, :, :, 0]   # [3,C]
        T = y.shape[1]
        if s == 2:
            out = -(-T // 2); total = max((out - 1) * 2 + 3 - T, 0)
            y = F.pad(y, (0, 0, total // 2, total - total // 2))
        else:
            out = T - 2
        y = sum(y[:, j:j + s * (out - 1) + 1:s] * d[j] for j in range(3))
        y = bn_relu6(y @ torch.as_tensor(w[f"conv1d_{i + 1}/kernel"]).double()[0], i + 1)
    assert y.shape == (3, 3, 512)
    z = torch.clamp(y.mean(1) @ torch.as_tensor(w["dense_1/kernel"]).double(), 0, 6)
    ref = torch.softmax(z @ torch.as_tensor(w["dense_2/kernel"]).double(), -1).numpy()
    got = network.forward(x.numpy(), w, 716, dtype=torch.float64)
    np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-12)


def test_steffenet_oracle_structure():
    """steffeNet restatement (model.py:1663-1726): layer-creation-order names (the shortcut Conv1D / BN of a stride-2 residual
    block precede the block's own layers), lengths 320 -> 5, 20.1 M parameters, a probability vector; the product-side
    shape table (arch.py) is the same table."""
    import torch
    from oracle import network
    from speech_recognition_b200 import synth, arch
    shapes = network.steffenet_weight_shapes(12)
    assert list(shapes.items()) == list(arch.steffenet_weight_shapes(12).items())
    assert shapes["conv1d_1/kernel"] == (75, 1, 256) and shapes["conv1d_3/kernel"] == (1, 256, 320)      # first shortcut
    assert shapes["depthwise_conv2d_2/depthwise_kernel"] == (1, 3, 256, 1) and shapes["conv1d_4/kernel"] == (1, 256, 320)
    assert shapes["conv1d_32/kernel"] == (1, 1536, 1536) and shapes["dense_1/kernel"] == (3072, 12) and len(shapes) == 186
    assert sum(int(np.prod(v)) for v in shapes.values()) == 20102912
    w = synth.synthetic_weights(1663)
    x = synth.make_clips(2, seed=8)
    p = network.forward_steffenet(x, w)
    assert p.shape == (2, 12) and np.allclose(p.sum(1), 1.0) and np.isfinite(p).all()
    # a roll by a multiple of the stem's stride (50) shifts the stem's output by whole frames: the pooled head is
    # invariant up to the SAME-padding edges, so the label survives
    q = network.forward_steffenet(np.roll(x, 100, axis=1), w)
    assert np.array_equal(p.argmax(1), q.argmax(1))
