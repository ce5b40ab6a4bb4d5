"""The Python drop-in surface on the GPU, against the oracle: AudioProcessor.get_data (input_data.py:395-541),
load_model(path).predict (make_submission.py:64-71,120), the host entry points with int16 PCM / pageable /
pinned buffers, the submission writers and the pseudo-label helpers."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import h5_writer  # noqa: E402
from oracle import augment, frontend, network, driver, stretch  # noqa: E402
from speech_recognition_b200 import (AudioProcessor, Engine, load_model, prepare_model_settings, synth,  # noqa: E402
                                     submission, pseudo, TTA_SHIPPED, TTA_8)
from speech_recognition_b200.audio_processor import draw_augmentation_params  # noqa: E402
from speech_recognition_b200 import weights as W  # noqa: E402

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / (np.abs(b).max() + 1e-30))


@pytest.fixture(scope="module")
def dataset():
    clips = synth.make_clips(40, seed=901)
    labels = (np.arange(40) % 12).astype(np.int64)            # every third clip of a dozen is _silence_ (index 0)
    pseudo_clips = synth.make_clips(8, seed=902)
    bank, offs = synth.make_noise_bank(seconds=3)
    bg = [bank[offs[i]:offs[i + 1]] for i in range(len(offs) - 1)]
    return dict(clips=clips, labels=labels, pseudo=(pseudo_clips, (np.arange(8) % 12).astype(np.int64)), bg=bg,
                bank=bank, offs=offs)


@pytest.mark.parametrize("rep", ["raw", "spec", "mfcc", "mfcc_and_raw"])
def test_audio_processor_get_data(engine, dataset, rep):
    """Return contract of get_data: float64 containers, one-hot float64 labels, the reference's draw order from the
    GLOBAL np.random state (training mode: shift / background / volumes / flip / pseudo), values == oracle."""
    engine.set_precision("fp32")
    ms = prepare_model_settings(12, 16000, 1000, 30.0, 10.0, 80, 60, output_representation=rep)
    data = {"training": (dataset["clips"], dataset["labels"]), "validation": (dataset["clips"][:10], dataset["labels"][:10]),
            "pseudo": dataset["pseudo"]}
    ap = AudioProcessor(ms, output_representation=rep, data=data, background_data=dataset["bg"], engine=engine)
    assert ap.set_size("training") == 40 and ap.set_size("validation") == 10
    args = dict(how_many=24, offset=0, background_frequency=0.7, background_volume_range=0.15,
                foreground_frequency=0.5, foreground_volume_range=0.15, time_shift_frequency=0.6,
                time_shift_range=[-500, 0], mode="training", pseudo_frequency=0.25, flip_frequency=0.3,
                silence_volume_range=0.3)
    np.random.seed(1234)
    out, onehot = ap.get_data(sess=None, **args)
    # the same draws, replayed
    np.random.seed(1234)
    idx, from_pseudo, labels, p = draw_augmentation_params(ap.data_index, ap.background_data, ms, **args)
    assert len(idx) == 24 and from_pseudo.any() and (~from_pseudo).any()
    assert (p["time_shift"] != 0).any() and (p["fg_volume"] < 0).any() and (p["bg_volume"] > 0).any()
    src = np.where(from_pseudo[:, None], dataset["pseudo"][0][np.minimum(idx, 7)], dataset["clips"][np.minimum(idx, 39)])
    bgs = augment.gather_background(dataset["bank"], dataset["offs"], p["bg_index"], p["bg_offset"])
    r_raw = augment.augment_mix(src.astype(np.float32), p["time_shift"], bgs, p["bg_volume"], p["fg_volume"])
    assert onehot.dtype == np.float64 and onehot.shape == (24, 12)
    assert np.array_equal(onehot.argmax(1), labels) and np.array_equal(onehot.sum(1), np.ones(24))
    if rep == "raw":
        assert out.dtype == np.float64 and out.shape == (24, ms["fingerprint_size"]) == (24, 16000)
        assert np.array_equal(out, r_raw.astype(np.float64))
    elif rep == "spec":
        assert out.dtype == np.float64 and out.shape == (24, 98 * 257) == (24, ms["fingerprint_size"])
        assert rel_err(out.reshape(24, 98, 257), frontend.features(r_raw, kind="spec")) < 1e-5
    else:
        r_mf = frontend.features(r_raw, dct_coefficient_count=80, num_log_mel_features=60, kind="mfcc")
        mf = out[0] if rep == "mfcc_and_raw" else out
        assert mf.dtype == np.float64 and mf.shape == (24, 98 * 60) == (24, ms["fingerprint_size"])
        assert rel_err(mf.reshape(24, 98, 60), r_mf) < 1e-4
        if rep == "mfcc_and_raw":
            assert isinstance(out, list) and len(out) == 2 and np.array_equal(out[1], r_raw.astype(np.float64))
    # non-training modes are deterministic and un-augmented (utils.py:15-23); how_many = -1 returns the partition
    np.random.seed(5)
    v, vl = ap.get_data(-1, 0, 0.0, 0.0, 0.0, 0.0, 0.0, [0, 0], "validation", None)
    vv = v[1] if rep == "mfcc_and_raw" else v
    assert len(vl) == 10 and np.array_equal(vl.argmax(1), dataset["labels"][:10])
    if rep in ("raw", "mfcc_and_raw"):
        want = dataset["clips"][:10].astype(np.float64).copy()
        want[dataset["labels"][:10] == 0] = 0.0                      # silence label: foreground volume 0
        assert np.array_equal(vv, want)
    e, el = ap.get_data(5, 40, 0.0, 0.0, 0.0, 0.0, 0.0, [0, 0], "training", None)   # offset past the end
    assert (e[0] if rep == "mfcc_and_raw" else e).shape[0] == 0 and el.shape == (0, 12)
    engine.frontend_config(480, 160, 40, 40)


@pytest.mark.parametrize("container", ["npz", "hdf5", "pb"])
@pytest.mark.parametrize("arch", [195, 106])
def test_load_model_predict(tmp_path, engine, container, arch):
    """keras.models.load_model(path, custom_objects).predict(x) (make_submission.py:64-71,120) from each weight
    container, through the tensor-core tier on a TRAINED checkpoint: probabilities and labels vs the float64 oracle."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_weights_io_cpu import _frozen_graph_bytes
    w = synth.trained_weights(arch)
    if container == "npz":
        path = str(tmp_path / "w.npz"); W.save_npz(path, arch, w)
    elif container == "hdf5":
        path = str(tmp_path / "ep-085-vl-0.2231.hdf5")
        h5_writer.write_h5(path, h5_writer.keras_checkpoint_tree(w, model_config='{"class_name": "Model"}'))
    else:
        path = str(tmp_path / f"frozen_{arch}.pb")
        with open(path, "wb") as f:
            f.write(_frozen_graph_bytes(w))
    engine.set_precision("tc")
    try:
        model = load_model(path, custom_objects={"relu6": None, "DepthwiseConv2D": None}, engine=engine, slot=2)
        C = network.ARCHS[arch]["classes"]
        assert model.num_classes == C
        x = synth.make_word_clips(48, C, seed=77 + arch).numpy()
        probs = model.predict(x, batch_size=32, verbose=0)
        ref = network.forward(x, w, arch, dtype=torch.float64)
        assert probs.dtype == np.float32 and probs.shape == (48, C)
        assert np.abs(probs - ref).max() < 1e-2 and np.array_equal(probs.argmax(1), ref.argmax(1))
        p3, a3 = model.predict_tta(x)
        r3, ra3 = driver.tta_predict(lambda v: network.forward(v, w, arch, dtype=torch.float64), x, TTA_SHIPPED)
        assert np.abs(p3 - r3).max() < 1e-2 and np.array_equal(a3, ra3)
        with pytest.raises(ValueError, match="expected input_1 to have shape"):
            model.predict(x[:, :15999])
    finally:
        engine.set_precision("fp32")


def test_real_checkpoint_gated(engine):
    """Activates when the reference's own artefacts are supplied (they are missing from the mount):
    KWS_REAL_CKPT=<checkpoints_195/ep-085-vl-0.2231.hdf5 | tf_files/frozen_195.pb>.  Loads it through the pure-Python
    readers and checks the tensor-core tier against the oracle on the same weights (label for label)."""
    path = os.environ.get("KWS_REAL_CKPT", "")
    if not path or not os.path.exists(path):
        pytest.skip("set KWS_REAL_CKPT to a reference checkpoint (.hdf5) or frozen graph (.pb)")
    arch, w = W.load_weights(path)
    engine.set_precision("tc")
    try:
        model = load_model(path, engine=engine, slot=2)
        x = synth.make_clips(64, seed=5150)
        p, a = model.predict_tta(x)
        r, ra = driver.tta_predict(lambda v: network.forward(v, w, arch, dtype=torch.float64), x, TTA_SHIPPED)
        assert np.abs(p - r).max() < 1e-2 and (a == ra).mean() >= 0.999
    finally:
        engine.set_precision("fp32")


def test_pcm16_and_staging_paths(synth_small):
    """Host entry points: int16 PCM input (both divisors) == fp32 input bit for bit; pageable buffers staged through
    the handle's pinned slots == pinned buffers copied in place == forced 'never' staging; ragged chunk schedule."""
    eng = Engine(device=0, max_rows=2048, precision="tc")
    try:
        w = synth.synthetic_weights(195)
        eng.load_model(0, 195, w)
        eng.frontend_config(480, 160, 40, 40)
        bank, offs = synth_small["bank"], synth_small["offsets"]
        eng.set_noise_bank(torch.from_numpy(bank).cuda(), offs)
        B = 1500                                                   # several chunks of 256 clips + ramp + tail
        base, pcm0 = synth.make_clips(128, seed=2024, return_pcm=True)
        reps = (B + 127) // 128
        clips, pcm = np.tile(base, (reps, 1))[:B], np.tile(pcm0, (reps, 1))[:B]
        p = synth.make_params(B, offs, seed=2025)
        p64 = {k: v.astype(np.int64 if v.dtype.kind == "i" else np.float64) for k, v in p.items()}   # NumPy defaults
        bg = augment.gather_background(bank, offs, p["bg_index"], p["bg_offset"])
        r_aug = augment.augment_mix(clips, p["time_shift"], bg, p["bg_volume"], p["fg_volume"])
        results = {}
        for mode in ("auto", "always", "never"):
            eng.set_host_staging(mode)
            results[mode] = eng.pipeline_host(clips, p64, feat_kind="logmel", views=TTA_8)
        f0, p0, a0 = results["auto"]
        assert rel_err(f0.reshape(B, 98, 40)[:64], frontend.features(r_aug[:64], kind="logmel")) < 1e-4
        for mode in ("always", "never"):
            for got, want in zip(results[mode], (f0, p0, a0)):
                assert np.array_equal(got, want), mode
        eng.set_host_staging("auto")
        # pinned caller buffers take the in-place DMA path
        hx = torch.from_numpy(clips).pin_memory()
        hf = torch.empty((B, 98 * 40), dtype=torch.float32).pin_memory()
        hp = torch.empty((B, 12), dtype=torch.float32).pin_memory()
        ha = torch.empty((B,), dtype=torch.int32).pin_memory()
        eng.pipeline_host(hx.numpy(), p, feat_kind="logmel", views=TTA_8, feat_out=hf.numpy(), probs_out=hp.numpy(),
                          argmax_out=ha.numpy())
        assert np.array_equal(hf.numpy(), f0) and np.array_equal(hp.numpy(), p0) and np.array_equal(ha.numpy(), a0)
        # int16 PCM in: decode fused into the augment kernel's load; features left on the device
        f1, p1, a1 = eng.pipeline_host(pcm, p, feat_kind="logmel", views=TTA_8, want_features=False)
        assert f1 is None and np.array_equal(p1, p0) and np.array_equal(a1, a0)
        pp, pa = eng.predict_host(pcm, views=TTA_SHIPPED)
        qp, qa = eng.predict_host(clips, views=TTA_SHIPPED)
        assert np.array_equal(pp, qp) and np.array_equal(pa, qa)
        sp, _ = eng.predict_host(pcm[:40], views=((0, 1.0),), pcm_divisor=32767.0)   # the scipy paths' scale
        tp, _ = eng.predict_host(augment.decode_pcm16(pcm[:40], scale="scipy"), views=((0, 1.0),))
        assert np.array_equal(sp, tp)
        # buffers that would corrupt memory are refused before they reach the C ABI
        with pytest.raises(ValueError):
            eng.pipeline_host(clips, p, views=TTA_8, probs_out=np.empty((B, 11), np.float32))
        with pytest.raises(ValueError):
            eng.pipeline_host(clips, p, views=TTA_8, feat_out=np.empty((B, 98 * 40), np.float64))
        with pytest.raises(ValueError):
            eng.predict_host(clips[:, ::2])
        with pytest.raises(ValueError):
            eng.forward(torch.from_numpy(clips[:8]).cuda().double())
    finally:
        eng.close()


def test_submission_and_pseudo_surface(tmp_path, engine):
    """make_submission.py:83-213 on in-memory clips (predict -> TTA mean -> argmax -> label maps -> 3 CSVs),
    convert_from_see_v3_bugfix.py:76-110 (32 -> 12, uint8 memmap), create_pseudo_with_thresh.py:14-43,
    majority_vote.py / REPR_106_pseudo.py -- the Python surface end to end against the oracle's driver math."""
    import pandas as pd
    from speech_recognition_b200 import Model
    from speech_recognition_b200.classes import get_int2label
    engine.set_precision("fp32")
    w = synth.trained_weights(106)
    model = Model(106, w, engine=engine, slot=3)
    n = 90
    x = synth.make_word_clips(n, 32, seed=4711).numpy()
    probs, pred = submission.predict_clips(model, x, views=TTA_SHIPPED, batch_size=32)      # 3 ragged host batches
    r_probs, r_pred = driver.tta_predict(lambda v: network.forward(v, w, 106, dtype=torch.float64), x, TTA_SHIPPED)
    np.testing.assert_allclose(probs, r_probs, rtol=1e-4, atol=1e-5)
    assert np.array_equal(pred, r_pred)
    fns = [f"clip_{i:05d}.wav" for i in range(n)]
    prefix = str(tmp_path / "REPR_submission_106_tta_leftloud")
    labels, wanted = submission.write_submission_csvs(prefix, fns, probs, pred, wanted_only=False)
    int2label = get_int2label(wanted_only=False)
    all_probs = pd.read_csv(prefix + "_all_labels_probs.csv")
    assert list(all_probs.columns) == ["fname", "label"] + [int2label[i] for i in range(32)]
    assert list(pd.read_csv(prefix + ".csv").label) == wanted and list(pd.read_csv(prefix + "_all_labels.csv").label) == labels
    # convert_from_see_v3_bugfix.py: CSV probabilities -> 12 classes (Heng order) -> uint8 memmap
    cols = all_probs.iloc[:, 2:].to_numpy(np.float32)
    see, u8 = pseudo.convert_32_to_12(engine, cols, "heng")
    r_see, r_u8 = driver.convert_32_to_12(cols, "heng")
    np.testing.assert_allclose(see, r_see, rtol=2e-6, atol=1e-7)
    assert np.abs(u8.astype(int) - r_u8.astype(int)).max() <= 1
    mm_path = str(tmp_path / "submit_probs.uint8.memmap")
    submission.write_probs_memmap(mm_path, u8)
    mm = np.memmap(mm_path, dtype="uint8", mode="r", shape=(n, 12))
    preds, keep = pseudo.threshold_select(engine, np.asarray(mm), 0.6)
    r_preds, r_keep = driver.threshold_select(np.asarray(mm), 0.6)
    assert np.array_equal(preds, r_preds) and np.array_equal(keep, r_keep)
    names = pseudo.pseudo_label_names(preds)
    assert set(names) <= set(pseudo.AUDIO_NAMES) and len(names) == n
    three = np.stack([preds, np.roll(preds, 1), preds]).astype(np.int32)
    voted, clear = pseudo.majority_vote(engine, three, min_count=2)
    r_voted, r_clear = driver.majority_vote(three, 2)
    assert np.array_equal(voted, r_voted) and np.array_equal(clear, r_clear)
    assert np.array_equal(pseudo.unanimity(engine, three[0], three[1], three[2]),
                          (three[0] == three[1]) & (three[0] == three[2]))


def test_time_stretch_against_oracle(engine):
    """Speed-TTA view (create_tta_set.py:10-22): the phase-vocoder kernel against the restatement of librosa's published
    algorithm on int16 PCM in / int16 PCM out.  Transcendentals (atan2f / sincosf / hypotf) and the FFT summation order
    differ from NumPy's by rounding, and np.int16(x * 32767) truncates: a sample may land on the other side of an
    integer, never further -- |delta| <= 1 LSB, and only on a small fraction of the samples."""
    _, pcm = synth.make_clips(6, seed=321, return_pcm=True)
    pcm[0, :] = 0                                                     # digital silence
    # a loud tone over a +-20 LSB noise floor.  (Without the floor -- a bin-centred tone whose other bins are EXACTLY zero
    # until the reflect-padded last frames -- the phase accumulators of those bins integrate the angles of rounding
    # noise for 30 frames, and the output of the last 2000 samples depends on the FFT's summation order: NumPy's
    # direct transform and a two-real-frames-per-complex-transform one already differ by thousands of LSB there.
    # That is a property of the plain phase vocoder, librosa's included, not of an implementation.)
    rs = np.random.RandomState(5)
    pcm[1] = np.int16(np.round(8000 * np.sin(2 * np.pi * 1000.0 * np.arange(16000) / 16000.0) + rs.normal(0, 20, 16000)))
    ref = stretch.create_tta_batch(pcm, 0.9)
    got = engine.time_stretch(torch.from_numpy(pcm).cuda(), 0.9).cpu().numpy()
    assert got.dtype == np.int16 and got.shape == (6, 16000)
    assert np.array_equal(got[0], ref[0]) and not got[0].any()
    d = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    assert d.max() <= 1, (d.max(), np.unravel_index(d.argmax(), d.shape))
    assert (d > 0).mean() < 0.02, (d > 0).mean()
    assert np.abs(ref[1:]).max() > 1000                               # the comparison is not between silences
    host = engine.time_stretch_host(pcm, 0.9)
    assert np.array_equal(host, got)
    # other rates take the same path (output = the last 16000 samples of a longer clip)
    got75 = engine.time_stretch(torch.from_numpy(pcm[2:4]).cuda(), 0.75).cpu().numpy()
    ref75 = stretch.create_tta_batch(pcm[2:4], 0.75)
    assert np.abs(got75.astype(np.int32) - ref75.astype(np.int32)).max() <= 1
    with pytest.raises(Exception, match="rate"):
        engine.time_stretch(torch.from_numpy(pcm[:1]).cuda(), 1.5)


def test_speed_tta_end_to_end(engine):
    """make_submission.py:124-146 with use_speed_tta, the slowed set produced on the device: six views / 10 against the
    oracle (time stretch + driver arithmetic + float64 network) on a trained checkpoint."""
    from speech_recognition_b200 import Model
    engine.set_precision("fp32")
    w = synth.trained_weights(195)
    m = Model(195, w, engine=engine, slot=1)
    x, lab = synth.make_word_clips(12, 12, seed=99, return_labels=True)
    pcm = np.int16(np.round(x.numpy() * 32768.0))
    probs, amax = m.predict_speed_tta(pcm)
    xf = pcm.astype(np.float32) / np.float32(32768.0)
    slow = stretch.create_tta_batch(pcm, 0.9).astype(np.float32) / np.float32(32768.0)
    r_probs, r_amax = driver.speed_tta_predict(lambda v: network.forward(v, w, 195, dtype=torch.float64), xf, slow)
    assert np.abs(probs - r_probs).max() < 2e-3                       # +-1 LSB differences of the slowed PCM reach the softmax
    assert np.array_equal(amax, r_amax)
