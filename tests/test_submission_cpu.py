"""Output formats of the inference driver (make_submission.py:147-153,198-212; convert_from_see_v3_bugfix.py:107-110):
the three CSVs byte for byte against pandas' DataFrame.to_csv -- the call the reference makes -- and the uint8
probability memmap.  No GPU: the writers take probabilities / predictions as arrays."""
import io
import os

import numpy as np
import pandas as pd
import pytest

from speech_recognition_b200 import submission
from speech_recognition_b200.classes import get_int2label, get_classes, prepare_words_list, map_to_valid, map_to_wanted


def reference_csvs(fns, probs, pred, wanted_only):
    """make_submission.py:147-153,198-212 restated with the same pandas calls."""
    int2label = get_int2label(wanted_only=wanted_only)
    wanted_words = prepare_words_list(get_classes(wanted_only=True))
    labels = map_to_valid([int2label[int(p)] for p in pred])
    wanted = map_to_wanted(labels, wanted_words)
    out = []
    for frame in (pd.DataFrame({'fname': fns, 'label': wanted}), pd.DataFrame({'fname': fns, 'label': labels})):
        s = io.StringIO(); frame.to_csv(s, index=False, compression=None); out.append(s.getvalue())
    all_data = pd.DataFrame({'fname': fns, 'label': labels})
    for i, l in int2label.items():
        all_data[l] = probs[:, i]
    s = io.StringIO(); all_data.to_csv(s, index=False, compression=None); out.append(s.getvalue())
    return out


@pytest.mark.parametrize("wanted_only,C", [(False, 32), (True, 12)])
def test_submission_csvs_match_pandas(tmp_path, wanted_only, C):
    rs = np.random.RandomState(3)
    n = 200
    probs = rs.dirichlet(np.ones(C) * 0.05, size=n).astype(np.float32)     # includes 1e-30-sized and ~1.0 entries
    probs[0] = 0.0; probs[0, 3] = 1.0
    pred = probs.argmax(1)
    fns = [f"clip_{i:08x}.wav" for i in range(n)]
    prefix = str(tmp_path / "REPR_submission")
    labels, wanted = submission.write_submission_csvs(prefix, fns, probs, pred, wanted_only=wanted_only)
    ref = reference_csvs(fns, probs, pred, wanted_only)
    for suffix, want in zip((".csv", "_all_labels.csv", "_all_labels_probs.csv"), ref):
        with open(prefix + suffix, newline='') as f:
            got = f.read()
        assert got.replace("\r\n", "\n") == want, suffix
    assert set(wanted) <= set(['silence', 'unknown'] + get_classes(wanted_only=True))
    assert len(labels) == n
    # the files read back the way create_pseudo / majority_vote read them (pd.read_csv)
    back = pd.read_csv(prefix + "_all_labels_probs.csv")
    assert list(back.columns[:2]) == ['fname', 'label'] and back.shape == (n, 2 + C)
    np.testing.assert_array_equal(back.iloc[:, 2:].to_numpy(np.float32), probs)


def test_probs_memmap_roundtrip(tmp_path):
    rs = np.random.RandomState(4)
    u8 = rs.randint(0, 256, (1000, 12)).astype(np.uint8)
    path = str(tmp_path / "submit_probs.uint8.memmap")
    mm = submission.write_probs_memmap(path, u8)
    del mm
    assert os.path.getsize(path) == 1000 * 12
    back = np.memmap(path, dtype='uint8', mode='r', shape=(1000, 12))        # create_pseudo_with_thresh.py:15-16
    assert np.array_equal(np.asarray(back), u8)
