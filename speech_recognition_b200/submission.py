"""Inference driver: batched predict + TTA mean + argmax + label maps + the three CSVs
and the uint8 probability memmap (reference make_submission.py:34-213,
convert_from_see_v3_bugfix.py:94-110)."""
from __future__ import annotations

import csv

import numpy as np

from .classes import get_classes, get_int2label, prepare_words_list, map_to_valid, map_to_wanted
from .model import Model, TTA_SHIPPED


def predict_clips(model: Model, clips: np.ndarray, views=TTA_SHIPPED, batch_size=384 * 16):
    """make_submission.py:83-153 for in-memory clips: returns (probs f32 [N,C], pred int [N])."""
    N = len(clips)
    probs = np.empty((N, model.num_classes), np.float32)
    pred = np.empty((N,), np.int32)
    for s in range(0, N, batch_size):
        e = min(N, s + batch_size)
        model.engine.predict_host(clips[s:e], views=views, slot=model.slot, probs_out=probs[s:e],
                                  argmax_out=pred[s:e])
    return probs, pred


def labels_from_pred(pred, wanted_only=False):
    """int2label -> map_to_valid -> map_to_wanted (make_submission.py:147-153)."""
    int2label = get_int2label(wanted_only=wanted_only)
    wanted_words = prepare_words_list(get_classes(wanted_only=True))
    labels = map_to_valid([int2label[int(p)] for p in pred])
    return labels, map_to_wanted(labels, wanted_words)


def write_submission_csvs(prefix, fnames, probs, pred, wanted_only=False):
    """The three CSVs of make_submission.py:198-212: <prefix>.csv (12-way labels),
    <prefix>_all_labels.csv, <prefix>_all_labels_probs.csv (one float column per class)."""
    labels, wanted = labels_from_pred(pred, wanted_only)
    int2label = get_int2label(wanted_only=wanted_only)
    with open(prefix + '.csv', 'w', newline='') as f:
        w = csv.writer(f); w.writerow(['fname', 'label']); w.writerows(zip(fnames, wanted))
    with open(prefix + '_all_labels.csv', 'w', newline='') as f:
        w = csv.writer(f); w.writerow(['fname', 'label']); w.writerows(zip(fnames, labels))
    with open(prefix + '_all_labels_probs.csv', 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['fname', 'label'] + [int2label[i] for i in range(len(int2label))])
        for fn, l, p in zip(fnames, labels, probs):
            w.writerow([fn, l] + [str(v) for v in np.asarray(p, np.float32)])   # pandas' float32 text: shortest round-trip repr
    return labels, wanted


def write_probs_memmap(path, probs_u8):
    """np.memmap(uint8, shape=(N,12)) like convert_from_see_v3_bugfix.py:107-110."""
    mm = np.memmap(path, dtype='uint8', mode='w+', shape=probs_u8.shape)
    mm[...] = probs_u8
    mm.flush()
    return mm
