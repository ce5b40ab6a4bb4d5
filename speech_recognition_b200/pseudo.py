"""Pseudo-labelling / ensembling arithmetic on the device (reference
create_pseudo_with_thresh.py:14-43, convert_from_see_v3_bugfix.py:76-110,
majority_vote.py:26-56, REPR_106_pseudo.py:12).  NumPy in, NumPy out; the compute
runs in libkws.so (kws_convert_classes / kws_select / kws_vote)."""
from __future__ import annotations

import numpy as np

from .classes import class_map_32_to_12, AUDIO_NAMES
from .engine import Engine


def _dev(engine, a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(f"cuda:{engine.device}")


def convert_32_to_12(engine: Engine, all_probs: np.ndarray, order='heng'):
    """-> (see_probs f32 [N,12], uint8 [N,12])."""
    out, u8 = engine.convert_classes(_dev(engine, np.asarray(all_probs, np.float32)), class_map_32_to_12(order), 12)
    return out.cpu().numpy(), u8.cpu().numpy()


def threshold_select(engine: Engine, probs_u8: np.ndarray, prob_thresh: float):
    """-> (preds int32 [N], keep bool [N]); a clip is dropped iff float32(max)/255 < prob_thresh."""
    label, keep = engine.select(_dev(engine, np.asarray(probs_u8, np.uint8)), prob_thresh)
    return label.cpu().numpy(), keep.cpu().numpy().astype(bool)


def pseudo_label_names(preds):
    return [AUDIO_NAMES[int(p)] for p in preds]


def majority_vote(engine: Engine, labels: np.ndarray, min_count=3):
    """labels int [M,N] -> (voted int32 [N], clear_majority bool [N])."""
    voted, clear = engine.vote(_dev(engine, np.asarray(labels, np.int32)), min_count)
    return voted.cpu().numpy(), clear.cpu().numpy().astype(bool)


def unanimity(engine: Engine, a, b, c):
    """REPR_106_pseudo.py:12: (a == b) & (a == c)."""
    _, clear = majority_vote(engine, np.stack([a, b, c]), min_count=3)
    return clear
