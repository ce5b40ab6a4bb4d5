"""``load_model`` / ``Model.predict`` -- the Keras-shaped surface of the reference
(make_submission.py:64-71,120-146) on top of libkws.so."""
from __future__ import annotations

import numpy as np

from .engine import Engine
from .weights import load_weights, validate, infer_arch

# make_submission.py:126-128, in the order they are summed at :142-144
TTA_SHIPPED = ((0, 1.0), (0, 1.2), (-1500, 1.0))
# synthetic 8-view table of SURVEY.md 8d (built from make_submission.py:126-130 primitives)
TTA_8 = ((0, 1.0), (-1500, 1.0), (0, 1.2), (0, 0.9), (0, -1.0), (-1500, 1.2), (-1500, 0.9), (-3000, 1.0))


class Model:
    """Stateless, synchronous ``predict`` like keras.Model (inference mode: BN moving
    statistics, dropout off -- K.set_learning_phase(0), make_submission.py:37)."""

    def __init__(self, arch: int, weights: dict, engine: Engine | None = None, slot: int = 0,
                 device: int = 0, precision="tc"):
        self.engine = engine if engine is not None else Engine(device=device, precision=precision)
        self.slot = slot
        self.arch = arch
        self.weights = validate(arch, weights)
        self.num_classes = self.engine.load_model(slot, arch, self.weights)

    def predict(self, x, batch_size=32, verbose=0, views=((0, 1.0),)):
        """x: float32 [N,16000] -> float32 [N,C] softmax probabilities.  ``batch_size`` is
        accepted for API compatibility; the device batches internally."""
        x = np.asarray(x)
        x = np.ascontiguousarray(x, np.int16 if x.dtype == np.int16 else np.float32)   # int16 PCM is decoded on the device
        if x.ndim != 2 or x.shape[1] != 16000:
            raise ValueError("Error when checking input: expected input_1 to have shape (None, 16000) "
                             "but got array with shape %s" % (x.shape,))
        probs, _ = self.engine.predict_host(x, views=views, slot=self.slot)
        return probs

    def predict_tta(self, x, views=TTA_SHIPPED):
        """(probs + loud_probs + left_probs) / 3 and argmax (make_submission.py:120-146)."""
        return self.engine.predict_host(x, views=views, slot=self.slot)

    def predict_speed_tta(self, x, x_slow=None, tta_speed=0.9):
        """make_submission.py:124-146 with ``use_speed_tta``: x_slow holds the time-stretched copies that
        create_tta_set.py wrote (librosa, offline).  Views: x, roll(x, -1500), 1.2 x, x_slow,
        clip(1.1 x_slow, -1, 1), 0.9 x_slow; the reference divides the SIX probability vectors by 10
        (sic, :140) -- kept, it does not change the argmax.  Returns (probs f32 [N,C], argmax int32 [N]).

        With ``x_slow=None`` the slowed set is produced here, on the device, the way create_tta_set.py:16-22 produces
        its WAV files: ``x`` must then be the clips' int16 PCM; both sets are decoded like the reference's DecodeWav
        (/ 32768, input_data.py:334-336)."""
        import torch
        eng = self.engine
        dev = f"cuda:{eng.device}"
        if x_slow is None:
            x = np.asarray(x)
            if x.dtype != np.int16:
                raise ValueError("predict_speed_tta without x_slow needs the int16 PCM of the clips")
            pcm = torch.from_numpy(np.ascontiguousarray(x)).to(dev)
            slow = eng.time_stretch(pcm, tta_speed)
            xt = pcm.to(torch.float32) * (1.0 / 32768.0)
            st = slow.to(torch.float32) * (1.0 / 32768.0)
        else:
            xt = torch.from_numpy(np.ascontiguousarray(x, np.float32)).to(dev)
            st = torch.from_numpy(np.ascontiguousarray(x_slow, np.float32)).to(dev)
        n = xt.shape[0]
        p_a, _ = eng.forward(xt, views=TTA_SHIPPED, slot=self.slot)                 # mean of 3
        p_b, _ = eng.forward(st, views=((0, 1.0), (0, 0.9)), slot=self.slot)        # mean of 2
        zeros_i = torch.zeros((n,), dtype=torch.int32, device=dev)
        loud = eng.augment(st, zeros_i, zeros_i - 1, zeros_i, torch.zeros((n,), dtype=torch.float32, device=dev),
                           torch.full((n,), 1.1, dtype=torch.float32, device=dev), clamp=True)   # clip(1.1 x, -1, 1)
        p_c, _ = eng.forward(loud, views=((0, 1.0),), slot=self.slot)
        probs = (p_a * 3.0 + p_b * 2.0 + p_c) / 10.0
        return probs.cpu().numpy(), probs.argmax(dim=1).to(torch.int32).cpu().numpy()


def load_model(filepath, custom_objects=None, engine: Engine | None = None, slot: int = 0, **kw):
    """keras.models.load_model stand-in: .npz / Keras .hdf5 / frozen .pb -> Model.
    ``custom_objects`` is accepted and ignored (relu6, DepthwiseConv2D, ... are built in)."""
    arch, weights = load_weights(filepath)
    return Model(arch, weights, engine=engine, slot=slot, **kw)
