"""Calibrate the BatchNorm moving statistics of the synthetic weights.

Runs the CPU oracle (float64) layer by layer on a fixed batch of synthetic clips
(plus two of the TTA gains so that 1.2x / 0.9x views stay in range) and stores,
per BatchNorm layer, the per-channel mean / variance of its input, rounded to
~12 mantissa bits so the file is reproducible.  Output:
speech_recognition_b200/data/synth_bn_<arch>.npz (keys "s<seed-arch>/<keras name>").
Development tool (uses oracle/, which is test infrastructure).
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import network  # noqa: E402
from speech_recognition_b200 import synth  # noqa: E402


def quant(v):
    m, e = np.frexp(v.astype(np.float64))
    return np.ldexp(np.round(m * 4096) / 4096, e).astype(np.float32)


def calibrate(arch, clips):
    w = synth.raw_synthetic_weights(arch)
    a = network.ARCHS[arch]
    dt = torch.float64
    x = torch.as_tensor(clips, dtype=dt)
    p = network.time_slice_stack(x)
    k1 = torch.as_tensor(w["conv1d_1/kernel"], dtype=dt)
    y = F.conv1d(p.transpose(1, 2), k1.permute(2, 1, 0).contiguous(), stride=2)
    stats = {}

    def fix(i, y):
        m = y.mean(dim=(0, 2)).numpy()
        v = y.var(dim=(0, 2), unbiased=False).numpy()
        w[f"batch_normalization_{i}/moving_mean"] = quant(m)
        w[f"batch_normalization_{i}/moving_variance"] = quant(np.maximum(v, 1e-6))
        stats[f"batch_normalization_{i}/moving_mean"] = w[f"batch_normalization_{i}/moving_mean"]
        stats[f"batch_normalization_{i}/moving_variance"] = w[f"batch_normalization_{i}/moving_variance"]
        return network._bn_relu6(y, w, i, dt)
    y = fix(1, y)
    for i, (co, s) in enumerate(a["blocks"], start=1):
        dk = torch.as_tensor(w[f"depthwise_conv2d_{i}/depthwise_kernel"], dtype=dt)
        C = dk.shape[2]
        dkt = dk[0, :, :, 0].t().reshape(C, 1, 3).contiguous()
        if s == 1:
            y = F.conv1d(y, dkt, groups=C)
        else:
            _, pl, pr = network.same_pad(y.shape[-1], 3, 2)
            y = F.conv1d(F.pad(y, (pl, pr)), dkt, stride=2, groups=C)
        pk = torch.as_tensor(w[f"conv1d_{i + 1}/kernel"], dtype=dt)
        y = F.conv1d(y, pk[0].t().reshape(co, C, 1).contiguous())
        y = fix(i + 1, y)
    return stats, w


def main():
    clips = synth.make_clips(96, seed=synth.SEED + 1000)
    clips = np.concatenate([clips, 1.2 * clips[:16], 0.9 * clips[16:32]]).astype(np.float32)
    for key, seeds in ((195, (195, 206)), (106, (106,)), (716, (716,))):
        out = {}
        for s in seeds:
            stats, w = calibrate(s, clips)
            for k, v in stats.items():
                out[f"s{s}/{k}"] = v
            probs = network.forward(clips[:96], w, s)
            print(s, "argmax histogram", np.bincount(probs.argmax(1), minlength=probs.shape[1]),
                  "max-prob quantiles", np.quantile(probs.max(1), [0.1, 0.5, 0.9]))
        path = os.path.join(os.path.dirname(os.path.abspath(synth.__file__)), "data",
                            f"synth_bn_{key}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
