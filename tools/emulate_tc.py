"""CPU emulation of the tensor-core tier's NUMERICS (development tool; uses oracle/, test infrastructure).

Mirrors, rounding for rounding, what csrc/tc_net.cu computes -- fp16 waveform window, fp16 weights with the
BatchNorm scale folded in, fp32 accumulation, fmaf(acc, gain, shift) -> ReLU6 -> fp16 storage between the
layers, the depthwise FIR either as the packed-half chain (hmul2, hfma2, hfma2) or with fp32 accumulation and one
rounding -- so that the label-agreement rate against the float64 oracle and the effect of a numerics change
can be estimated without a GPU.  Summation ORDER inside a GEMM differs from the tensor core's (fp32 either way).

  python tools/emulate_tc.py --arch 195 --clips 1024 [--weights trained]
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import network, driver  # noqa: E402
from speech_recognition_b200 import synth, TTA_8  # noqa: E402


def h16(t):
    return t.to(torch.float16).to(torch.float32)


def bn_fold(w, i):
    g, b = w[f"batch_normalization_{i}/gamma"], w[f"batch_normalization_{i}/beta"]
    m, v = w[f"batch_normalization_{i}/moving_mean"], w[f"batch_normalization_{i}/moving_variance"]
    s = (1.0 / np.sqrt(v.astype(np.float32) + np.float32(1e-3))).astype(np.float32) * g
    return s.astype(np.float32), (b - m * s).astype(np.float32)


def relu6_h(y):
    return torch.clamp(h16(torch.clamp(y, min=0.0)), max=6.0)


def forward_tc(x, w, arch, views, fir="fp16"):
    """x [B,16000] f32 -> mean probabilities [B,C] following the TC tier's roundings."""
    a = network.ARCHS[arch]
    B = x.shape[0]
    probs_sum = None
    xt = torch.as_tensor(x)
    s1, sh1 = bn_fold(w, 1)
    k1 = torch.as_tensor(w["conv1d_1/kernel"]) * torch.as_tensor(s1)[None, None, :]
    # fold the 3 overlapping patches into 80 taps (model_build_tc), scale folded BEFORE the fp16 rounding
    c0 = k1.shape[2]
    w80 = torch.zeros((80, c0))
    k1u = torch.as_tensor(w["conv1d_1/kernel"])
    for f in range(3):
        w80[20 * f:20 * f + 40] += k1u[f]
    w80 = h16(w80 * torch.as_tensor(s1)[None, :])
    out = []
    for shift, gain in views:
        xv = h16(torch.roll(xt, int(shift), dims=1))                 # fp16 window, gain applied in the epilogue
        xp = F.pad(xv, (10, 70))
        # rows j: samples [40 j - 10, 40 j + 70)
        win = xp.unfold(1, 80, 40)[:, :399]                          # [B,399,80]
        acc = win @ w80                                              # fp32 accumulate
        y = relu6_h(acc * np.float32(gain) + torch.as_tensor(sh1))   # fmaf(acc, gain, shift)
        y = y.transpose(1, 2).contiguous()                           # [B,C,T]
        for i, (co, s) in enumerate(a["blocks"], start=1):
            dk = torch.as_tensor(w[f"depthwise_conv2d_{i}/depthwise_kernel"])[0, :, :, 0]   # [3,C]
            C = dk.shape[1]
            dk = h16(dk)
            if s == 2:
                _, pl, pr = network.same_pad(y.shape[-1], 3, 2)
                y = F.pad(y, (pl, pr))
            T = y.shape[-1]
            n = (T - 3) // s + 1
            x0 = y[:, :, 0:0 + s * (n - 1) + 1:s]
            x1 = y[:, :, 1:1 + s * (n - 1) + 1:s]
            x2 = y[:, :, 2:2 + s * (n - 1) + 1:s]
            k0, k1_, k2 = dk[0][None, :, None], dk[1][None, :, None], dk[2][None, :, None]
            if fir == "fp16":
                t = h16(x0 * k0)                                     # hmul2
                t = (x1.double() * k1_.double() + t.double()).to(torch.float16).to(torch.float32)   # hfma2
                t = (x2.double() * k2.double() + t.double()).to(torch.float16).to(torch.float32)
            else:
                t = h16(x2 * k2 + (x1 * k1_ + x0 * k0))             # fp32 accumulate, one rounding
            sc, sh = bn_fold(w, i + 1)
            pk = h16(torch.as_tensor(w[f"conv1d_{i + 1}/kernel"])[0] * torch.as_tensor(sc)[None, :])   # [Cin,Cout]
            acc = torch.einsum("bct,co->bot", t, pk)
            y = relu6_h(acc + torch.as_tensor(sh)[None, :, None])
        # head in fp32 on the fp16 activations
        xl = y.transpose(1, 2).contiguous()
        Bv, T, C = xl.shape
        att = xl.reshape(Bv, T * C) @ torch.as_tensor(w["dense_1/kernel"])
        if a["dense1_bias"]:
            att = att + torch.as_tensor(w["dense_1/bias"])
        att = torch.softmax(att, dim=-1)
        wt = xl * att[:, :, None]
        z = torch.cat([wt.max(dim=1).values, xl.mean(dim=1)], dim=1) if a["pool"] == "max_avg" else wt.mean(dim=1)
        p = torch.softmax(z @ torch.as_tensor(w["dense_2/kernel"]), dim=-1)
        out.append(p)
    probs_sum = out[0].clone()
    for p in out[1:]:
        probs_sum = probs_sum + p
    return (probs_sum / np.float32(len(views))).numpy()


def report(name, got, ref):
    err = np.abs(got - ref)
    lab, rlab = got.argmax(1), ref.argmax(1)
    srt = np.sort(ref, axis=1)
    margin = srt[:, -1] - srt[:, -2]
    dis = lab != rlab
    print(f"{name}: N={len(ref)} agreement {1 - dis.mean():.5f}  |dp| p99 {np.quantile(err, 0.99):.2e} "
          f"p99.9 {np.quantile(err, 0.999):.2e} max {err.max():.2e}  max margin among flips "
          f"{margin[dis].max() if dis.any() else 0:.3e}  median max-prob {np.median(ref.max(1)):.3f}")
    for lo, hi in ((0, 1e-3), (1e-3, 1e-2), (1e-2, 1e-1), (1e-1, 1.01)):
        sel = (margin >= lo) & (margin < hi)
        if sel.any():
            print(f"   margin [{lo:g},{hi:g}): {sel.sum():6d} clips, {dis[sel].sum():4d} flips")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", type=int, default=195)
    ap.add_argument("--clips", type=int, default=512)
    ap.add_argument("--seed", type=int, default=4242)
    ap.add_argument("--weights", default="random", choices=["random", "trained"])
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    if args.weights == "random":
        w = synth.synthetic_weights(args.arch)
        x = synth.make_clips(args.clips, seed=args.seed)
    else:
        w = synth.trained_weights(args.arch)
        x = synth.make_word_clips(args.clips, network.ARCHS[args.arch]["classes"], seed=args.seed).numpy()
    t0 = time.time()
    ref, _ = driver.tta_predict(lambda v: network.forward(v, w, args.arch, dtype=torch.float64), x, TTA_8)
    print(f"oracle f64: {time.time() - t0:.1f} s")
    for fir in ("fp16", "fp32"):
        t0 = time.time()
        got = forward_tc(x, w, args.arch, TTA_8, fir=fir)
        report(f"arch {args.arch} FIR {fir} ({time.time() - t0:.0f} s)", got, ref)


if __name__ == "__main__":
    main()
