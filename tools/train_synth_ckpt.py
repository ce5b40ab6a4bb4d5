"""Train a checkpoint of the exp-195 / exp-106 architecture on the synthetic keyword task.

The reference's trained checkpoints are missing from the mount and random weights make a
chaotic network: it amplifies input and rounding noise from layer to layer and its softmax sits near ties, so
"label agreement" on it measures tie-breaking, not numerics.  This tool trains the torch restatement of
`conv_1d_time_sliced_with_attention_model` (model.py:775-838; exp 106: the logs_106 variant) on
`synth.make_word_clips` -- RMSprop(1e-3) and label smoothing 0.1 as in model.py:833-836, dropout 0.4 on the two
head inputs, augmentation in the spirit of input_data.py:457-514 (time shift, gain, sign flip, background
noise) wide enough to cover the TTA views -- then replaces the BatchNorm moving statistics by exact statistics
over a calibration set and writes the Keras-named tensors as float16 to
speech_recognition_b200/data/trained_<arch>.npz.  Development tool (a few minutes of CPU per architecture).
"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import network, driver  # noqa: E402
from speech_recognition_b200 import synth, TTA_8  # noqa: E402


class Net(nn.Module):
    def __init__(self, arch):
        super().__init__()
        a = network.ARCHS[arch]
        self.a = a
        c = a["conv1"]
        self.conv1 = nn.Conv1d(40, c, 3, stride=2, bias=False)
        self.bn = nn.ModuleList([nn.BatchNorm1d(c, eps=1e-3, momentum=0.01)])
        self.dw, self.pw = nn.ModuleList(), nn.ModuleList()
        for co, s in a["blocks"]:
            self.dw.append(nn.Conv1d(c, c, 3, stride=s, groups=c, bias=False))
            self.pw.append(nn.Conv1d(c, co, 1, bias=False))
            self.bn.append(nn.BatchNorm1d(co, eps=1e-3, momentum=0.01))
            c = co
        T = network.layer_lengths(arch)[-1]
        self.T, self.C = T, c
        self.d1 = nn.Linear(T * c, T, bias=a["dense1_bias"])
        feat = 2 * c if a["pool"] == "max_avg" else c
        self.d2 = nn.Linear(feat, a["classes"], bias=False)
        self.drop = nn.Dropout(0.4)

    def forward(self, x):
        p = network.time_slice_stack(x)
        y = torch.clamp(self.bn[0](self.conv1(p.transpose(1, 2))), 0.0, 6.0)
        for i, (co, s) in enumerate(self.a["blocks"]):
            if s == 2:
                _, pl, pr = network.same_pad(y.shape[-1], 3, 2)
                y = F.pad(y, (pl, pr))
            y = torch.clamp(self.bn[i + 1](self.pw[i](self.dw[i](y))), 0.0, 6.0)
        xt = y.transpose(1, 2)                                     # [B,T,C]
        att = torch.softmax(self.d1(self.drop(xt.reshape(xt.shape[0], -1))), dim=-1)
        wt = xt * att[:, :, None]
        z = torch.cat([wt.max(dim=1).values, xt.mean(dim=1)], dim=1) if self.a["pool"] == "max_avg" else wt.mean(dim=1)
        return self.d2(self.drop(z))

    def export(self):
        """Keras-named tensors (layouts of SURVEY.md 8b)."""
        w = {"conv1d_1/kernel": self.conv1.weight.detach().permute(2, 1, 0).numpy()}           # [3,40,C0]
        def bn(i, m):
            w[f"batch_normalization_{i}/gamma"] = m.weight.detach().numpy()
            w[f"batch_normalization_{i}/beta"] = m.bias.detach().numpy()
            w[f"batch_normalization_{i}/moving_mean"] = m.running_mean.numpy()
            w[f"batch_normalization_{i}/moving_variance"] = m.running_var.numpy()
        bn(1, self.bn[0])
        for i in range(len(self.dw)):
            C = self.dw[i].weight.shape[0]
            w[f"depthwise_conv2d_{i + 1}/depthwise_kernel"] = self.dw[i].weight.detach().reshape(C, 3).t().reshape(1, 3, C, 1).numpy()
            w[f"conv1d_{i + 2}/kernel"] = self.pw[i].weight.detach()[:, :, 0].t().reshape(1, C, -1).numpy()
            bn(i + 2, self.bn[i + 1])
        w["dense_1/kernel"] = self.d1.weight.detach().t().numpy()
        if self.d1.bias is not None:
            w["dense_1/bias"] = self.d1.bias.detach().numpy()
        w["dense_2/kernel"] = self.d2.weight.detach().t().numpy()
        return {k: np.ascontiguousarray(v, np.float32) for k, v in w.items()}


def augment(x, g, bank):
    n = x.shape[0]
    shift = torch.randint(-3300, 300, (n,), generator=g)
    idx = (torch.arange(16000)[None, :] - shift[:, None]) % 16000
    x = torch.gather(x, 1, idx)
    gain = 0.8 + 0.5 * torch.rand((n, 1), generator=g)
    sign = torch.where(torch.rand((n, 1), generator=g) < 0.4, -1.0, 1.0)
    off = torch.randint(0, bank.shape[0] - 16000, (n,), generator=g)
    bg = torch.stack([bank[o:o + 16000] for o in off.tolist()])
    vol = torch.where(torch.rand((n, 1), generator=g) < 0.3, 0.15 * torch.rand((n, 1), generator=g), torch.zeros(n, 1))
    return x * gain * sign + bg * vol


def train(arch, steps, batch=64, pool=8192):
    """arch 206 = the exp-195 architecture trained from another seed on another draw of the clips (the reference's
    exp 206 is exactly that: the same model retrained, README.md)."""
    torch.manual_seed(arch)
    g = torch.Generator().manual_seed(arch + 1)
    C = network.ARCHS[arch]["classes"]
    x_all, y_all = synth.make_word_clips(pool, C, seed=synth.SEED + 31 * arch, return_labels=True)
    bank = torch.from_numpy(synth.make_noise_bank(seconds=4)[0])
    net = Net(arch)
    opt = torch.optim.RMSprop(net.parameters(), lr=1e-3, alpha=0.9, eps=1e-7, weight_decay=1e-5)
    t0 = time.time()
    net.train()
    for step in range(steps):
        for gq in opt.param_groups:
            gq["lr"] = 1e-3 * (0.1 if step > 0.8 * steps else 1.0)
        i = torch.randint(0, pool, (batch,), generator=g)
        xb, yb = augment(x_all[i], g, bank), y_all[i]
        loss = F.cross_entropy(net(xb), yb, label_smoothing=0.1)
        opt.zero_grad(); loss.backward(); opt.step()
        if step % 50 == 0 or step == steps - 1:
            print(f"arch {arch} step {step} loss {loss.item():.3f} ({time.time() - t0:.0f} s)", flush=True)
    # exact BatchNorm statistics over a calibration set (the moving averages of a short run lag behind)
    for m in net.bn:
        m.reset_running_stats(); m.momentum = None
    with torch.no_grad():
        for k in range(8):
            i = torch.randint(0, pool, (128,), generator=g)
            net(augment(x_all[i], g, bank))
    net.eval()
    w = net.export()
    w16 = {k: v.astype(np.float16) for k, v in w.items()}
    w = {k: v.astype(np.float32) for k, v in w16.items()}
    # held-out check through the float64 oracle and the TTA views
    xt, yt = synth.make_word_clips(512, C, seed=synth.SEED + 77 * arch, return_labels=True)
    p, lab = driver.tta_predict(lambda v: network.forward(v, w, arch, dtype=torch.float64), xt.numpy(), TTA_8)
    srt = np.sort(p, axis=1)
    print(f"arch {arch}: held-out accuracy (8-view TTA, oracle f64) {(lab == yt.numpy()).mean():.4f}, median max-prob "
          f"{np.median(p.max(1)):.3f}, top-2 margin < 0.01 on {((srt[:, -1] - srt[:, -2]) < 0.01).mean():.4f} of the clips")
    path = os.path.join(os.path.dirname(os.path.abspath(synth.__file__)), "data", f"trained_{arch}.npz")
    np.savez_compressed(path, **w16)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    for arch in ([int(sys.argv[1])] if len(sys.argv) > 1 else [195, 106, 206]):
        train(arch, steps)
