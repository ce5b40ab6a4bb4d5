"""Oracle, stage 2: raw-waveform Depthwise1D / separable-conv networks
(TEST INFRASTRUCTURE ONLY).

Restates with torch-CPU ops (channels-last semantics of Keras kept explicit):
  * ``overlapping_time_slice_stack``                (model.py:67-76)
  * ``_depthwise_conv_block``                       (model.py:34-52)
  * ``conv_1d_time_sliced_with_attention_model``    (model.py:775-838)  = exp 195 / 206
  * the exp-106 variant recovered from the logs_106 GraphDef (not in HEAD model.py)
BatchNorm is the Keras/TF inference form (graph nodes
batch_normalization_*/cond/batchnorm/*): s = rsqrt(var + 1e-3) * gamma;
y = x*s + (beta - mean*s).  ReLU6 = model.py:30-31.

Weights are a dict keyed by the Keras variable names of the reference graph
(``conv1d_1/kernel`` [3,40,C0], ``depthwise_conv2d_k/depthwise_kernel`` [1,3,C,1],
``conv1d_k/kernel`` [1,Cin,Cout], ``batch_normalization_k/{gamma,beta,
moving_mean,moving_variance}``, ``dense_1/kernel`` (+``dense_1/bias``),
``dense_2/kernel``).  ``tests/golden/graph_net_*.npz`` (the reference GraphDef
evaluated node-by-node on the same synthetic weights) pins this file.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3

# (out_channels, stride) of the 11 depthwise-separable blocks after conv1d_1.
ARCHS = {
    # model.py:807-817 with filter_mult=1
    195: dict(conv1=128, blocks=[(128, 1), (192, 2), (192, 1), (256, 2), (256, 1), (320, 2),
                                 (320, 1), (384, 2), (384, 1), (512, 2), (512, 1)],
              dense1_bias=True, pool="max_avg", classes=12),
    # logs_106 GraphDef
    106: dict(conv1=64, blocks=[(128, 1), (192, 2), (192, 1), (256, 2), (256, 1), (320, 2),
                                (320, 1), (384, 2), (384, 1), (448, 2), (448, 1)],
              dense1_bias=False, pool="attn_mean", classes=32),
}
ARCHS[206] = ARCHS[195]
# conv_1d_time_sliced_model(filter_mult=1), model.py:716-772
ARCHS[716] = dict(conv1=32, blocks=[(64, 1), (128, 2), (128, 1), (192, 2), (192, 1), (256, 2), (256, 1), (320, 2),
                                    (320, 1), (384, 2), (384, 1), (512, 2), (512, 1)],
                  dense1_bias=False, pool="gap_dense", hidden=256, classes=12)


def same_pad(T: int, k: int, s: int):
    """TF 'SAME' rule: out = ceil(T/s); pad_total = max((out-1)*s + k - T, 0);
    pad_left = pad_total // 2 (the extra sample goes right)."""
    out = -(-T // s)
    total = max((out - 1) * s + k - T, 0)
    return out, total // 2, total - total // 2


def layer_lengths(arch: int, input_size: int = 16000):
    """Time lengths after the patch stack, conv1d_1 and each block."""
    a = ARCHS[arch]
    n_patch, _, _ = same_pad(input_size, 40, 20)
    T = (n_patch - 3) // 2 + 1
    out = [n_patch, T]
    for _, s in a["blocks"]:
        T = T - 2 if s == 1 else same_pad(T, 3, 2)[0]
        out.append(T)
    return out


def weight_shapes(arch: int):
    """Ordered {keras_name: shape} for an architecture (74 tensors for 195, 73 for 106)."""
    a = ARCHS[arch]
    shapes = {}
    c = a["conv1"]
    shapes["conv1d_1/kernel"] = (3, 40, c)

    def bn(i, ch):
        for nm in ("gamma", "beta", "moving_mean", "moving_variance"):
            shapes[f"batch_normalization_{i}/{nm}"] = (ch,)
    bn(1, c)
    for i, (co, _) in enumerate(a["blocks"], start=1):
        shapes[f"depthwise_conv2d_{i}/depthwise_kernel"] = (1, 3, c, 1)
        shapes[f"conv1d_{i + 1}/kernel"] = (1, c, co)
        bn(i + 1, co)
        c = co
    if a["pool"] == "gap_dense":
        shapes["dense_1/kernel"] = (c, a["hidden"])
        shapes["dense_2/kernel"] = (a["hidden"], a["classes"])
        return shapes
    T_last = layer_lengths(arch)[-1]
    shapes["dense_1/kernel"] = (T_last * c, T_last)
    if a["dense1_bias"]:
        shapes["dense_1/bias"] = (T_last,)
    feat = 2 * c if a["pool"] == "max_avg" else c
    shapes["dense_2/kernel"] = (feat, a["classes"])
    return shapes


def _bn_relu6(x, w, i, dtype):
    g = torch.as_tensor(w[f"batch_normalization_{i}/gamma"], dtype=dtype)
    b = torch.as_tensor(w[f"batch_normalization_{i}/beta"], dtype=dtype)
    m = torch.as_tensor(w[f"batch_normalization_{i}/moving_mean"], dtype=dtype)
    v = torch.as_tensor(w[f"batch_normalization_{i}/moving_variance"], dtype=dtype)
    s = torch.rsqrt(v + BN_EPS) * g
    y = x * s[None, :, None] + (b - m * s)[None, :, None]
    return torch.clamp(y, 0.0, 6.0)


def time_slice_stack(x: torch.Tensor, ksize: int = 40, stride: int = 20):
    """model.py:67-76: extract_image_patches SAME -> [N, n_patch, ksize];
    P[j,i] = x[stride*j - pad_left + i], zero outside."""
    N, W = x.shape
    n, pl, pr = same_pad(W, ksize, stride)
    xp = F.pad(x, (pl, pr))
    return xp.unfold(1, ksize, stride)[:, :n]


def forward(x, w, arch: int = 195, dtype=torch.float32, return_activations: bool = False):
    """x [B,16000] -> softmax probabilities [B,C].  Internally channels-first
    [B,C,T] (torch); every contraction is the same sum as the Keras channels-last op."""
    a = ARCHS[arch]
    x = torch.as_tensor(np.asarray(x), dtype=dtype)
    acts = []
    p = time_slice_stack(x)                                   # [B,800,40]
    k1 = torch.as_tensor(w["conv1d_1/kernel"], dtype=dtype)   # [3,40,C0] (f,i,co)
    # Conv1D over the patch axis with the 40 patch samples as channels.
    y = F.conv1d(p.transpose(1, 2), k1.permute(2, 1, 0).contiguous(), stride=2)  # [B,C0,399]
    y = _bn_relu6(y, w, 1, dtype)
    acts.append(y)
    for i, (co, s) in enumerate(a["blocks"], start=1):
        dk = torch.as_tensor(w[f"depthwise_conv2d_{i}/depthwise_kernel"], dtype=dtype)  # [1,3,C,1]
        C = dk.shape[2]
        dkt = dk[0, :, :, 0].t().reshape(C, 1, 3).contiguous()
        if s == 1:
            y = F.conv1d(y, dkt, groups=C)                    # VALID
        else:
            _, pl, pr = same_pad(y.shape[-1], 3, 2)
            y = F.conv1d(F.pad(y, (pl, pr)), dkt, stride=2, groups=C)
        pk = torch.as_tensor(w[f"conv1d_{i + 1}/kernel"], dtype=dtype)  # [1,Cin,Cout]
        y = F.conv1d(y, pk[0].t().reshape(co, C, 1).contiguous())
        y = _bn_relu6(y, w, i + 1, dtype)
        acts.append(y)
    xt = y.transpose(1, 2).contiguous()                        # [B,T,C] channels-last
    B, T, C = xt.shape
    d1 = torch.as_tensor(w["dense_1/kernel"], dtype=dtype)
    if a["pool"] == "gap_dense":                                # model.py:759-765
        z = torch.clamp(xt.mean(dim=1) @ d1, 0.0, 6.0)         # GlobalAveragePooling1D -> Dense(256, no bias) -> relu6
        logits = z @ torch.as_tensor(w["dense_2/kernel"], dtype=dtype)
        probs = torch.softmax(logits, dim=-1)
        if return_activations:
            return probs.numpy(), logits.numpy(), [t.transpose(1, 2).contiguous().numpy() for t in acts]
        return probs.numpy()
    att = xt.reshape(B, T * C) @ d1                            # flatten index t*C+c
    if a["dense1_bias"]:
        att = att + torch.as_tensor(w["dense_1/bias"], dtype=dtype)
    att = torch.softmax(att, dim=-1)                           # [B,T]
    weighted = xt * att[:, :, None]
    if a["pool"] == "max_avg":
        z = torch.cat([weighted.max(dim=1).values, xt.mean(dim=1)], dim=1)
    else:
        z = weighted.mean(dim=1)
    logits = z @ torch.as_tensor(w["dense_2/kernel"], dtype=dtype)
    probs = torch.softmax(logits, dim=-1)
    if return_activations:
        return probs.numpy(), logits.numpy(), [t.transpose(1, 2).contiguous().numpy() for t in acts]
    return probs.numpy()


# ---------------------------------------------------------------------------------------------
# steffeNet (model.py:1663-1726): Conv1D(256, 75, strides=50, SAME) on the raw waveform -> BN -> ReLU6 -> one SAME
# depthwise-separable block -> 6 x [residual block (stride 2, 1x1 strided shortcut + BN), residual block (identity
# shortcut)], each residual block = two SAME depthwise-separable blocks + Add -> GlobalMaxPooling || GlobalAveragePooling
# -> Dense(num_classes, softmax, no bias).  Keras numbers the layers in creation order: the shortcut Conv1D / BN of a
# stride-2 residual block are created BEFORE the block's depthwise / pointwise layers.
# ---------------------------------------------------------------------------------------------
STEFFE_WIDTHS = (320, 384, 512, 768, 1024, 1536)


def steffenet_plan():
    """[(kind, ...)] in layer-creation order with the Keras layer numbers each step consumes."""
    plan, conv, bn, dw = [], 1, 1, 0
    plan.append(("conv75", conv, bn)); conv += 1; bn += 1
    dw += 1; plan.append(("dwpw", dw, conv, bn, 1)); conv += 1; bn += 1
    for nh in STEFFE_WIDTHS:
        for stride in (2, 1):
            if stride == 2:
                plan.append(("shortcut", conv, bn)); conv += 1; bn += 1
            else:
                plan.append(("identity",))
            dw += 1; plan.append(("dwpw", dw, conv, bn, stride)); conv += 1; bn += 1
            dw += 1; plan.append(("dwpw", dw, conv, bn, 1)); conv += 1; bn += 1
            plan.append(("add",))
    return plan


def steffenet_weight_shapes(classes: int = 12):
    shapes, c = {}, 256

    def bn(i, ch):
        for nm in ("gamma", "beta", "moving_mean", "moving_variance"):
            shapes[f"batch_normalization_{i}/{nm}"] = (ch,)
    widths = iter(w for w in STEFFE_WIDTHS for _ in range(2))
    nh = c
    for step in steffenet_plan():
        if step[0] == "conv75":
            shapes[f"conv1d_{step[1]}/kernel"] = (75, 1, 256); bn(step[2], 256)
        elif step[0] in ("shortcut", "identity"):
            nh = next(widths)
            if step[0] == "shortcut":
                shapes[f"conv1d_{step[1]}/kernel"] = (1, c, nh); bn(step[2], nh)
        elif step[0] == "dwpw":
            shapes[f"depthwise_conv2d_{step[1]}/depthwise_kernel"] = (1, 3, c, 1)
            shapes[f"conv1d_{step[2]}/kernel"] = (1, c, nh); bn(step[3], nh)
            c = nh
    shapes["dense_1/kernel"] = (2 * c, classes)
    return shapes


def forward_steffenet(x, w, dtype=torch.float64):
    """x [B,16000] -> softmax probabilities [B,classes] (channels-first internally)."""
    x = torch.as_tensor(np.asarray(x), dtype=dtype)

    def bnorm(y, i, relu6=True):
        g, b, m, v = (torch.as_tensor(w[f"batch_normalization_{i}/{n}"], dtype=dtype) for n in ("gamma", "beta", "moving_mean", "moving_variance"))
        sc = torch.rsqrt(v + BN_EPS) * g
        y = y * sc[None, :, None] + (b - m * sc)[None, :, None]
        return torch.clamp(y, 0.0, 6.0) if relu6 else y
    y, res = None, None
    for step in steffenet_plan():
        if step[0] == "conv75":
            k = torch.as_tensor(w[f"conv1d_{step[1]}/kernel"], dtype=dtype)            # [75,1,256]
            out, pl, pr = same_pad(x.shape[1], 75, 50)
            y = F.conv1d(F.pad(x, (pl, pr))[:, None, :], k.permute(2, 1, 0).contiguous(), stride=50)
            y = bnorm(y, step[2])
        elif step[0] == "shortcut":
            k = torch.as_tensor(w[f"conv1d_{step[1]}/kernel"], dtype=dtype)[0]         # [C,nh]
            res = bnorm(F.conv1d(y, k.t()[:, :, None].contiguous(), stride=2), step[2], relu6=False)
        elif step[0] == "identity":
            res = y
        elif step[0] == "dwpw":
            _, dwi, ci, bi, stride = step
            dk = torch.as_tensor(w[f"depthwise_conv2d_{dwi}/depthwise_kernel"], dtype=dtype)
            C = dk.shape[2]
            _, pl, pr = same_pad(y.shape[-1], 3, stride)
            y = F.conv1d(F.pad(y, (pl, pr)), dk[0, :, :, 0].t().reshape(C, 1, 3).contiguous(), stride=stride, groups=C)
            pk = torch.as_tensor(w[f"conv1d_{ci}/kernel"], dtype=dtype)[0]
            y = bnorm(F.conv1d(y, pk.t()[:, :, None].contiguous()), bi)
        else:
            y = y + res
    z = torch.cat([y.max(dim=2).values, y.mean(dim=2)], dim=1)
    return torch.softmax(z @ torch.as_tensor(w["dense_1/kernel"], dtype=dtype), dim=-1).numpy()
