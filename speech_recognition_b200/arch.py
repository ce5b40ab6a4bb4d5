"""Architecture tables of the shipped raw-waveform networks.

exp 195 / 206 = ``conv_1d_time_sliced_with_attention_model(filter_mult=1,
num_classes=12)`` (reference model.py:775-838); exp 106 = the older 32-class
variant recovered from the reference's ``logs_106`` GraphDef.
Variable names are the Keras names found in those graphs, which is also how a
Keras 2.1.2 HDF5 checkpoint / frozen .pb names them.
"""
from __future__ import annotations

INPUT_SAMPLES = 16000
PATCH, PATCH_STRIDE = 40, 20          # model.py:805  overlapping_time_slice_stack(x, 40, 20)
BN_EPS = 1e-3                         # Keras BatchNormalization default

ARCHS = {
    195: dict(conv1=128, blocks=[(128, 1), (192, 2), (192, 1), (256, 2), (256, 1), (320, 2),
                                 (320, 1), (384, 2), (384, 1), (512, 2), (512, 1)],
              dense1_bias=True, pool="max_avg", classes=12),
    106: dict(conv1=64, blocks=[(128, 1), (192, 2), (192, 1), (256, 2), (256, 1), (320, 2),
                                (320, 1), (384, 2), (384, 1), (448, 2), (448, 1)],
              dense1_bias=False, pool="attn_mean", classes=32),
}
ARCHS[206] = ARCHS[195]
# conv_1d_time_sliced_model(filter_mult=1) (reference model.py:716-772): conv1d_1 with 32 filters, 13 blocks, head =
# GlobalAveragePooling1D -> Dense(256, no bias) -> ReLU6 -> Dense(num_classes, softmax, no bias).  No checkpoint of it is
# shipped; num_classes is an argument of the builder (12 here, the competition's label set).
TIME_SLICED = 716
ARCHS[TIME_SLICED] = dict(conv1=32, blocks=[(64, 1), (128, 2), (128, 1), (192, 2), (192, 1), (256, 2), (256, 1), (320, 2),
                                            (320, 1), (384, 2), (384, 1), (512, 2), (512, 1)],
                          dense1_bias=False, pool="gap_dense", hidden=256, classes=12)


# steffeNet (reference model.py:1663-1726): Conv1D(256, 75, strides=50) stem, one SAME depthwise-separable block, then per
# width two residual blocks (the first with stride 2 and a strided 1x1 shortcut + BN), max || average pooling, Dense.
STEFFENET = 1663
STEFFE_WIDTHS = (320, 384, 512, 768, 1024, 1536)
ARCHS[STEFFENET] = dict(conv1=256, blocks=[], dense1_bias=False, pool="max_avg_dense", classes=12)


def steffenet_weight_shapes(classes: int = 12):
    """Keras variable names / shapes in layer-creation order: the shortcut Conv1D / BN of a stride-2 residual block are
    created before the block's depthwise / pointwise layers."""
    shapes, c = {}, 256
    conv, bn, dw = 1, 1, 0

    def add_bn(ch):
        nonlocal bn
        for nm in ("gamma", "beta", "moving_mean", "moving_variance"):
            shapes[f"batch_normalization_{bn}/{nm}"] = (ch,)
        bn += 1

    def add_dwpw(cout):
        nonlocal conv, dw, c
        dw += 1
        shapes[f"depthwise_conv2d_{dw}/depthwise_kernel"] = (1, 3, c, 1)
        shapes[f"conv1d_{conv}/kernel"] = (1, c, cout); conv += 1
        add_bn(cout)
        c = cout
    shapes["conv1d_1/kernel"] = (75, 1, 256); conv += 1
    add_bn(256)
    add_dwpw(256)
    for nh in STEFFE_WIDTHS:
        for stride in (2, 1):
            if stride == 2:
                shapes[f"conv1d_{conv}/kernel"] = (1, c, nh); conv += 1
                add_bn(nh)
            add_dwpw(nh)
            add_dwpw(nh)
    shapes["dense_1/kernel"] = (2 * c, classes)
    return shapes


def same_pad(T: int, k: int, s: int):
    """TF 'SAME': out = ceil(T/s), pad_left = pad_total // 2."""
    out = -(-T // s)
    total = max((out - 1) * s + k - T, 0)
    return out, total // 2, total - total // 2


def layer_lengths(arch: int, input_size: int = INPUT_SAMPLES):
    """[n_patches, T after conv1d_1, T after each block]."""
    n_patch, _, _ = same_pad(input_size, PATCH, PATCH_STRIDE)
    T = (n_patch - 3) // 2 + 1
    out = [n_patch, T]
    for _, s in ARCHS[arch]["blocks"]:
        T = T - 2 if s == 1 else same_pad(T, 3, 2)[0]
        out.append(T)
    return out


def weight_shapes(arch: int):
    if arch == STEFFENET:
        return steffenet_weight_shapes(ARCHS[arch]["classes"])
    a = ARCHS[arch]
    shapes = {}
    c = a["conv1"]
    shapes["conv1d_1/kernel"] = (3, PATCH, c)

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
    shapes["dense_2/kernel"] = ((2 * c if a["pool"] == "max_avg" else c), a["classes"])
    return shapes
