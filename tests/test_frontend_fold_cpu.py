"""The radix-2 step the tensor-core front end takes out of its DFT-as-GEMM (csrc/tc_frontend.cu): even and odd bins of the
512-point transform of a windowed frame are K = 256 contractions of the sum / difference of the two window halves.
Checked here in NumPy against np.fft.rfft on the oracle's own window, for both window lengths the GPU tests use."""
import numpy as np
import pytest

from oracle import frontend


@pytest.mark.parametrize("win", [480, 400])
def test_even_odd_fold_equals_rfft(win):
    rng = np.random.RandomState(7)
    x = rng.randn(5, win)
    w = frontend.hann_window_periodic(win).astype(np.float64)
    y = np.zeros((5, 512))
    y[:, :win] = x * w
    ref = np.fft.rfft(y, 512)[:, :256]
    n = np.arange(256)
    s, d = y[:, :256] + y[:, 256:], y[:, :256] - y[:, 256:]
    j = np.arange(128)
    even = s @ np.exp(-2j * np.pi * np.outer(n, 2 * j) / 512)
    odd = d @ np.exp(-2j * np.pi * np.outer(n, 2 * j + 1) / 512)
    got = np.empty_like(ref)
    got[:, 0::2], got[:, 1::2] = even, odd
    assert np.abs(got - ref).max() < 1e-9 * np.abs(ref).max()


def test_mel_row_division_magic():
    """idx / n_mel as __umulhi(idx, ceil(2^32 / n_mel)) for every index of a 128-row tile (the coalesced log-mel store)."""
    for d in range(2, 129):
        m = (2 ** 32 + d - 1) // d
        assert m < 2 ** 32
        idx = np.arange(128 * d, dtype=np.uint64)
        assert np.array_equal((idx * np.uint64(m)) >> np.uint64(32), idx // np.uint64(d)), d


def _split16(a):
    """fp32 -> (hi, lo) fp16 parts as the kernel's producers / the host basis builder make them."""
    a = a.astype(np.float32)
    hi = a.astype(np.float16).astype(np.float32)
    lo = (a - hi).astype(np.float16).astype(np.float32)
    return hi, lo


@pytest.mark.parametrize("win,hop,n_mel", [(480, 160, 40), (400, 240, 80)])
def test_split_fp16_fold_reaches_the_fp32_tier(win, hop, n_mel):
    """NumPy emulation of the tensor-core front end's arithmetic (window in fp32, sum / difference of the window halves in
    fp32, hi / lo fp16 split of both operands, the three products a_hi b_hi + a_lo b_hi + a_hi b_lo) against the oracle:
    the scheme itself must sit inside the 1e-4 bar the GPU test holds the kernel to (measured here: 1.1e-5 for 480 / 160 / 40,
    3.6e-5 for 400 / 240 / 80, where more filters are buried near the 1e-6 floor inside the log)."""
    from speech_recognition_b200 import synth
    x = synth.make_clips(3, seed=78)
    w = frontend.hann_window_periodic(win)
    y = np.zeros(x.shape[:-1] + (frontend.frame(x, win, hop).shape[-2], 512), np.float32)
    y[..., :win] = (frontend.frame(x, win, hop) * w).astype(np.float32)
    s = (y[..., :256] + y[..., 256:]).astype(np.float32)
    d = (y[..., :256] - y[..., 256:]).astype(np.float32)
    n = np.arange(256)[:, None]
    spec = np.empty(y.shape[:-1] + (256,), np.float32)
    for odd, a in ((0, s), (1, d)):
        bins = 2 * np.arange(128)[None, :] + odd
        ang = 2.0 * np.pi * ((n * bins) % 512) / 512.0
        a_hi, a_lo = _split16(a)
        acc = []
        for basis in (np.cos(ang), -np.sin(ang)):
            b_hi, b_lo = _split16(basis.astype(np.float32))
            acc.append((a_hi.astype(np.float64) @ b_hi + a_lo.astype(np.float64) @ b_hi +
                        a_hi.astype(np.float64) @ b_lo).astype(np.float32))
        spec[..., odd::2] = np.sqrt(acc[0] * acc[0] + acc[1] * acc[1])
    mel_w = frontend.linear_to_mel_weight_matrix(n_mel, 257)
    assert not mel_w[256].any()                                  # the Nyquist bin carries no weight: 256 bins suffice
    got = frontend.log_mel(spec, mel_w[:256])
    ref = frontend.features(x, window_size_samples=win, window_stride_samples=hop, dct_coefficient_count=n_mel, kind="logmel")
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err < 5e-5, err
