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
