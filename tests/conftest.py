import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def driver_fixtures():
    with np.load(os.path.join(GOLDEN, "driver_fixtures.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def engine():
    """One libkws handle for the GPU tests.  The product path has no CPU fallback, so this
    fails (not skips) on a box without a B200."""
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from speech_recognition_b200 import Engine
    eng = Engine(device=0, max_rows=512, precision="fp32")
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def synth_small():
    from speech_recognition_b200 import synth
    clips = synth.make_clips(48, seed=synth.SEED + 5)
    bank, offsets = synth.make_noise_bank(seconds=4)
    return dict(clips=clips, bank=bank, offsets=offsets)
