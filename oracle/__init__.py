"""CPU oracle for the keyword-spotting hot path of see--/speech_recognition.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement (NumPy + torch-CPU)
of the reference's algorithm for the path named in BASELINE.json.  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` may import it, and only as the checker or as the
timed CPU baseline -- never as (part of) the product path.  The product
(``speech_recognition_b200``) never imports ``oracle`` and fails loudly when its
CUDA library is missing.

Parity status (see DESIGN.md "Oracle"):

* Integer / label / byte paths (threshold selection, argmax, majority vote,
  unanimity, 32->12 conversion order): PINNED against the reference's committed
  fixtures (``submit_50_probs.uint8.memmap``, ``submission_*.csv``) through
  ``tests/golden/driver_fixtures.npz``.
* Graph structure, paddings, strides and every front-end constant (stage 1b) and
  the exp-195/206/106 networks (stage 2): PINNED against the reference's own
  serialized TensorFlow ``GraphDef`` (``logs_*/events.out.tfevents.*``), which
  ``tests/golden/make_golden.py`` evaluates node by node with a small NumPy
  GraphDef interpreter (``tests/golden/graphdef_eval.py``) to produce
  ``tests/golden/graph_*.npz``.
* TF 1.4 / Keras 2.1.2 *kernel numerics* (Eigen FFT, cuDNN conv summation
  order) and the trained weights: the libraries and checkpoints are absent
  (README.md:46-47, .MISSING_LARGE_BLOBS) => floating-point parity at the
  TF-kernel level is UNPINNED; the claim is "GPU == this restatement within
  the stated tolerance on identical inputs, weights and pre-drawn parameters".

All citations are file:line relative to /root/reference.
"""

from . import augment, frontend, network, driver  # noqa: F401
