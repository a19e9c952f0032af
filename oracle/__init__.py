"""CPU oracle for the homography-adaptation / extraction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker (or as the
timed CPU arm), never as the thing shipped.  The product path
(``pytorch_superpoint_b200``) never imports this package and raises if its CUDA
library is missing.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified
reference from ``/root/reference`` (build container only) and writes small
input/output fixtures to ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
replays them against this restatement, together with the reference's own
known-answer artefact for the valid mask (``notebooks/h2.npz``) and the
descriptor-matching artefacts.
"""
from .oracle import *  # noqa: F401,F403
