"""CPU oracle for the SatNeRF / Semantic-NeRF ray-rendering hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker (or as the CPU
arm that is timed *beside* the CUDA path) - never as the thing shipped.

Parity status: PINNED.  ``oracle/pin_against_reference.py`` (run in the build
container, where ``/root/reference`` is importable) checks every function here
against the reference's own PyTorch implementation and freezes the golden
vectors under ``tests/golden/``; ``tests/test_oracle_golden.py`` re-checks the
oracle against those vectors on any machine.
"""
from .render_oracle import *  # noqa: F401,F403
