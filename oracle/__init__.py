"""CPU oracle for the dynamic-fixed-point (DFXP) training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``lbt_b200/`` may import this package; only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` do, and there only as the checker / the CPU baseline, never as the product path.

PARITY UNPINNED: the reference (freudh/lbt) ships no tests, no golden vectors and cannot run in
this image (TensorFlow 1.x is absent).  The arithmetic lives in TensorFlow (unpinned, >=1.7,<2.0);
this package restates the documented TF1 semantics of every call the reference makes on the hot
path.  The pins are the hand-derived known-answer vectors of SURVEY.md App. A.4
(``tests/golden/kat_quantizer.json``) which ``tests/test_oracle_kat.py`` reproduces.
"""
