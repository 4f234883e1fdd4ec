"""CPU oracle for the HuPR hot path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU restatement (numpy fp64 / torch-CPU fp32)
of the reference's algorithm for the hot path, used solely as the *checker*:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it.  The product package
(``hupr_b200``) never imports from here and has no CPU fallback.

Parity pinning: the reference repo ships no tests / golden vectors
(SURVEY.md §4), so the oracle is pinned against outputs of the reference's own
files executed in the build container (``oracle/ref_shim.py`` +
``oracle/make_golden.py``); the resulting fixtures live in ``tests/golden/``.
"""
