"""CPU oracle for the GAViKO 3D-ViT hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``gaviko_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there
only as the checker / the timed CPU arm, never as the product path.

Parity status: the reference ships no tests or golden vectors,
so the restatement in ``gaviko_oracle.py`` is pinned against outputs of the
reference itself, run in the build container by ``oracle/make_golden.py`` and
committed under ``tests/golden/``.
"""
