"""TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a checker: a CPU restatement of the reference's algorithm for the
registration hot path (and, in ``oracle/_ref``, the unmodified reference CPU ops behind a C shim).
Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.  The product package ``geotransformer_b200`` never does.
"""
