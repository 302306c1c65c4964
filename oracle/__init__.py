"""ORACLE — CPU restatements of the reference's algorithms for the hot path.

Test infrastructure only: nothing under ``efg_b200/`` imports this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may.
Each module cites the reference file:line it restates and says how it is pinned.
"""
