"""CPU oracle for the profile-HMM hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under `protpore_b200/` may import this package; only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s CPU-baseline / `--impl reference` legs do.  See oracle/README.md.
"""
