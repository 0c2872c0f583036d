"""ORACLE — test infrastructure, not product code.

CPU restatement of the reference's algorithm for the simulation hot path (C++17 under
oracle/csrc, built into oracle/_build/liboracle.so) plus a pure-Python restatement of the
incerto library semantics (oracle/incerto_ref.py).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this package.

Parity status: the reference is Rust on Bevy and cannot be built here (no cargo/rustc); every
RNG-dependent result of the reference is unpinned by its own tests, so for those
paths this oracle is "parity unpinned" and is anchored on the draw contract of DESIGN.md §2.
The RNG-free behaviour (step/series cadence, aggregators, builder conflicts, neighbour sets,
bounds) IS pinned: tests/test_oracle_kats.py checks the oracle against every known-answer
test the reference's tests/ directory holds.
"""
from .lib import load, build  # noqa: F401
