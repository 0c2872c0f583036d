"""Draw contract, CPU side: Philox4x32-10 known answers and the draw mappings."""
import ctypes as C

import numpy as np

from philox_py import make_key, philox4x32_10, threshold

# Known-answer vectors of the Random123 distribution (kat_vectors, philox4x32 10 rounds)
KATS = [
    ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
    ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
    ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
     [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
]


def _oracle_philox(lib, ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib.oracle_philox4x32_10(c, k, o)
    return list(o)


def test_philox_known_answers_python():
    for ctr, key, want in KATS:
        assert philox4x32_10(ctr, key) == want


def test_philox_known_answers_oracle(oracle_lib):
    for ctr, key, want in KATS:
        assert _oracle_philox(oracle_lib, ctr, key) == want


def test_oracle_matches_python_on_random_counters(oracle_lib):
    rng = np.random.default_rng(7)
    for _ in range(200):
        ctr = [int(x) for x in rng.integers(0, 2**32, 4)]
        key = [int(x) for x in rng.integers(0, 2**32, 2)]
        assert _oracle_philox(oracle_lib, ctr, key) == philox4x32_10(ctr, key)


def test_draw_mappings(oracle_lib):
    # bernoulli(p): u < floor(p * 2^32); p >= 1 always, p <= 0 never
    assert oracle_lib.oracle_threshold(0.5) == 1 << 31 == threshold(0.5)
    assert oracle_lib.oracle_threshold(1.0) == 1 << 32
    assert oracle_lib.oracle_threshold(1.5) == 1 << 32
    assert oracle_lib.oracle_threshold(0.0) == 0
    assert oracle_lib.oracle_threshold(-1.0) == 0
    for p in (0.001, 0.6, 0.7, 0.02, 0.0005, 0.256, 0.05, 0.15, 0.03):
        assert oracle_lib.oracle_threshold(p) == threshold(p) == int(np.floor(p * 2.0**32))
    # uniform integer in [0, n): mulhi32(u, n)
    for u, n in ((0, 10), (0xFFFFFFFF, 10), (0x80000000, 3), (123456789, 4096 * 4096), (0xFFFFFFFF, 1)):
        assert oracle_lib.oracle_below(u, n) == (u * n) >> 32 < n


def test_key_derivation():
    seed = 0x1CE27002025
    k = make_key(seed, 5, 0x434F494E)
    assert k == [(0x27002025) ^ 5, 0x1CE ^ 0x434F494E]
