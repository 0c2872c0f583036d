"""Block lottery: the level-1 draw of rare per-entity Bernoulli events, 256 ids per Philox call.
CPU side: the threshold tables of the engine and the oracle against exact rational arithmetic, the oracle against a
pure-Python restatement, and the distribution the scheme has to reproduce (256 independent Bernoulli(2^-k) draws)."""
import ctypes as C
import os
import re
import sys
from math import comb

import numpy as np
import pytest

from philox_py import philox4x32_10

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gen_lottery_tables import table  # noqa: E402  (exact rational arithmetic)

TABLES = {k: table(k) for k in (5, 6, 7, 8)}


def hex_lists(path, start):
    src = open(os.path.join(ROOT, path)).read()
    src = src[src.index(start):]
    return [[int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", body)] for body in re.findall(r"\{([^{}]*)\}", src)]


def lottery_py(key, block, sub, step, kexp=8):
    """winners of a block, in drawing order (independent restatement of the header comment of lottery.cuh)"""
    first = philox4x32_10([block, sub, step, 0], key)
    u = (first[0] << 32) | first[1]
    k = sum(1 for t in TABLES[kexp] if u >= t)
    stream = [(first[w] >> (8 * b)) & 255 for w in (2, 3) for b in range(4)]
    winners, extra = [], 0
    i = 0
    while len(winners) < k:
        if i == len(stream):
            e = philox4x32_10([block, sub, step, 2 + extra], key)
            extra += 1
            stream += [(e[w] >> (8 * b)) & 255 for w in range(4) for b in range(4)]
        b = stream[i]
        i += 1
        if b not in winners:
            winners.append(b)
    return winners


def oracle_lottery(lib, key, block, sub, step, kexp=8):
    out = (C.c_uint8 * 64)()
    n = lib.oracle_lottery256(key[0], key[1], block, sub, step, kexp, out)
    return list(out[:n])


def test_threshold_tables_are_the_exact_binomial_cdf():
    for k, t in TABLES.items():
        assert t[-1] == 2**64 - 1 and t[-2] < t[-1]  # the table carries every bit of the 64-bit uniform
    assert [len(TABLES[k]) for k in (5, 6, 7, 8)] == [45, 33, 26, 21]
    # engine: one padded row per k
    rows = hex_lists("incerto_b200/csrc/lottery.cuh", "kLotteryCdf[4][kLotteryRows]")[:4]
    for k, row in zip((5, 6, 7, 8), rows):
        assert len(row) == 45
        assert row[:len(TABLES[k])] == TABLES[k] and all(v == 2**64 - 1 for v in row[len(TABLES[k]):])
    # oracle: one array per k
    for k in (5, 6, 7, 8):
        assert hex_lists("oracle/csrc/lottery.hpp", f"LOTTERY_CDF_K{k}[")[0] == TABLES[k]
    # the four immediates of lottery_count (k = 8)
    src = open(os.path.join(ROOT, "incerto_b200/csrc/lottery.cuh")).read()
    imm = re.findall(r"u >= (0x[0-9a-f]+)ull \? 1u", src)
    assert [int(x, 16) for x in imm] == TABLES[8][:4]


@pytest.mark.parametrize("kexp", [5, 6, 7, 8])
def test_oracle_matches_python_restatement(oracle_lib, kexp):
    rng = np.random.default_rng(11 + kexp)
    many = 0
    for _ in range(1500):
        key = [int(x) for x in rng.integers(0, 2**32, 2)]
        block, sub, step = int(rng.integers(0, 2**24)), int(rng.integers(0, 3)), int(rng.integers(0, 2**20))
        w = oracle_lottery(oracle_lib, key, block, sub, step, kexp)
        assert w == lottery_py(key, block, sub, step, kexp)
        assert len(set(w)) == len(w)
        many += len(w) > 8
    assert kexp > 6 or many > 5  # the continuation calls of the byte stream were exercised


@pytest.mark.parametrize("kexp", [6, 8])
def test_distribution_is_independent_bernoulli(oracle_lib, kexp):
    """winner counts follow Binomial(256, 2^-k), every position wins with 2^-k, pairs of positions are independent"""
    key = [0x27002025, 0x1CE ^ 0x46524547]
    nblocks = 120_000
    p = 2.0 ** -kexp
    counts = np.zeros(64, np.int64)
    per_pos = np.zeros(256, np.int64)
    pair = 0  # positions 3 and 77 both win
    for b in range(nblocks):
        w = oracle_lottery(oracle_lib, key, b, 1, 17, kexp)
        counts[len(w)] += 1
        per_pos[w] += 1
        pair += (3 in w) and (77 in w)
    pmf = [comb(256, k) * p**k * (1 - p) ** (256 - k) for k in range(12)]
    for k in range(6):
        exp = nblocks * pmf[k]
        assert abs(counts[k] - exp) < 5 * np.sqrt(exp) + 1, (k, counts[k], exp)
    exp = nblocks * p
    assert np.all(np.abs(per_pos - exp) < 5.5 * np.sqrt(exp))
    assert abs(per_pos.sum() / (nblocks * 256 * p) - 1.0) < 0.01
    exp_pair = nblocks * p * p
    assert abs(pair - exp_pair) < 6 * np.sqrt(exp_pair) + 3
