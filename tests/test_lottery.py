"""Block lottery: the level-1 draw of rare per-entity Bernoulli events, 256 ids per Philox call.
CPU side: the threshold tables of the engine and the oracle against exact rational arithmetic, the oracle against a
pure-Python restatement, and the distribution the scheme has to reproduce (256 independent Bernoulli(2^-8) draws)."""
import ctypes as C
import os
import re
from fractions import Fraction
from math import comb

import numpy as np

from philox_py import philox4x32_10

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def exact_thresholds():
    p, cdf, out = Fraction(1, 256), Fraction(0), []
    for k in range(21):
        cdf += comb(256, k) * p**k * (1 - p) ** (256 - k)
        out.append((cdf * 2**64).__floor__())
    return out


def table_in(path, name):
    src = open(os.path.join(ROOT, path)).read()
    body = src[src.index(name):]
    body = body[body.index("{") + 1:body.index("}")]
    return [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", body)]


def lottery_py(key, block, step):
    """winners of a block, in drawing order (independent restatement of the header comment of lottery.cuh)"""
    first = philox4x32_10([block, 0, step, 0], key)
    u = (first[0] << 32) | first[1]
    k = sum(1 for t in exact_thresholds() if u >= t)
    stream = [(first[w] >> (8 * b)) & 255 for w in (2, 3) for b in range(4)]
    winners, extra = [], 0
    i = 0
    while len(winners) < k:
        if i == len(stream):
            e = philox4x32_10([block, 0, step, 2 + extra], key)
            extra += 1
            stream += [(e[w] >> (8 * b)) & 255 for w in range(4) for b in range(4)]
        b = stream[i]
        i += 1
        if b not in winners:
            winners.append(b)
    return winners


def oracle_lottery(lib, key, block, step):
    out = (C.c_uint8 * 32)()
    n = lib.oracle_lottery256(key[0], key[1], block, step, out)
    return list(out[:n])


def test_threshold_tables_are_the_exact_binomial_cdf():
    want = exact_thresholds()
    assert want[-1] == 2**64 - 1 and want[-2] < want[-1]  # 21 entries carry every bit of the 64-bit uniform
    assert table_in("incerto_b200/csrc/lottery.cuh", "kLotteryCdf[21]") == want
    assert table_in("oracle/csrc/lottery.hpp", "LOTTERY_CDF[21]") == want
    # the four immediates of lottery_count
    src = open(os.path.join(ROOT, "incerto_b200/csrc/lottery.cuh")).read()
    imm = re.findall(r"u >= (0x[0-9a-f]+)ull \? 1u", src)
    assert [int(x, 16) for x in imm] == want[:4]


def test_oracle_matches_python_restatement(oracle_lib):
    rng = np.random.default_rng(11)
    many = 0
    for _ in range(3000):
        key = [int(x) for x in rng.integers(0, 2**32, 2)]
        block, step = int(rng.integers(0, 2**24)), int(rng.integers(0, 2**20))
        w = oracle_lottery(oracle_lib, key, block, step)
        assert w == lottery_py(key, block, step)
        assert len(set(w)) == len(w)
        many += len(w) >= 4
    assert many > 20  # the table part of the count (K >= 4) was exercised


def test_distribution_is_independent_bernoulli(oracle_lib):
    """winner counts follow Binomial(256, 1/256), every position wins with 2^-8, pairs of positions are independent"""
    key = [0x27002025, 0x1CE ^ 0x46524547]
    nblocks = 200_000
    counts = np.zeros(24, np.int64)
    per_pos = np.zeros(256, np.int64)
    pair = 0  # positions 3 and 77 both win
    for b in range(nblocks):
        w = oracle_lottery(oracle_lib, key, b, 17)
        counts[len(w)] += 1
        per_pos[w] += 1
        pair += (3 in w) and (77 in w)
    pmf = [comb(256, k) * (1 / 256) ** k * (255 / 256) ** (256 - k) for k in range(24)]
    for k in range(6):
        exp = nblocks * pmf[k]
        assert abs(counts[k] - exp) < 5 * np.sqrt(exp) + 1, (k, counts[k], exp)
    exp = nblocks / 256
    assert np.all(np.abs(per_pos - exp) < 5.5 * np.sqrt(exp))
    assert abs(per_pos.sum() / nblocks - 1.0) < 0.01
    assert pair < 20  # expectation 200000 / 65536 = 3
