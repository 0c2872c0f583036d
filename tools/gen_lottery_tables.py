"""Threshold tables of the block lottery: floor(P(Binomial(256, 2^-k) <= j) * 2^64) for j = 0, 1, ...
until the value reaches 2^64 - 1, for k = 5..8, in exact rational arithmetic.  The engine (incerto_b200/csrc/lottery.cuh)
and the oracle (oracle/csrc/lottery.hpp) each carry a copy; tests/test_lottery.py recomputes and compares both.
    python tools/gen_lottery_tables.py          prints the C initialisers"""
from fractions import Fraction
from math import comb


def table(k, n=256):
    p, cdf, out = Fraction(1, 1 << k), Fraction(0), []
    for j in range(n + 1):
        cdf += comb(n, j) * p**j * (1 - p) ** (n - j)
        out.append(min((cdf * 2**64).__floor__(), 2**64 - 1))
        if out[-1] == 2**64 - 1:
            break
    return out


if __name__ == "__main__":
    for k in (5, 6, 7, 8):
        t = table(k)
        print(f"// k = {k}: Binomial(256, 2^-{k}), {len(t)} thresholds")
        body = ", ".join(f"0x{v:016x}ull" for v in t)
        print("{" + body + "},")
