"""Third, pure-Python statement of Philox4x32-10 used to cross-check the oracle (tests only)."""
M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return [c0, c1, c2, c3]


def make_key(seed, replica, tag):
    return [(seed & MASK) ^ replica, ((seed >> 32) & MASK) ^ tag]


def threshold(p):
    if not p > 0.0:
        return 0
    if p >= 1.0:
        return 1 << 32
    return int(p * 4294967296.0)
