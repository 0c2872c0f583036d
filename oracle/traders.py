"""ORACLE — test infrastructure, not product code.  numpy wrapper over the C++ traders oracle."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .lib import load

_F64P = C.POINTER(C.c_double)
_U8P = C.POINTER(C.c_uint8)


class TradersOracle:
    def __init__(self, seed, replica, num_traders=10, num_stocks=100, cash=100.0, p_buy=0.01, p_sell=0.005,
                 shares=(1, 5), price=5.0, mean=0.001, std_dev=0.2, systems=15):
        self.lib = load()
        self.nt, self.ns = num_traders, num_stocks
        self.h = self.lib.oracle_traders_new(seed, replica, num_traders, num_stocks, cash, p_buy, p_sell, shares[0],
                                             shares[1], price, mean, std_dev, systems)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.oracle_traders_free(self.h)
            self.h = None

    def set_prices(self, prices):
        p = np.ascontiguousarray(prices, dtype=np.float64)
        assert len(p) == self.ns
        self.lib.oracle_traders_set_prices(self.h, p.ctypes.data_as(_F64P))
        return self

    def run(self, steps):
        self.lib.oracle_traders_run(self.h, steps)
        return self

    def state(self):
        cash = np.zeros(self.nt)
        nw = np.zeros(self.nt)
        shares = np.zeros((self.nt, self.ns), np.uint8)
        price = np.zeros(self.ns)
        delisted = np.zeros(self.ns, np.uint8)
        self.lib.oracle_traders_state(self.h, cash.ctypes.data_as(_F64P), nw.ctypes.data_as(_F64P),
                                      shares.ctypes.data_as(_U8P), price.ctypes.data_as(_F64P),
                                      delisted.ctypes.data_as(_U8P))
        return dict(cash=cash, net_worth=nw, shares=shares, price=price, delisted=delisted.astype(bool))

    def series(self, trader):
        n = self.lib.oracle_traders_series(self.h, trader, None, 0)
        out = np.zeros(n)
        self.lib.oracle_traders_series(self.h, trader, out.ctypes.data_as(_F64P), n)
        return out


def det_normal(words):
    lib = load()
    return lib.oracle_det_normal(*[int(w) for w in words])
