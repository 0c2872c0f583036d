// traders kernels (examples/traders.rs). See kernels_traders.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.hpp"

namespace incerto
{

constexpr uint32_t kTradersMaxStocks = 256; // portfolio words are held in registers: 32 x 8 nibbles

struct TradersArgs
{
    // traders table
    double* cash;        // Trader.cash
    double* net_worth;   // TraderNetWorth
    uint32_t* portfolio; // Trader.portfolio: word w (stocks 8w..8w+7, 4 bits each = #shares, 0 = not owned) of trader i at [w * n_cap + i]
    const uint8_t* trader_flags; // nullptr: all alive
    const uint32_t* trader_uid;  // nullptr: uid == row
    uint64_t n, n_cap;
    // stocks table
    double* price;        // StockPrice
    uint8_t* stock_flags; // alive + StockDelisted marker bit
    uint32_t num_stocks;
    uint32_t bit_delisted;
    uint32_t step;
    uint32_t do_buy, do_sell, do_net_worth;
    uint32_t buy_t1, sell_t1;   // level-1 byte thresholds: min(256, ceil(256 p))
    uint64_t buy_thr2, sell_thr2; // level-2: floor(256 p / T1 * 2^32), or floor(2^k p * 2^32) with the lottery
    uint32_t buy_kexp, sell_kexp; // level 1 by block lottery with P = 2^-k (k = 5..8); 0: byte thresholds
    uint32_t shares_min, shares_range; // num_shares = shares_min + mulhi(u, shares_range)
    double change_mean, change_std_dev;
    PhiloxRoundKeys rk_buy, rk_sell, rk_price;
};

void launch_traders_step(const TradersArgs& a, cudaStream_t st);
void launch_stocks_price_change(const TradersArgs& a, cudaStream_t st);
void launch_fill_f64(double* p, double v, uint64_t n, cudaStream_t st);
void launch_iota_u64(uint64_t* p, uint64_t n, cudaStream_t st);

} // namespace incerto
