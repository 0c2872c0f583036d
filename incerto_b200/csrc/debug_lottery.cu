// Test hook (not part of the ABI of include/incerto_b200.h): the device side of the block lottery (lottery.cuh) for a
// range of blocks, so that tests/test_gpu_lottery.py can compare it with the oracle's winner lists directly - the
// kernels only show it through their results.
#include <cuda_runtime.h>

#include "lottery.cuh"

namespace incerto
{
namespace
{
constexpr int kMaxOut = 64;
__global__ void debug_lottery_kernel(PhiloxRoundKeys rk, uint32_t kexp, uint32_t first_block, uint32_t sub, uint32_t step, uint32_t nblocks,
                                     uint8_t* out /* nblocks x (1 + kMaxOut - 1): count, then the winners in drawing order */)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nblocks) return;
    uint8_t* o = out + static_cast<size_t>(i) * kMaxOut;
    uint32_t n = 0;
    auto emit = [&](uint32_t b) {
        if (n + 1 < kMaxOut) o[1 + n] = static_cast<uint8_t>(b);
        ++n;
    };
    if (kexp == 8)
        lottery256(rk, first_block + i, sub, step, emit);
    else
        lottery256_k(rk, kexp, first_block + i, sub, step, emit);
    o[0] = static_cast<uint8_t>(n);
}
} // namespace
} // namespace incerto

extern "C" int incerto_debug_lottery(uint32_t k0, uint32_t k1, uint32_t kexp, uint32_t first_block, uint32_t sub, uint32_t step,
                                     uint32_t nblocks, uint8_t* host_out /* nblocks x 64 */)
{
    using namespace incerto;
    if (kexp < 5 || kexp > 8 || !nblocks) return 2;
    uint8_t* d = nullptr;
    const size_t bytes = static_cast<size_t>(nblocks) * kMaxOut;
    if (cudaMalloc(&d, bytes) != cudaSuccess) return 30;
    PhiloxKey key;
    key.k0 = k0;
    key.k1 = k1;
    debug_lottery_kernel<<<(nblocks + 127) / 128, 128>>>(make_round_keys(key), kexp, first_block, sub, step, nblocks, d);
    cudaError_t e = cudaMemcpy(host_out, d, bytes, cudaMemcpyDeviceToHost);
    cudaFree(d);
    return e == cudaSuccess ? 0 : 30;
}
