// What does one dependent kernel in a CUDA graph cost on this GPU, as a function of what the kernel does first?
// (the forest step's floor: 8.8 us per replayed step for a 4-CTA grid, 13.6 us cold)   nvcc -arch=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

struct Big { uint64_t w[150]; };   // 1200 bytes of parameters, like ForestConst

__global__ void k_empty() {}
__global__ void k_big(const __grid_constant__ Big b, uint32_t* out) { if (b.w[149] == 12345) out[0] = 1; }
__global__ void k_ld(const uint32_t* in, uint32_t* out) { out[blockIdx.x * blockDim.x + threadIdx.x] = in[blockIdx.x * blockDim.x + threadIdx.x] + 1; }
__global__ void k_chain(const uint32_t* in, uint32_t* out, int depth)
{
    // `depth` dependent global loads before the store (pointer chasing through L2)
    uint32_t i = threadIdx.x & 31;
    for (int d = 0; d < depth; ++d) i = __ldcg(in + i);
    out[blockIdx.x * blockDim.x + threadIdx.x] = i;
}
__global__ void k_atom(uint32_t* out) { if (threadIdx.x == 0) atomicAdd(out + 64, 1u); }
__global__ void k_fence(uint32_t* out) { out[threadIdx.x] = 1; __threadfence(); if (threadIdx.x == 0) atomicAdd(out + 64, 1u); }

template <typename F> static float graph_time(cudaStream_t st, int n, F launch)
{
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < n; ++i) launch();
    cudaStreamEndCapture(st, &g);
    cudaGraphInstantiate(&ge, g, 0);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
    cudaEventRecord(a, st);
    for (int r = 0; r < 10; ++r) cudaGraphLaunch(ge, st);
    cudaEventRecord(b, st); cudaStreamSynchronize(st);
    float ms; cudaEventElapsedTime(&ms, a, b);
    cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
    return ms * 1e3f / (10 * n);
}

int main()
{
    cudaStream_t st; CK(cudaStreamCreate(&st));
    uint32_t *in, *out; CK(cudaMalloc(&in, 1 << 20)); CK(cudaMalloc(&out, 1 << 20));
    CK(cudaMemset(in, 0, 1 << 20)); CK(cudaMemset(out, 0, 1 << 20));
    Big big{}; 
    const int n = 64;
    printf("empty <<<1,32>>>            %.2f us/kernel\n", graph_time(st, n, [&] { k_empty<<<1, 32, 0, st>>>(); }));
    printf("empty <<<1368,128>>>        %.2f\n", graph_time(st, n, [&] { k_empty<<<1368, 128, 0, st>>>(); }));
    printf("1200 B params <<<4,128>>>   %.2f\n", graph_time(st, n, [&] { k_big<<<4, 128, 0, st>>>(big, out); }));
    printf("load+store <<<4,128>>>      %.2f\n", graph_time(st, n, [&] { k_ld<<<4, 128, 0, st>>>(in, out); }));
    for (int d : {1, 2, 4, 8}) printf("chain depth %d <<<4,128>>>   %.2f\n", d, graph_time(st, n, [&] { k_chain<<<4, 128, 0, st>>>(in, out, d); }));
    printf("atomic <<<4,128>>>          %.2f\n", graph_time(st, n, [&] { k_atom<<<4, 128, 0, st>>>(out); }));
    printf("store+fence+atomic          %.2f\n", graph_time(st, n, [&] { k_fence<<<4, 128, 0, st>>>(out); }));
    // plain stream launches (no graph)
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 100; ++i) k_empty<<<1, 32, 0, st>>>();
    cudaStreamSynchronize(st);
    cudaEventRecord(a, st);
    for (int i = 0; i < 1000; ++i) k_ld<<<4, 128, 0, st>>>(in, out);
    cudaEventRecord(b, st); cudaStreamSynchronize(st);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("stream launches load+store  %.2f us/kernel\n", ms);
    // one kernel between two events
    float best = 1e9f;
    for (int r = 0; r < 20; ++r)
    {
        cudaEventRecord(a, st); k_ld<<<4, 128, 0, st>>>(in, out); cudaEventRecord(b, st); cudaStreamSynchronize(st);
        cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    printf("event | load+store | event  %.2f us (best of 20)\n", best * 1e3f);
    return 0;
}
