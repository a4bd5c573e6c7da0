// ubench.cu -- B200 micro-benchmarks that size the design choices of pfb200.cu:
// random L2 atomics (sketch update), random L2 sector reads (bitmap probe), shared-memory
// atomics / reads (staged alternatives).  Prints G ops/s.  Build: nvcc -arch=sm_100a -O3.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ uint32_t rnd(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x;
}

template <int MODE>  // 0 red.or, 1 atom.or + conditional second, 2 ld, 3 ld + dependent compare of a second table
__global__ void k_l2(uint32_t *tab, uint32_t nwords, uint32_t iters, uint32_t *sink) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (uint32_t i = 0; i < iters; i++) {
        uint32_t r = rnd(t * 0x9E3779B9u + i * 0x85EBCA6Bu + 1u);
        uint32_t w = (uint32_t)(((uint64_t)r * nwords) >> 32);
        uint32_t bit = 1u << (r & 15);
        if (MODE == 0) atomicOr(&tab[w], bit);
        if (MODE == 1) { uint32_t old = atomicOr(&tab[w], bit); if (old & bit) atomicOr(&tab[w], bit << 16); }
        if (MODE == 2) acc += __ldg(&tab[w]) & bit;
    }
    if (acc == 0xFFFFFFFFu) sink[0] = acc;
}

template <int MODE>  // 0 smem atomicOr, 1 smem read
__global__ void k_smem(uint32_t iters, uint32_t *sink) {
    extern __shared__ uint32_t sm[];
    const uint32_t nwords = 40960;  // 160 KB
    for (uint32_t i = threadIdx.x; i < nwords; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (uint32_t i = 0; i < iters; i++) {
        uint32_t r = rnd(t * 0x9E3779B9u + i * 0x85EBCA6Bu + 1u);
        uint32_t w = (uint32_t)(((uint64_t)r * nwords) >> 32);
        uint32_t bit = 1u << (r & 31);
        if (MODE == 0) atomicOr(&sm[w], bit);
        if (MODE == 1) acc += sm[w] & bit;
    }
    __syncthreads();
    if (acc == 0xFFFFFFFFu || sm[threadIdx.x] == 0x12345678u) sink[0] = acc;
}

template <typename F>
float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
    uint32_t *tab, *sink;
    size_t maxb = 512u << 20;
    cudaMalloc(&tab, maxb); cudaMalloc(&sink, 64);
    cudaMemset(tab, 0, maxb);
    const uint32_t iters = 64;
    const int blocks = 148 * 32, threads = 256;
    double nops = (double)blocks * threads * iters;
    size_t sizes[] = {1250000, 2500000, 12500000, 25000000, 100000000, 400000000};
    for (size_t sz : sizes) {
        uint32_t nw = (uint32_t)(sz / 4);
        float m0 = timeit([&] { k_l2<0><<<blocks, threads>>>(tab, nw, iters, sink); });
        cudaMemset(tab, 0, sz);
        float m1 = timeit([&] { k_l2<1><<<blocks, threads>>>(tab, nw, iters, sink); });
        float m2 = timeit([&] { k_l2<2><<<blocks, threads>>>(tab, nw, iters, sink); });
        printf("table %9zu B: red.or %7.1f G/s   atom.or(+cond) %7.1f G/s   ld.32 %7.1f G/s\n", sz,
               nops / m0 * 1e-6, nops / m1 * 1e-6, nops / m2 * 1e-6);
    }
    cudaFuncSetAttribute(k_smem<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 163840);
    cudaFuncSetAttribute(k_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 163840);
    const uint32_t it2 = 1024;
    double nops2 = 148.0 * 1024 * it2;
    float s0 = timeit([&] { k_smem<0><<<148, 1024, 163840>>>(it2, sink); });
    float s1 = timeit([&] { k_smem<1><<<148, 1024, 163840>>>(it2, sink); });
    printf("smem 160 KB/CTA, 148 CTAs x 1024 thr: atomicOr %7.1f G/s   ld %7.1f G/s\n", nops2 / s0 * 1e-6, nops2 / s1 * 1e-6);
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
