// pfb_devscan.cuh -- device-wide exclusive prefix sum of a uint32 array in three launches (tile sums, scan of the tile
// sums by one CTA, tile scans); out has n + 1 entries (out[n] = the total).  Shared by the outgroup screen and the
// pair search.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace devscan {

constexpr int TILE = 1024;
static __global__ void __launch_bounds__(256) k_tile_sums(const uint32_t *in, uint32_t n, uint32_t *tileSums) {
    __shared__ uint32_t ws[8];
    const uint32_t base = blockIdx.x * TILE;
    uint32_t s = 0;
    for (uint32_t j = threadIdx.x; j < TILE; j += 256) s += (base + j < n) ? in[base + j] : 0u;
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(~0u, s, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int j = 0; j < 8; j++) t += ws[j];
        tileSums[blockIdx.x] = t;
    }
}
static __global__ void __launch_bounds__(1024) k_scan_tiles(uint32_t *tileSums, uint32_t nTiles, uint32_t *total) {
    __shared__ uint32_t ws[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t b = 0; b < nTiles; b += 1024) {
        const uint32_t j = b + threadIdx.x;
        const uint32_t v = j < nTiles ? tileSums[j] : 0u;
        uint32_t x = v;
        for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(~0u, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t z = ws[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(~0u, z, o); if (threadIdx.x >= o) z += y; }
            ws[threadIdx.x] = z;
        }
        __syncthreads();
        const uint32_t before = carry + (threadIdx.x >= 32 ? ws[(threadIdx.x >> 5) - 1] : 0u) + x - v;
        if (j < nTiles) tileSums[j] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
static __global__ void __launch_bounds__(256) k_tile_scan(const uint32_t *in, uint32_t n, const uint32_t *tileSums, uint32_t *out) {
    __shared__ uint32_t ws[8];
    const uint32_t base = blockIdx.x * TILE + threadIdx.x * 4;
    uint32_t v[4], s = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) { v[j] = (base + j < n) ? in[base + j] : 0u; s += v[j]; }
    uint32_t x = s;
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(~0u, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
    __syncthreads();
    uint32_t before = tileSums[blockIdx.x] + x - s;
    for (int j = 0; j < (int)(threadIdx.x >> 5); j++) before += ws[j];
#pragma unroll
    for (int j = 0; j < 4; j++) { if (base + j < n) out[base + j] = before; before += v[j]; }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 255) out[n] = before;
}

inline uint32_t n_tiles(uint32_t n) { return (n + TILE) / TILE; }     // covers n + 1 entries

// in[0 .. n) -> out[0 .. n], tiles = scratch of n_tiles(n) + 1 words, total_dev = one word that receives the total
inline void exscan(const uint32_t *in, uint32_t n, uint32_t *out, uint32_t *tiles, uint32_t *total_dev, cudaStream_t st) {
    const uint32_t nt = n_tiles(n);
    k_tile_sums<<<nt, 256, 0, st>>>(in, n, tiles);
    k_scan_tiles<<<1, 1024, 0, st>>>(tiles, nt, total_dev);
    k_tile_scan<<<nt, 256, 0, st>>>(in, n, tiles, out);
}

}  // namespace devscan
