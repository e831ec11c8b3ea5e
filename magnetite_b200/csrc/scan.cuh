// scan.cuh — device-wide exclusive prefix sum over uint32 (reduce / scan / apply,
// recursive over tile sums).  Used for free-DOF maps, CSR row pointers, segment
// ids of the sorted COO stream and the radix-sort digit offsets.
#pragma once
#include "common.cuh"

namespace mag {

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

// Exclusive scan of one value per thread over a 256-thread block.
__device__ __forceinline__ uint32_t block_exscan_256(uint32_t v, uint32_t *warp_sums /*[8]*/,
                                                     uint32_t &block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    uint32_t woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) {
        uint32_t s = warp_sums[w];
        if (w < warp) woff += s;
        tot += s;
    }
    block_total = tot;
    __syncthreads();   // warp_sums reusable afterwards
    return woff + inc - v;
}

__global__ void __launch_bounds__(kScanThreads)
scan_tile_sums_kernel(const uint32_t *__restrict__ in, size_t n_in, uint32_t *__restrict__ sums) {
    __shared__ uint32_t ws[kScanThreads / 32];
    const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        size_t idx = base + i;
        if (idx < n_in) s += in[idx];
    }
    uint32_t tot;
    (void)block_exscan_256(s, ws, tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// out[i] = tile_offs[tile] + exclusive scan within the tile, for i < n_out;
// inputs at i >= n_in read as 0 (so n_out = n_in + 1 appends the grand total).
__global__ void __launch_bounds__(kScanThreads)
scan_tiles_kernel(const uint32_t *in, size_t n_in, uint32_t *out, size_t n_out,
                  const uint32_t *__restrict__ tile_offs) {
    __shared__ uint32_t ws[kScanThreads / 32];
    const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        size_t idx = base + i;
        v[i] = (idx < n_in) ? in[idx] : 0u;
        s += v[i];
    }
    uint32_t tot;
    uint32_t off = block_exscan_256(s, ws, tot) + (tile_offs ? tile_offs[blockIdx.x] : 0u);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        size_t idx = base + i;
        if (idx < n_out) out[idx] = off;
        off += v[i];
    }
}

// out may alias in.  n_out is n_in or n_in + 1.
inline void exclusive_scan_u32(mag_ctx *ctx, const uint32_t *in, size_t n_in, uint32_t *out,
                               size_t n_out) {
    if (n_out == 0) return;
    const unsigned tiles = cdiv(n_out, kScanTile);
    if (tiles == 1) {
        MAG_LAUNCH(ctx, scan_tiles_kernel, 1, kScanThreads, 0, in, n_in, out, n_out,
                   (const uint32_t *)nullptr);
        return;
    }
    DevBuf<uint32_t> sums(ctx, tiles);
    MAG_LAUNCH(ctx, scan_tile_sums_kernel, tiles, kScanThreads, 0, in, n_in, sums.p);
    exclusive_scan_u32(ctx, sums.p, tiles, sums.p, tiles);
    MAG_LAUNCH(ctx, scan_tiles_kernel, tiles, kScanThreads, 0, in, n_in, out, n_out,
               (const uint32_t *)sums.p);
}

}  // namespace mag
