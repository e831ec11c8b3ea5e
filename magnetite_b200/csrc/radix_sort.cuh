// radix_sort.cuh — stable LSD radix sort of (uint64 key, uint32 payload) pairs,
// 8-bit digits.  Stability is what makes the assembly deterministic: pairs with
// equal (row,col) keys keep their emission order (= ascending element index), so
// the segmented reduction adds contributions in the reference's `+=` order
// (reference src/solver.rs:299-323).
//
// Per pass: (1) per-tile digit histogram, (2) exclusive scan of the
// digit-major [256][tiles] table, (3) stable scatter.  The scatter ranks keys
// with warp match_any (no atomics on data), stages the tile in shared memory in
// digit order and writes it out as coalesced runs.
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace mag {

constexpr int kRsThreads = 256;
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsItems = 16;                       // keys per thread
constexpr int kRsTile = kRsThreads * kRsItems;     // 4096 keys per CTA
constexpr int kRadix = 256;

template <class KeyT>
__device__ __forceinline__ uint32_t rs_digit(KeyT key, int shift) {
    return (uint32_t)(key >> shift) & 0xffu;
}

// hist[d * n_tiles + tile] = number of keys of the tile whose digit is d.
template <class KeyT>
__global__ void __launch_bounds__(kRsThreads)
rs_hist_kernel(const KeyT *__restrict__ keys, size_t n, int shift, uint32_t n_tiles,
               uint32_t *__restrict__ hist) {
    __shared__ uint32_t cnt[kRadix];
    cnt[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * kRsTile;
    const int lane = threadIdx.x & 31;
#pragma unroll 4
    for (int i = 0; i < kRsItems; ++i) {
        const size_t idx = base + (size_t)i * kRsThreads + threadIdx.x;
        const bool ok = idx < n;
        const uint32_t d = ok ? rs_digit(keys[idx], shift) : 0xffffffffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        if (ok && lane == (__ffs(peers) - 1)) atomicAdd(&cnt[d], (uint32_t)__popc(peers));
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * n_tiles + blockIdx.x] = cnt[threadIdx.x];
}

// Dynamic shared memory: staged keys (tile*sizeof(KeyT)) then staged payloads (tile*4 B).
template <class KeyT>
__global__ void __launch_bounds__(kRsThreads)
rs_scatter_kernel(const KeyT *__restrict__ kin, const uint32_t *__restrict__ pin,
                  KeyT *__restrict__ kout, uint32_t *__restrict__ pout, size_t n, int shift,
                  uint32_t n_tiles, const uint32_t *__restrict__ offs) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    KeyT *skey = reinterpret_cast<KeyT *>(rs_smem);
    uint32_t *spay = reinterpret_cast<uint32_t *>(rs_smem + (size_t)kRsTile * sizeof(KeyT));
    __shared__ uint32_t warp_cnt[kRsWarps][kRadix];   // running per-warp digit counts
    __shared__ uint32_t digit_base[kRadix];           // first staged slot of each digit
    __shared__ uint32_t global_base[kRadix];          // first output slot of each digit
    __shared__ uint32_t ws[kRsWarps];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int w = 0; w < kRsWarps; ++w) warp_cnt[w][threadIdx.x] = 0;
    __syncthreads();

    const size_t tile_base = (size_t)blockIdx.x * kRsTile;
    const size_t remaining = n - tile_base;
    const uint32_t tile_n = remaining < (size_t)kRsTile ? (uint32_t)remaining : (uint32_t)kRsTile;
    // warp w owns tile items [w*512, (w+1)*512), visited in order: step i covers
    // the 32 consecutive items starting at w*512 + i*32.
    const uint32_t wbase = warp * (kRsItems * 32);

    KeyT key[kRsItems];
    uint32_t pay[kRsItems];
    uint32_t rank[kRsItems];
#pragma unroll
    for (int i = 0; i < kRsItems; ++i) {
        const uint32_t t = wbase + i * 32 + lane;
        const bool ok = t < tile_n;
        key[i] = ok ? kin[tile_base + t] : (KeyT)0;
        pay[i] = ok ? pin[tile_base + t] : 0u;
    }
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < kRsItems; ++i) {
        const uint32_t t = wbase + i * 32 + lane;
        const bool ok = t < tile_n;
        const uint32_t d = ok ? rs_digit(key[i], shift) : 0xffffffffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t before = 0;
        if (ok && lane == leader) {
            before = warp_cnt[warp][d];
            warp_cnt[warp][d] = before + (uint32_t)__popc(peers);
        }
        before = __shfl_sync(0xffffffffu, before, leader);
        rank[i] = before + (uint32_t)__popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();

    // thread d: turn per-warp counts of digit d into exclusive warp offsets
    {
        const uint32_t d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kRsWarps; ++w) {
            const uint32_t c = warp_cnt[w][d];
            warp_cnt[w][d] = run;
            run += c;
        }
        uint32_t tot;
        const uint32_t ex = block_exscan_256(run, ws, tot);
        digit_base[d] = ex;
        global_base[d] = offs[(size_t)d * n_tiles + blockIdx.x];
    }
    __syncthreads();

#pragma unroll
    for (int i = 0; i < kRsItems; ++i) {
        const uint32_t t = wbase + i * 32 + lane;
        if (t < tile_n) {
            const uint32_t d = rs_digit(key[i], shift);
            const uint32_t slot = digit_base[d] + warp_cnt[warp][d] + rank[i];
            skey[slot] = key[i];
            spay[slot] = pay[i];
        }
    }
    __syncthreads();

    for (uint32_t j = threadIdx.x; j < tile_n; j += kRsThreads) {
        const KeyT k = skey[j];
        const uint32_t d = rs_digit(k, shift);
        const size_t g = (size_t)global_base[d] + (j - digit_base[d]);
        kout[g] = k;
        pout[g] = spay[j];
    }
}

// Sorts n pairs by the low `key_bits` bits of the key (uint64 COO keys, or uint32 node ids: a third less
// traffic per pass).  keys/payload hold the input and receive the output; keys_alt/payload_alt are scratch
// of equal size.
template <class KeyT>
inline void radix_sort_pairs(mag_ctx *ctx, KeyT *keys, uint32_t *payload, KeyT *keys_alt,
                             uint32_t *payload_alt, size_t n, int key_bits) {
    if (n < 2 || key_bits <= 0) return;
    if (n > 0xffffffffull) fail(MAG_ERR_BAD_ARG, "radix_sort_pairs: more than 2^32 pairs");
    constexpr size_t kRsSmemBytes = (size_t)kRsTile * (sizeof(KeyT) + sizeof(uint32_t));
    bool &attr_set = sizeof(KeyT) == 8 ? ctx->rs_attr_set : ctx->rs32_attr_set;
    if (!attr_set) {              // a per-device attribute: kept per context, not per process
        MAG_CUDA(cudaFuncSetAttribute(rs_scatter_kernel<KeyT>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kRsSmemBytes));
        attr_set = true;
    }
    const uint32_t n_tiles = cdiv(n, kRsTile);
    const size_t hist_n = (size_t)kRadix * n_tiles;
    DevBuf<uint32_t> hist(ctx, hist_n);
    int passes = (key_bits + 7) / 8;
    KeyT *kin = keys, *kout = keys_alt;
    uint32_t *pin = payload, *pout = payload_alt;
    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        MAG_LAUNCH(ctx, rs_hist_kernel<KeyT>, n_tiles, kRsThreads, 0, (const KeyT *)kin, n, shift, n_tiles, hist.p);
        exclusive_scan_u32(ctx, hist.p, hist_n, hist.p, hist_n);
        MAG_LAUNCH(ctx, rs_scatter_kernel<KeyT>, n_tiles, kRsThreads, kRsSmemBytes, (const KeyT *)kin,
                   (const uint32_t *)pin, kout, pout, n, shift, n_tiles, (const uint32_t *)hist.p);
        std::swap(kin, kout);
        std::swap(pin, pout);
    }
    if (kin != keys) {   // odd number of passes: result sits in the alt buffers
        MAG_CUDA(cudaMemcpyAsync(keys, kin, n * sizeof(KeyT), cudaMemcpyDeviceToDevice, ctx->stream));
        MAG_CUDA(cudaMemcpyAsync(payload, pin, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    }
}

}  // namespace mag
