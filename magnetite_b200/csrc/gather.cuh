// gather.cuh — gather assembly (mag_options.assembly = 0, the default): the full K as 2x2-block CSR without sorting
// COO keys and without materialising K_e.  Same output, bit for bit, as the sort-and-reduce path of
// assembly.cuh (reference src/solver.rs:290-331); see gather_core.h for the per-node algorithm and why
// the accumulation order is the reference's.
//
//   emit    3 (node, incidence) pairs per triangle, incidence = local_element*3 + corner; pairs whose node
//           another rank owns get the sentinel key; per-node counts by integer atomics (deterministic)
//   sort    the stable LSD radix sort of radix_sort.cuh by the log2(N)-bit node key: 3E pairs x 3 passes at
//           8 M nodes, against 9E pairs x 6 passes for the COO keys
//   count   one thread per owned node: distinct column nodes of its row -> browptr (scan)
//   fill    one thread per owned node: recompute the two K_e rows of every incident (element, corner) and
//           add the blocks column by column in ascending (element, corner) order -> bcol, bval
//
// HBM traffic per triangle (plate): 12 B connectivity + 3 x 12 B pairs per sort pass + ~10 x 12 B of
// connectivity re-reads and 16-byte coordinate gathers that mostly hit L1/L2 + 7/2 blocks x 36 B out,
// against 288 B of K_e written and read back plus 9 x 12 B pairs per pass on the sorted-key path.
#pragma once
#include "assembly.cuh"
#include "common.cuh"
#include "element.cuh"
#include "gather_core.h"
#include "radix_sort.cuh"
#include "scan.cuh"

namespace mag {

constexpr int kGatherThreads = 128;

// One thread per local element.  cnt[node - node_lo] counts the incidences of every owned node.
__global__ void emit_incidence_kernel(const uint32_t *__restrict__ n0, const uint32_t *__restrict__ n1,
                                      const uint32_t *__restrict__ n2, const uint32_t *__restrict__ elist,
                                      size_t n_local, uint32_t node_lo, uint32_t node_hi,
                                      uint32_t *__restrict__ keys, uint32_t *__restrict__ payload,
                                      uint32_t *__restrict__ cnt) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_local) return;
    const gather::Conn conn{n0, n1, n2, elist};
    uint32_t nd[3];
    gather::corner_nodes(conn, (uint32_t)i, nd);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const bool mine = nd[k] >= node_lo && nd[k] < node_hi;
        keys[i * 3 + k] = mine ? nd[k] : 0xffffffffu;       // sentinel: sorts behind every node id
        payload[i * 3 + k] = (uint32_t)(i * 3 + k);
        if (mine) atomicAdd(&cnt[nd[k] - node_lo], 1u);          // integer: deterministic
    }
}

// nblk[r] = number of distinct column nodes of owned node row r.
__global__ void __launch_bounds__(kGatherThreads)
gather_count_kernel(const uint32_t *__restrict__ n0, const uint32_t *__restrict__ n1,
                    const uint32_t *__restrict__ n2, const uint32_t *__restrict__ elist,
                    const uint32_t *__restrict__ payload, const uint32_t *__restrict__ nptr, uint32_t n_own,
                    uint32_t *__restrict__ nblk) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_own) return;
    const gather::Conn conn{n0, n1, n2, elist};
    nblk[r] = gather::count_cols(conn, payload, nptr[r], nptr[r + 1]);
}

// Block row r of K: bcol/bval at browptr[r].  The thread-local table of fill_row lives in local memory (L1): a
// shared-memory table (tried: profiles/r2_fused_assembly_experiments.txt) brings the DRAM writes down to the
// 2.2 GB of the blocks themselves but costs occupancy, and the kernel is bound by its dependent gathers, not by DRAM.
__global__ void __launch_bounds__(kGatherThreads)
gather_fill_kernel(const double2 *__restrict__ xy, const uint32_t *__restrict__ n0,
                   const uint32_t *__restrict__ n1, const uint32_t *__restrict__ n2,
                   const uint32_t *__restrict__ elist, const uint32_t *__restrict__ payload,
                   const uint32_t *__restrict__ nptr, uint32_t n_own, const uint32_t *__restrict__ browptr,
                   uint32_t *__restrict__ brow, uint32_t *__restrict__ bcol, double *__restrict__ bval) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_own) return;
    const gather::Conn conn{n0, n1, n2, elist};
    const uint32_t b0 = browptr[r], b1 = browptr[r + 1];
    for (uint32_t b = b0; b < b1; ++b) brow[b] = r;        // row of every block: the elimination works per block
    gather::fill_row(conn, xy, c_mat.D, c_mat.t, payload, nptr[r], nptr[r + 1], b1 - b0, bcol + b0,
                     bval + (size_t)b0 * 4);
}

// Fills K (owned node rows [K.node_lo, K.node_hi)) from the local element list.  The material must have
// been uploaded (upload_material).  ms_sort / ms_reduce receive the incidence sort and the two gather passes.
static void assemble_gather(mag_ctx *ctx, const DevBuf<double2> &xy, const DevBuf<uint32_t> &n0,
                            const DevBuf<uint32_t> &n1, const DevBuf<uint32_t> &n2, const uint32_t *elist,
                            size_t n_local, size_t n_nodes, BsrMatrix &K, float *ms_sort, float *ms_reduce) {
    EventTimer phase(ctx->stream);
    phase.start();
    const uint32_t n_own = K.node_hi - K.node_lo;
    const size_t n_inc = n_local * 3;
    const int bits = bits_for(n_nodes + 1);          // the sentinel's low bits exceed every node id
    DevBuf<uint32_t> nptr(ctx, (size_t)n_own + 1);
    nptr.zero();
    DevBuf<uint32_t> keys(ctx, n_inc), keys_alt(ctx, n_inc);     // node ids fit 32 bits: a third less sort traffic
    DevBuf<uint32_t> pay(ctx, n_inc), pay_alt(ctx, n_inc);
    if (n_local) {
        MAG_LAUNCH(ctx, emit_incidence_kernel, cdiv(n_local, 256), 256, 0, (const uint32_t *)n0.p,
                   (const uint32_t *)n1.p, (const uint32_t *)n2.p, elist, n_local, K.node_lo, K.node_hi, keys.p,
                   pay.p, nptr.p);
        radix_sort_pairs<uint32_t>(ctx, keys.p, pay.p, keys_alt.p, pay_alt.p, n_inc, bits);
    }
    keys_alt.release();
    pay_alt.release();
    keys.release();                                  // the per-node offsets replace the sorted keys
    exclusive_scan_u32(ctx, nptr.p, n_own, nptr.p, (size_t)n_own + 1);
    *ms_sort = phase.stop();

    phase.start();
    K.browptr.alloc(ctx, (size_t)n_own + 1);
    K.browptr.zero();
    if (n_own)
        MAG_LAUNCH(ctx, gather_count_kernel, cdiv(n_own, kGatherThreads), kGatherThreads, 0,
                   (const uint32_t *)n0.p, (const uint32_t *)n1.p, (const uint32_t *)n2.p, elist,
                   (const uint32_t *)pay.p, (const uint32_t *)nptr.p, n_own, K.browptr.p);
    exclusive_scan_u32(ctx, K.browptr.p, n_own, K.browptr.p, (size_t)n_own + 1);
    uint32_t n_blocks = 0;
    MAG_CUDA(cudaMemcpyAsync(&n_blocks, K.browptr.p + n_own, sizeof n_blocks, cudaMemcpyDeviceToHost, ctx->stream));
    MAG_CUDA(cudaStreamSynchronize(ctx->stream));
    K.n_blocks = n_blocks;
    K.brow.alloc(ctx, K.n_blocks);
    K.bcol.alloc(ctx, K.n_blocks);
    K.bval.alloc(ctx, (size_t)K.n_blocks * 4);
    if (n_own && K.n_blocks)
        MAG_LAUNCH(ctx, gather_fill_kernel, cdiv(n_own, kGatherThreads), kGatherThreads, 0,
                   (const double2 *)xy.p, (const uint32_t *)n0.p, (const uint32_t *)n1.p, (const uint32_t *)n2.p,
                   elist, (const uint32_t *)pay.p, (const uint32_t *)nptr.p, n_own, (const uint32_t *)K.browptr.p,
                   K.brow.p, K.bcol.p, K.bval.p);
    *ms_reduce = phase.stop();
}

}  // namespace mag
