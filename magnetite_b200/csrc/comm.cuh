// comm.cuh — multi-GPU plumbing: one process per GPU, an NCCL communicator for
// the scalar allreduces and the one-off exchanges, CUDA IPC mappings of the
// neighbours' halo buffers for the peer-to-peer stores over NVLink.
#pragma once
#include <nccl.h>

#include "common.cuh"

namespace mag {

#define MAG_NCCL(expr)                                                                   \
    do {                                                                                 \
        ncclResult_t r_ = (expr);                                                        \
        if (r_ != ncclSuccess)                                                           \
            ::mag::fail(MAG_ERR_NCCL, "NCCL error at %s:%d: %s", __FILE__, __LINE__,     \
                        ncclGetErrorString(r_));                                         \
    } while (0)

struct Comm {
    ncclComm_t nccl = nullptr;
    int rank = 0, nranks = 1;
    // The slab the other ranks store into (Dinv halo, mailboxes, coarse partials, halo of r: solve.cuh) and its CUDA
    // IPC mappings belong to the communicator, not to a system: allocating, exporting and mapping 100+ MB costs tens
    // of milliseconds, a solve a few.  It grows when a system needs more; `generation` tells a system that holds
    // pointers into an older slab to look them up again.  `epoch` numbers the solves: the sequence numbers of all
    // messages derive from it, so nothing in the slab ever has to be reset between solves or systems.
    double *slab = nullptr;
    size_t slab_bytes = 0;
    std::vector<double *> peer_slab;        // rank r's slab as mapped here (own rank: slab)
    unsigned long long generation = 0, epoch = 0;
    size_t layout_rows = 0;                 // ext_len of the system layout the slab was last used with
};

// Blocking allgather of `count` uint32 per rank (host in, host out).
inline void allgather_u32(mag_ctx *ctx, const uint32_t *mine, int count, uint32_t *all) {
    Comm *c = ctx->comm;
    if (!c || c->nranks == 1) {
        for (int i = 0; i < count; ++i) all[i] = mine[i];
        return;
    }
    DevBuf<uint32_t> send(ctx, count), recv(ctx, (size_t)count * c->nranks);
    MAG_CUDA(cudaMemcpyAsync(send.p, mine, count * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    MAG_NCCL(ncclAllGather(send.p, recv.p, (size_t)count * sizeof(uint32_t), ncclChar, c->nccl, ctx->stream));
    MAG_CUDA(cudaMemcpyAsync(all, recv.p, (size_t)count * c->nranks * sizeof(uint32_t),
                             cudaMemcpyDeviceToHost, ctx->stream));
    MAG_CUDA(cudaStreamSynchronize(ctx->stream));
}

inline void allgather_bytes(mag_ctx *ctx, const void *mine, size_t bytes, void *all) {
    Comm *c = ctx->comm;
    DevBuf<unsigned char> send(ctx, bytes), recv(ctx, bytes * c->nranks);
    MAG_CUDA(cudaMemcpyAsync(send.p, mine, bytes, cudaMemcpyHostToDevice, ctx->stream));
    MAG_NCCL(ncclAllGather(send.p, recv.p, bytes, ncclChar, c->nccl, ctx->stream));
    MAG_CUDA(cudaMemcpyAsync(all, recv.p, bytes * c->nranks, cudaMemcpyDeviceToHost, ctx->stream));
    MAG_CUDA(cudaStreamSynchronize(ctx->stream));
}

// In-stream (graph-capturable) sum over ranks of `count` doubles, in place.
inline void allreduce_sum(mag_ctx *ctx, double *dptr, int count) {
    Comm *c = ctx->comm;
    if (!c || c->nranks == 1) return;
    MAG_NCCL(ncclAllReduce(dptr, dptr, (size_t)count, ncclDouble, ncclSum, c->nccl, ctx->stream));
}

// Every rank contributes the slice [lo[r], lo[r+1]) of a vector it owns; afterwards all
// ranks hold the whole vector (grouped broadcasts: the slices have different lengths).
inline void allgather_slices(mag_ctx *ctx, double *vec, const std::vector<uint32_t> &lo) {
    Comm *c = ctx->comm;
    if (!c || c->nranks == 1) return;
    MAG_NCCL(ncclGroupStart());
    for (int r = 0; r < c->nranks; ++r) {
        const size_t cnt = lo[r + 1] - lo[r];
        if (cnt) MAG_NCCL(ncclBroadcast(vec + lo[r], vec + lo[r], cnt, ncclDouble, r, c->nccl, ctx->stream));
    }
    MAG_NCCL(ncclGroupEnd());
}

}  // namespace mag
