// dist.cu — multi-GPU plumbing: one process per GPU, NCCL communicator
// bootstrap.  (Row-block partitioned solve: see dist_solve.cuh.)
#include <nccl.h>

#include <cstring>

#include "common.cuh"

namespace mag {
struct Comm {
    ncclComm_t nccl = nullptr;
    int rank = 0, nranks = 1;
};
}  // namespace mag

using namespace mag;

extern thread_local std::string g_last_error_dist;
thread_local std::string g_last_error_dist;

extern "C" int mag_comm_unique_id(void *id128) {
    if (!id128) return MAG_ERR_BAD_ARG;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return MAG_ERR_NCCL;
    std::memcpy(id128, &id, sizeof id);
    return MAG_OK;
}

extern "C" int mag_comm_init(mag_ctx *ctx, int rank, int nranks, const void *id128) {
    if (!ctx || !id128 || nranks <= 0 || rank < 0 || rank >= nranks) return MAG_ERR_BAD_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return MAG_ERR_CUDA;
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof id);
    Comm *c = new Comm;
    c->rank = rank; c->nranks = nranks;
    if (ncclCommInitRank(&c->nccl, nranks, id, rank) != ncclSuccess) {
        delete c;
        return MAG_ERR_NCCL;
    }
    ctx->comm = c;
    return MAG_OK;
}

extern "C" int mag_comm_rank(const mag_ctx *ctx, int *rank, int *nranks) {
    if (!ctx || !rank || !nranks) return MAG_ERR_BAD_ARG;
    *rank = ctx->comm ? ctx->comm->rank : 0;
    *nranks = ctx->comm ? ctx->comm->nranks : 1;
    return MAG_OK;
}
