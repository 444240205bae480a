// NCCL resolved at run time (the copy already loaded by the process -- torch's -- else
// SA_NCCL_LIB, else libnccl.so.2): the library has no link-time dependency on it.  Shared by
// dist.cu (row-partitioned solve), tentative.cu and dist_setup.cu (sharded setup).
#pragma once
#include <nccl.h>

#include "sa_gpu_internal.cuh"

struct NcclApi
{
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

const NcclApi &sa_nccl(); // dist.cu

#define SA_NCCL(call)                                                              \
    do {                                                                           \
        ncclResult_t r__ = (call);                                                 \
        if (r__ != ncclSuccess)                                                    \
            SA_FAIL("%s:%d: %s -> %s", __FILE__, __LINE__, #call, sa_nccl().GetErrorString(r__)); \
    } while (0)

struct sa_gpu_comm
{
    sa_gpu_ctx *ctx = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
};

/* dist_setup.cu: in-place all-gather-v of a device array laid out for the full set: rank q's
   slice is [offs[q], offs[q + 1]) (elements of `elem_bytes` bytes); one grouped ncclSend /
   ncclRecv exchange on the context's stream. */
void sa_dev_allgatherv(sa_gpu_comm *C, void *buf, const int64_t *offs, size_t elem_bytes);
/* rank that owns AE `ae` under the contiguous ranges ae_part[0..nranks] */
static inline int sa_rank_of(const int *part, int nranks, int i)
{
    int lo = 0, hi = nranks - 1;
    while (lo < hi)
    {
        const int mid = (lo + hi + 1) >> 1;
        if (part[mid] <= i)
            lo = mid;
        else
            hi = mid - 1;
    }
    return lo;
}
