// nccl_dyn.cuh — NCCL bound at run time (dlopen), so libxrsfm_b200.so has no link-time
// dependency on it: a single-GPU process never touches NCCL, a multi-GPU process picks up the
// libnccl.so.2 that is already mapped (torch's bundled copy) or the system one.  Only the handful
// of entry points the BA exchange needs are declared, with NCCL's stable C ABI (nccl.h 2.x).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace xrb {

struct NcclApi {
    typedef struct { char internal[128]; } UniqueId;  // ncclUniqueId
    typedef void *Comm;                               // ncclComm_t
    enum { kSum = 0 };                                // ncclRedOp_t::ncclSum
    enum { kUint8 = 1, kFloat64 = 8 };                // ncclDataType_t::ncclUint8 / ncclDouble
    int (*GetUniqueId)(UniqueId *) = nullptr;
    int (*CommInitRank)(Comm *, int nranks, UniqueId id, int rank) = nullptr;
    int (*CommDestroy)(Comm) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int dtype, int op, Comm, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int dtype, int root, Comm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int *) = nullptr;
};

// nullptr (and xrb_last_error set) when no usable libnccl.so.2 can be found.
const NcclApi *nccl_api();

}  // namespace xrb
