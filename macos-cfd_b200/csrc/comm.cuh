// comm.cuh -- row-slab communication: NCCL over NVLink 5 / NVSwitch, loaded lazily with dlopen so that a
// single-GPU handle has no dependency on libnccl at all.
//
// What moves between slabs: whole rows (contiguous in memory -- no packing) to the two neighbours,
// and a handful of doubles through allreduce.  Every message is a few hundred KiB at most, i.e. latency-, not
// bandwidth-bound; all calls are enqueued on the solver's stream, so they stay ordered with the kernels and never
// need the host.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstddef>

// the few NCCL declarations we need (kept ABI-compatible with nccl.h 2.x)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclChar = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5,
               ncclFloat16 = 6, ncclFloat32 = 7, ncclFloat64 = 8, ncclDouble = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    const char *error = nullptr;

    bool load() {
        if (handle) return true;
        // torch bundles its own libnccl.so.2; if it is already mapped the loader hands us that one
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) {
            error = "libnccl.so.2 not found";
            return false;
        }
#define NCCL_SYM(field, name)                                                                                   \
    field = reinterpret_cast<decltype(field)>(dlsym(handle, name));                                             \
    if (!field) {                                                                                               \
        error = "symbol " name " missing in libnccl";                                                           \
        return false;                                                                                           \
    }
        NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        NCCL_SYM(CommInitRank, "ncclCommInitRank")
        NCCL_SYM(CommDestroy, "ncclCommDestroy")
        NCCL_SYM(AllReduce, "ncclAllReduce")
        NCCL_SYM(AllGather, "ncclAllGather")
        NCCL_SYM(Send, "ncclSend")
        NCCL_SYM(Recv, "ncclRecv")
        NCCL_SYM(GroupStart, "ncclGroupStart")
        NCCL_SYM(GroupEnd, "ncclGroupEnd")
        NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef NCCL_SYM
        return true;
    }
};

inline NcclApi &nccl_api() {
    static NcclApi api;
    return api;
}

// Rows this rank owns: [ny*rank/nranks, ny*(rank+1)/nranks)  (64-bit product: ny*rank can exceed 2^31)
inline void slab_partition(int ny, int rank, int nranks, int *row0, int *nrows) {
    long long a = (long long)ny * rank / nranks, b = (long long)ny * (rank + 1) / nranks;
    *row0 = (int)a;
    *nrows = (int)(b - a);
}
