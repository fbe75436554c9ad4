// Communicator and halo engine of the device data path: the B200 counterpart of the reference's MPI wrapper (comm/MpiComm.hpp:404-600,
// as far as the hot path uses it) and of comm::ImportExportContext / Import / Export (comm/ImportExport.hpp:29-72, 131-215, 296-470).
//
// One process per GPU; ranks talk through NCCL point-to-point calls grouped per exchange (one ncclGroup of ncclSend / ncclRecv per Import
// or Export, all right-hand sides in it — the reference posts one MPI message per vector and neighbour) and ncclAllReduce for the Krylov
// scalars. NCCL is resolved at run time (dlopen of libnccl.so.2): a process that already carries a copy — torch's — shares it, and the
// library still loads where no NCCL exists (single-rank use needs none).
//
// Streams: packing, unpacking and all element work run on the context's stream S; transfers run on a communication stream C owned by
// the communicator. Ordering is by events only (no host synchronisation):
//     Import   S: pack owned -> buffer | C: wait, group{send buffer slices, recv into the ghost block of x} | S: wait (import_end)
//     Export   S: (elements done)      | C: wait, group{send ghost block slices of y, recv into buffer}    | S: wait, unpack-add
// so whatever the caller queues on S between *_begin and *_end overlaps the transfer (MatrixFreeSystem.hpp:1046-1122).
#ifndef L3B_COMM_CUH
#define L3B_COMM_CUH

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdint>
#include <string>
#include <vector>

namespace l3b::comm
{
struct NcclApi
{
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*)                                                                        = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int)                                                 = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t)                                                                           = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int*)                                                                 = nullptr;
    ncclResult_t (*CommUserRank)(const ncclComm_t, int*)                                                              = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)                          = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t)                                = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t)      = nullptr;
    ncclResult_t (*GroupStart)()                                                                                      = nullptr;
    ncclResult_t (*GroupEnd)()                                                                                        = nullptr;
    const char* (*GetErrorString)(ncclResult_t)                                                                       = nullptr;
    std::string error; // why the library could not be loaded
};

// the process-wide NCCL entry points, or an object with handle == nullptr and `error` set
inline const NcclApi& nccl()
{
    static const NcclApi api = [] {
        NcclApi a;
        for (const char* name : {"libnccl.so.2", "libnccl.so"})
        {
            a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (a.handle)
                break;
        }
        if (not a.handle)
        {
            a.error = std::string{"NCCL could not be loaded: "} + dlerror();
            return a;
        }
        bool       ok  = true;
        const auto sym = [&](auto& fn, const char* name) {
            fn = reinterpret_cast< std::remove_reference_t< decltype(fn) > >(dlsym(a.handle, name));
            ok = ok and fn != nullptr;
        };
        sym(a.GetUniqueId, "ncclGetUniqueId");
        sym(a.CommInitRank, "ncclCommInitRank");
        sym(a.CommDestroy, "ncclCommDestroy");
        sym(a.CommCount, "ncclCommCount");
        sym(a.CommUserRank, "ncclCommUserRank");
        sym(a.Send, "ncclSend");
        sym(a.Recv, "ncclRecv");
        sym(a.AllReduce, "ncclAllReduce");
        sym(a.GroupStart, "ncclGroupStart");
        sym(a.GroupEnd, "ncclGroupEnd");
        sym(a.GetErrorString, "ncclGetErrorString");
        if (not ok)
        {
            a.error  = "the NCCL library found lacks a required entry point";
            a.handle = nullptr;
        }
        return a;
    }();
    return api;
}

// dst[(nbr chunk) + c * n_i + i] = src[idx[i] + c * ld]: all neighbours and columns in one launch. `ptr` are the CSR offsets of the
// neighbours' index lists; the packed buffer holds, per neighbour, n_cols column slices back to back.
__global__ void haloPackKernel(const double* src, long long ld, const int32_t* idx, const long long* ptr, int n_nbrs, int n_cols, double* dst)
{
    const long long total = ptr[n_nbrs];
    for (long long t = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; t < total * n_cols; t += static_cast< long long >(gridDim.x) * blockDim.x)
    {
        const long long i = t % total;
        const int       c = static_cast< int >(t / total);
        int             lo = 0, hi = n_nbrs; // neighbour of entry i: last k with ptr[k] <= i
        while (hi - lo > 1)
        {
            const int mid = (lo + hi) / 2;
            if (ptr[mid] <= i)
                lo = mid;
            else
                hi = mid;
        }
        const long long n_k = ptr[lo + 1] - ptr[lo];
        dst[ptr[lo] * n_cols + c * n_k + (i - ptr[lo])] = src[idx[i] + c * ld];
    }
}
// dst[idx[i] + c * ld] += buffer entry (combine op of comm::Export: AtomicSumInto, util/Functional.hpp:102). An owned dof shared with
// several neighbours appears once per neighbour, hence the atomic.
__global__ void haloUnpackAddKernel(double* dst, long long ld, const int32_t* idx, const long long* ptr, int n_nbrs, int n_cols, const double* src)
{
    const long long total = ptr[n_nbrs];
    for (long long t = blockIdx.x * static_cast< long long >(blockDim.x) + threadIdx.x; t < total * n_cols; t += static_cast< long long >(gridDim.x) * blockDim.x)
    {
        const long long i = t % total;
        const int       c = static_cast< int >(t / total);
        int             lo = 0, hi = n_nbrs;
        while (hi - lo > 1)
        {
            const int mid = (lo + hi) / 2;
            if (ptr[mid] <= i)
                lo = mid;
            else
                hi = mid;
        }
        const long long n_k = ptr[lo + 1] - ptr[lo];
        atomicAdd(dst + idx[i] + c * ld, src[ptr[lo] * n_cols + c * n_k + (i - ptr[lo])]);
    }
}
} // namespace l3b::comm
#endif
