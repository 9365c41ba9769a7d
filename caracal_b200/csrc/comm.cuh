// comm.cuh -- the one exchange step of the path, behind the C-ABI: NCCL all-reduce of the kappa(t) sums
// (child_evol + 1 doubles) and of the umbrella-window statistics over the ranks of one job.
//
// Replaces the reference's point-to-point result traffic: recross.f90:390,411 (`message(child_evol+2)` from every
// worker to rank 0) and the statistics files of calc_rate.f90:1690-1734.  One process per GPU, one handle per
// process; rank 0 makes a unique id (crcl_comm_unique_id), the caller ships its 128 bytes to the other ranks
// (mpi_bcast in the Fortran drivers, torch.distributed / a file in Python), every rank calls crcl_comm_init.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, or $CRCL_NCCL_LIB): the library has no link-time dependency on
// it, loads on single-GPU installations without NCCL, and inside a process that already carries a NCCL (PyTorch's
// bundled copy) it resolves to that same copy instead of loading a second one.
#pragma once
#include <dlfcn.h>
#include <nccl.h>
#include <cstdlib>
#include <string>

namespace crcl {

struct NcclApi {
    void* dso = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string err;
};

// nullptr (and the reason in *err) when NCCL cannot be bound
inline NcclApi* nccl_api(std::string* err)
{
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* env = getenv("CRCL_NCCL_LIB");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        void* d = nullptr;
        for (const char* n : names) {
            if (!n || !*n) continue;
            d = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (d) break;
        }
        if (!d) {
            const char* e = dlerror();
            api.err = std::string("cannot load NCCL (libnccl.so.2; set CRCL_NCCL_LIB): ") + (e ? e : "");
        } else {
            bool ok = true;
            auto bind = [&](const char* name) {
                void* p = dlsym(d, name);
                if (!p) {
                    ok = false;
                    api.err = std::string("NCCL symbol missing: ") + name;
                }
                return p;
            };
            api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(bind("ncclGetUniqueId"));
            api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(bind("ncclCommInitRank"));
            api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(bind("ncclCommDestroy"));
            api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(bind("ncclAllReduce"));
            api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(bind("ncclGroupStart"));
            api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(bind("ncclGroupEnd"));
            api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(bind("ncclGetErrorString"));
            api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(bind("ncclGetVersion"));
            if (ok) api.dso = d;
        }
    }
    if (!api.dso) {
        if (err) *err = api.err;
        return nullptr;
    }
    return &api;
}

struct Comm {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
};

// contiguous block [start, start+count) of n units for `rank` of `world`; blocks differ by at most one unit
// (the same rule as caracal_b200/shard.py::shard_range)
inline void shard_range(long long n, int rank, int world, long long* start, long long* count)
{
    const long long base = n / world, rem = n % world;
    *count = base + (rank < rem ? 1 : 0);
    *start = rank * base + (rank < rem ? rank : rem);
}

}  // namespace crcl
