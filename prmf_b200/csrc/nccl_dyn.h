// Minimal run-time binding to the NCCL library the host process already uses (torch bundles
// libnccl.so.2).  Nothing is linked at build time: the four entry points are resolved with dlsym so the
// library and torch.distributed share one NCCL instance.  Prototypes follow nccl.h (2.x ABI).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>

namespace prmf {

struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm* NcclComm;
enum { kNcclSum = 0, kNcclFloat64 = 8 };   // ncclSum, ncclFloat64 / ncclDouble

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;

    bool loaded() const { return AllReduce != nullptr; }

    // Returns nullptr on success, else a static error string.
    const char* load(const char* path) {
        if (loaded()) return nullptr;
        const char* names[] = {path, "libnccl.so.2", "libnccl.so"};
        for (int i = 0; i < 3 && !lib; ++i) {
            if (!names[i]) continue;
            lib = dlopen(names[i], RTLD_NOW | RTLD_NOLOAD);      // the instance already in the process
            if (!lib) lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
        }
        if (!lib) return "libnccl.so.2 not found (pass its path to prmf_nccl_load)";
        GetUniqueId = (int (*)(NcclUniqueId*))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))dlsym(lib, "ncclCommInitRank");
        AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(lib, "ncclAllReduce");
        CommDestroy = (int (*)(NcclComm))dlsym(lib, "ncclCommDestroy");
        GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy) {
            AllReduce = nullptr;
            return "NCCL symbols missing in the loaded library";
        }
        return nullptr;
    }
};

}  // namespace prmf
