// nccl_dyn.cu — see nccl_dyn.cuh.
#include "nccl_dyn.cuh"

#include <dlfcn.h>

#include <mutex>

#include "common.cuh"

namespace xrb {

const NcclApi *nccl_api() {
    static NcclApi api;
    static bool tried = false, ok = false;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (tried) {
        if (!ok) set_error("NCCL is not available in this process (libnccl.so.2 not found)");
        return ok ? &api : nullptr;
    }
    tried = true;
    // prefer the copy that is already mapped (torch's), then the loader's search path
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        set_error("NCCL is not available in this process: %s", dlerror());
        return nullptr;
    }
    bool all = true;
    auto sym = [&](const char *name) {
        void *p = dlsym(h, name);
        if (!p) all = false;
        return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    if (!all) {
        set_error("libnccl.so.2 lacks an entry point the BA exchange needs");
        return nullptr;
    }
    ok = true;
    return &api;
}

}  // namespace xrb
