#include <dlfcn.h>
#include <cstdlib>

#include "nccl_dyn.h"

namespace mogp {

const NcclApi* nccl_api() {
    static NcclApi api;
    static int state = 0;  // 0 = untried, 1 = ok, -1 = failed
    if (state == 0) {
        state = -1;
        const char* override_path = getenv("MOGP_NCCL_LIB");
        const char* names[] = {override_path, "libnccl.so.2", "libnccl.so"};
        void* lib = nullptr;
        for (const char* nm : names) {
            if (!nm) continue;
            lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
            if (lib) break;
        }
        if (lib) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(lib, "ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(lib, "ncclCommDestroy");
            api.AllGather = (decltype(api.AllGather))dlsym(lib, "ncclAllGather");
            api.AllReduce = (decltype(api.AllReduce))dlsym(lib, "ncclAllReduce");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(lib, "ncclGetErrorString");
            if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.AllReduce &&
                api.GetErrorString)
                state = 1;
        }
    }
    return state == 1 ? &api : nullptr;
}

}  // namespace mogp
