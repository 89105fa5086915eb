// NCCL is bound at run time (dlopen of libnccl.so.2) so that libmogp_b200.so loads on machines without it and
// never clashes with another NCCL copy in the process; only the multi-GPU gather needs it.
#pragma once
#include <nccl.h>

namespace mogp {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(ncclResult_t);
};

// nullptr when the library or a symbol is missing
const NcclApi* nccl_api();

}  // namespace mogp
