// nccl_dl.cuh — NCCL resolved at run time (dlopen "libnccl.so.2"): libcslam.so has no link-time
// NCCL dependency, single-GPU users never load it, and inside a torch process the already-loaded
// (torch-bundled) NCCL is the one that gets used.  Types come from <nccl.h>.
#pragma once
#include <nccl.h>

namespace cslam {

struct NcclApi {
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
};
// nullptr (and cslam_last_error set) when NCCL cannot be loaded
const NcclApi* nccl_api();

#define CSLAM_NCCL(call)                                                                          \
    do {                                                                                          \
        ncclResult_t r__ = (call);                                                                \
        if (r__ != ncclSuccess) {                                                                 \
            ::cslam::set_last_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,                  \
                                    ::cslam::nccl_api()->GetErrorString(r__));                    \
            return CSLAM_ERR_NCCL;                                                                \
        }                                                                                         \
    } while (0)

}  // namespace cslam
