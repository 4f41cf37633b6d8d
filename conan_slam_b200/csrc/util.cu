// util.cu — error reporting and library-level entry points of libcslam.so.
#include <stdarg.h>

#include <atomic>

#include <dlfcn.h>

#include "common.cuh"
#include "nccl_dl.cuh"

namespace cslam {
static thread_local char g_last_error[512] = "";
void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}
static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
const NcclApi* nccl_api() {
    static NcclApi api;
    static int state = 0;  // 0 = not tried, 1 = ok, -1 = failed
    if (state == 1) return &api;
    if (state == -1) return nullptr;
    void* so = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!so) so = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!so) {
        set_last_error("cannot load libnccl.so.2: %s", dlerror());
        state = -1;
        return nullptr;
    }
#define LOAD(name)                                                            \
    api.name = reinterpret_cast<decltype(api.name)>(dlsym(so, "nccl" #name)); \
    if (!api.name) {                                                          \
        set_last_error("libnccl.so.2 lacks nccl" #name);                      \
        state = -1;                                                           \
        return nullptr;                                                       \
    }
    LOAD(GetUniqueId) LOAD(CommInitRank) LOAD(CommDestroy) LOAD(AllReduce) LOAD(Broadcast) LOAD(AllGather)
    LOAD(GetErrorString)
#undef LOAD
    state = 1;
    return &api;
}
}  // namespace cslam

extern "C" {
int cslam_nccl_unique_id(void* out128) {
    CSLAM_REQUIRE(out128 != nullptr, CSLAM_ERR_BAD_ARG, "out is null");
    const cslam::NcclApi* api = cslam::nccl_api();
    if (!api) return CSLAM_ERR_NCCL;
    ncclUniqueId id;
    CSLAM_NCCL(api->GetUniqueId(&id));
    memcpy(out128, &id, sizeof(id));
    return CSLAM_OK;
}
unsigned long long cslam_kernel_launches(void) { return cslam::g_launches.load(); }
const char* cslam_last_error(void) { return cslam::g_last_error; }
int cslam_version(void) { return CSLAM_VERSION; }
int cslam_device_count(int* count) {
    CSLAM_REQUIRE(count != nullptr, CSLAM_ERR_BAD_ARG, "count is null");
    *count = 0;
    CSLAM_CUDA(cudaGetDeviceCount(count));
    return CSLAM_OK;
}
}
