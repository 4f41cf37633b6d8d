// ptx_async.cuh — thin inline-PTX wrappers for the sm_100a asynchronous machinery used by the
// streaming covariance passes: mbarriers, TMA bulk copies (cp.async.bulk), tensor-map TMA tile
// loads / stores (cp.async.bulk.tensor.2d -> SASS UTMALDG / UTMASTG), the generic->async proxy
// fence, L2 eviction policies and the FP64 tensor-core instruction (mma.sync m8n8k4 -> DMMA.8x8x4).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace cslam {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Arrive that cannot be issued before `dep` — a word derived from registers that shared-memory loads fill — is
// available: used to hand a buffer back to its producer right after reading it.  A plain arrive placed behind the
// loads is issued while they may still wait in the load/store queue (e.g. behind streaming global stores that
// stall on a saturated HBM), the producer refills the buffer and the queued loads read the new contents.  The
// barrier address is selected between `bar` and `bar + skew` by a predicate computed from `dep`; `skew` must be a
// value the assembler cannot know (derived from a kernel argument) that is ZERO at run time, so both choices are
// the same barrier: a true register dependency at no semantic cost.
__device__ __forceinline__ void mbar_arrive_after(uint32_t bar, uint32_t skew, uint32_t dep) {
    asm volatile(
        "{\n .reg .pred p;\n .reg .b32 a;\n setp.eq.u32 p, %1, 0x7ff8dead;\n add.u32 a, %0, %2;\n selp.b32 a, a, %0, p;\n"
        " mbarrier.arrive.shared::cta.b64 _, [a];\n}\n" ::"r"(bar),
        "r"(dep), "r"(skew)
        : "memory");
}
__device__ __forceinline__ uint32_t hi32(double x) { return (uint32_t)__double2hiint(x); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}

// L2 eviction policy for data that is touched exactly once per pass (the covariance itself)
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// 1-D TMA bulk copy global -> shared (16-byte granularity), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                              uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            dst),
        "l"(src), "r"(bytes), "r"(bar), "l"(pol)
        : "memory");
}

__device__ __forceinline__ void tmap_prefetch(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
// 2-D tensor-map tile load global -> shared (c0 = inner / column coordinate, c1 = row coordinate, in
// elements); out-of-bounds elements of the box are zero-filled and still counted in the transaction bytes
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar,
                                            uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst),
        "l"(tm), "r"(c0), "r"(c1), "r"(bar), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_nohint(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            dst),
        "l"(tm), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d_nohint(const CUtensorMap* tm, int c0, int c1, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(c0),
                 "r"(c1), "r"(src)
                 : "memory");
}
// 2-D tensor-map tile store shared -> global (bulk async-group completion); out-of-bounds elements are dropped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int c0, int c1, uint32_t src, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;" ::"l"(tm),
        "r"(c0), "r"(c1), "r"(src), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups still READ their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// wait until at most N of this thread's bulk groups are incomplete (writes performed)
template <int N>
__device__ __forceinline__ void bulk_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// generic-proxy shared-memory writes (st.shared) become visible to the async proxy (TMA stores)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D(8x8) += A(8x4) * B(4x8) in FP64 on the tensor cores.  Lane l: a = A[l/4][l%4], b = B[l%4][l/4],
// (c0, c1) = C[l/4][2*(l%4) + {0, 1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

}  // namespace ptx
}  // namespace cslam
