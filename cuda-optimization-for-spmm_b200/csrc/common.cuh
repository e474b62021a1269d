// common.cuh -- shared helpers for libcuspmm_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <mutex>

#include "cuspmm_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libcuspmm_b200 is written for sm_100a (B200) only"
#endif

namespace cuspmm_b200 {

// thread-local last error + launch counter (capi.cu)
int set_error(int status, const char *fmt, ...);
void count_launch(unsigned n = 1);

#define CUSPMM_CUDA(call)                                                                 \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess)                                                           \
            return ::cuspmm_b200::set_error(CUSPMM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, \
                                            cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define CUSPMM_LAUNCH_CHECK(name)                                                         \
    do {                                                                                  \
        cudaError_t e__ = cudaGetLastError();                                             \
        if (e__ != cudaSuccess)                                                           \
            return ::cuspmm_b200::set_error(CUSPMM_ERR_CUDA, "launch of %s failed: %s", name, \
                                            cudaGetErrorString(e__));                     \
        ::cuspmm_b200::count_launch();                                                    \
    } while (0)

#define CUSPMM_REQUIRE(cond, ...)                                                         \
    do {                                                                                  \
        if (!(cond)) return ::cuspmm_b200::set_error(CUSPMM_ERR_INVALID, __VA_ARGS__);    \
    } while (0)

int sm_count();          // SMs of the current device (cached per device)
size_t l2_bytes();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device): thread-safe, any device ordinal.  Keyed by the
// kernel's ADDRESS: instantiations with the same signature share this function's type, so a per-type flag would be wrong.
int set_smem_once_impl(const void *kern, size_t bytes);      // capi.cu; returns a cudaError_t
template <typename Kern>
static inline cudaError_t set_smem_once(Kern kern, size_t bytes) {
    return (cudaError_t)set_smem_once_impl(reinterpret_cast<const void *>(kern), bytes);
}

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr uint32_t kPad = 0xFFFFFFFFu;   // ELL padding index (the file's "-1")

#ifdef __CUDACC__
// ----------------------------------------------------------------- device helpers
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// streaming (read-once) loads for A: keep L1/L2 for the B rows that get re-used
__device__ __forceinline__ uint32_t ld_stream(const uint32_t *p) { return __ldcs(p); }
__device__ __forceinline__ float ld_stream(const float *p) { return __ldcs(p); }

__device__ __forceinline__ void fma4(float4 &acc, float a, const float4 &b) {
    acc.x = fmaf(a, b.x, acc.x);
    acc.y = fmaf(a, b.y, acc.y);
    acc.z = fmaf(a, b.z, acc.z);
    acc.w = fmaf(a, b.w, acc.w);
}

// Smallest p in [0, n] with key(p) >= target, where key is non-decreasing and
// key(n) >= target.  32-ary search: every lane probes one point per round, so the
// dependent-load chain is ceil(log32(n)) long instead of log2(n).  All 32 lanes
// must call it with the same arguments; all lanes get the result.
template <typename KeyFn>
__device__ __forceinline__ uint32_t warp_lower_bound(uint32_t n, uint64_t target, KeyFn key) {
    uint32_t lo = 0, hi = n;
    const uint32_t lane = lane_id();
    while (lo < hi) {
        uint32_t span = hi - lo;
        uint32_t step = span / 32u + 1u;
        uint64_t p64 = (uint64_t)lo + (uint64_t)lane * step;
        uint32_t p = p64 > hi ? hi : (uint32_t)p64;
        bool ok = (p >= hi) || (key(p) >= target);
        uint32_t ballot = __ballot_sync(0xFFFFFFFFu, ok);
        int f = __ffs(ballot) - 1;            // >= 0: lane 31 may be clamped to hi, and p==hi is ok
        if (f < 0) {                          // all 32 probes lie below hi and are false
            lo = lo + 31u * step + 1u;        // = p_31 + 1 <= hi
            continue;
        }
        uint32_t pf = __shfl_sync(0xFFFFFFFFu, p, f);
        uint32_t pprev = __shfl_sync(0xFFFFFFFFu, p, f > 0 ? f - 1 : 0);
        hi = pf;
        if (f > 0) lo = pprev + 1;
    }
    return lo;
}

// Two searches in lockstep (the lower and the upper end of a warp's row range): the two probe loads of a
// round are issued back to back, so both searches together cost one dependent-load chain, not two.
template <typename KeyFn>
__device__ __forceinline__ void warp_lower_bound2(uint32_t n, uint64_t target0, uint64_t target1, KeyFn key,
                                                  uint32_t &out0, uint32_t &out1) {
    uint32_t lo0 = 0, hi0 = n, lo1 = 0, hi1 = n;
    const uint32_t lane = lane_id();
    while (lo0 < hi0 || lo1 < hi1) {
        const uint32_t step0 = (hi0 - lo0) / 32u + 1u, step1 = (hi1 - lo1) / 32u + 1u;
        const uint64_t q0 = (uint64_t)lo0 + (uint64_t)lane * step0, q1 = (uint64_t)lo1 + (uint64_t)lane * step1;
        const uint32_t p0 = q0 > hi0 ? hi0 : (uint32_t)q0, p1 = q1 > hi1 ? hi1 : (uint32_t)q1;
        const bool live0 = lo0 < hi0, live1 = lo1 < hi1;
        const uint64_t k0 = (live0 && p0 < hi0) ? key(p0) : ~0ull;
        const uint64_t k1 = (live1 && p1 < hi1) ? key(p1) : ~0ull;
        if (live0) {
            const int f = __ffs(__ballot_sync(0xFFFFFFFFu, k0 >= target0)) - 1;
            if (f < 0) lo0 = lo0 + 31u * step0 + 1u;
            else {
                const uint32_t pf = __shfl_sync(0xFFFFFFFFu, p0, f), pp = __shfl_sync(0xFFFFFFFFu, p0, f > 0 ? f - 1 : 0);
                hi0 = pf;
                if (f > 0) lo0 = pp + 1;
            }
        }
        if (live1) {
            const int f = __ffs(__ballot_sync(0xFFFFFFFFu, k1 >= target1)) - 1;
            if (f < 0) lo1 = lo1 + 31u * step1 + 1u;
            else {
                const uint32_t pf = __shfl_sync(0xFFFFFFFFu, p1, f), pp = __shfl_sync(0xFFFFFFFFu, p1, f > 0 ? f - 1 : 0);
                hi1 = pf;
                if (f > 0) lo1 = pp + 1;
            }
        }
    }
    out0 = lo0;
    out1 = lo1;
}
// ----------------------------------------------------------------- mbarrier / TMA bulk-copy primitives
// Shared by the staged CSR kernels (spmm_csr.cu), the dual-path kernel (spmm_csr_tmem.cu) and the tcgen05 BSR kernel
// (spmm_bsr_tc.cu).  The wait loops differ per kernel (plain spin / suspend-time hint) and stay with them.
namespace pipe {
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one try_wait on the phase with this parity (true: completed); hint_ns > 0: suspend-time hint for the hardware wait
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity, uint32_t hint_ns = 0) {
    uint32_t done;
    if (hint_ns)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
// blocking wait; a lost arrive must fail loudly (trap), not hang the GPU.  The bound is ELAPSED TIME (%globaltimer,
// checked every 4096 polls), not a poll count: under a debugger, MPS time-slicing or a long preemption a healthy wait
// may take arbitrarily many polls, but not 20 seconds of wall clock
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr uint64_t kWaitLimitNs = 20ull * 1000ull * 1000ull * 1000ull;
template <uint32_t HINT_NS = 0>
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t polls = 0;
    uint64_t t0 = 0;
    while (!mbar_try_wait(bar, parity, HINT_NS)) {
        if ((++polls & 4095u) == 0) {
            const uint64_t now = global_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > kWaitLimitNs) __trap();
        }
    }
}
// 1-D TMA bulk copy global -> shared, completion (bytes) on the mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
} // namespace pipe
#endif


} // namespace cuspmm_b200
