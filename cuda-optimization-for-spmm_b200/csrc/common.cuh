// common.cuh -- shared helpers for libcuspmm_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

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
#endif

} // namespace cuspmm_b200
