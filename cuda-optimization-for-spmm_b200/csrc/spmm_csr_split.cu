// spmm_csr_split.cu -- CSR SpMM with an nnz split that CUTS rows (merge-path style), CSR variant 6.
//
// The row kernels (variants 1, 2, 4) give whole rows to warps: with few, skewed rows (GL7d25 of the reference's
// data/large_21074: 2798 rows, 2..422 non-zeros) the longest row is the critical path and most of the machine idles
// (0.072 ms against cuSPARSE's 0.047 ms).  Here every warp gets the same number of consecutive NON-ZEROS:
//   * rows that lie completely inside a warp's range are accumulated and stored as in the row kernels;
//   * a row cut by a range boundary leaves a partial sum in a carry slot: slot 2w   = head of warp w (the row began in an
//     earlier warp), slot 2w+1 = tail of warp w (the row began here and continues in later warps);
//   * a second kernel adds the partial sums of each cut row IN ENTRY ORDER (tail slot of the warp where the row starts,
//     then the head slots of the following warps) and stores the row.
// No atomics, fixed order: run-to-run bit-reproducible.  A cut row is rounded differently from the sequential sum of
// the row kernels (partial sums are formed first), within the stated tolerance; rows that are not cut are bit-identical.
// Needs 2 * 129 * 4 bytes of caller-provided workspace per (warp, 128-column tile) (cuspmm_spmm_csr_workspace).
#include "common.cuh"

#include <stdlib.h>

namespace cuspmm_b200 {
namespace split {

constexpr int kJ = 8;                 // B rows in flight per lane
constexpr int kWarpsPerBlock = 8;

struct Plan { uint32_t warps, perWarp, tiles; };

// entries per warp: a multiple of 32, at least 64; enough warps for ~64 resident warps per SM over all column tiles
static Plan make_plan(uint32_t nnz, uint32_t N) {
    Plan p;
    p.tiles = (N + 127) / 128;
    const uint64_t maxWarps = (uint64_t)sm_count() * 64 / p.tiles + 1;
    static const uint32_t minPer = getenv("CUSPMM_SPLIT_PER") ? (uint32_t)atoi(getenv("CUSPMM_SPLIT_PER")) : 64u;   // tuning hook
    uint64_t warps = ((uint64_t)nnz + minPer - 1) / minPer;
    if (warps > maxWarps) warps = maxWarps;
    if (warps == 0) warps = 1;
    uint64_t per = ((uint64_t)nnz + warps - 1) / warps;
    per = (per + 31) / 32 * 32;
    if (per < minPer) per = minPer;
    p.perWarp = (uint32_t)per;
    p.warps = (uint32_t)(((uint64_t)nnz + per - 1) / per);
    if (p.warps == 0) p.warps = 1;
    return p;
}

// 4 consecutive columns of a row: one 128-bit access when B / C rows are 16-byte aligned (VEC), else guarded scalar ones
template <bool VEC>
__device__ __forceinline__ float4 load4(const float *__restrict__ p, uint32_t col, uint32_t N) {
    if constexpr (VEC) return __ldg(reinterpret_cast<const float4 *>(p));
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < N) v.x = __ldg(p);
    if (col + 1 < N) v.y = __ldg(p + 1);
    if (col + 2 < N) v.z = __ldg(p + 2);
    if (col + 3 < N) v.w = __ldg(p + 3);
    return v;
}
template <bool VEC>
__device__ __forceinline__ void store4(float *__restrict__ p, uint32_t col, uint32_t N, const float4 &v) {
    if constexpr (VEC) { __stcs(reinterpret_cast<float4 *>(p), v); return; }
    if (col < N) p[0] = v.x;
    if (col + 1 < N) p[1] = v.y;
    if (col + 2 < N) p[2] = v.z;
    if (col + 3 < N) p[3] = v.w;
}

// entries [s, t) of one row, in order, into acc (one float4 of the 128-column tile per lane)
template <bool VEC>
__device__ __forceinline__ void accumulate(const uint32_t *__restrict__ colIdxs, const float *__restrict__ vals, uint32_t s, uint32_t t,
                                           const float *__restrict__ Bcol, size_t ldb, bool valid, uint32_t lane, uint32_t col, uint32_t N,
                                           float4 &acc) {
    for (uint32_t base = s; base < t; base += 32) {
        uint32_t mc = 0;
        float mv = 0.f;
        if (base + lane < t) { mc = ld_stream(colIdxs + base + lane); mv = ld_stream(vals + base + lane); }
        const int cnt = (int)min(32u, t - base);
        int j = 0;
        for (; j + kJ <= cnt; j += kJ) {
            float4 b[kJ];
            float v[kJ];
#pragma unroll
            for (int k = 0; k < kJ; ++k) {
                const uint32_t c = __shfl_sync(0xFFFFFFFFu, mc, j + k);
                v[k] = __shfl_sync(0xFFFFFFFFu, mv, j + k);
                if (valid) b[k] = load4<VEC>(Bcol + (size_t)c * ldb, col, N);
            }
#pragma unroll
            for (int k = 0; k < kJ; ++k)
                if (valid) fma4(acc, v[k], b[k]);
        }
        for (; j < cnt; ++j) {
            const uint32_t c = __shfl_sync(0xFFFFFFFFu, mc, j);
            const float v = __shfl_sync(0xFFFFFFFFu, mv, j);
            if (valid) fma4(acc, v, load4<VEC>(Bcol + (size_t)c * ldb, col, N));
        }
    }
}

// carry slots of one column tile: slot (w, k) = 128 floats at carryVal[((tile * warps + w) * 2 + k) * 128], row number (or -1)
// at carryRow[(tile * warps + w) * 2 + k]
template <bool VEC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
csr_nnzsplit_kernel(const uint32_t *__restrict__ rowPtrs, const uint32_t *__restrict__ colIdxs, const float *__restrict__ vals,
                    uint32_t M, uint32_t nnz, uint32_t warps, uint32_t perWarp,
                    const float *__restrict__ B, uint32_t N, size_t ldb, float *__restrict__ C, size_t ldc,
                    int32_t *__restrict__ carryRow, float *__restrict__ carryVal) {
    const uint32_t lane = lane_id();
    const uint32_t w = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (w >= warps) return;
    const uint32_t col = blockIdx.y * 128u + lane * 4u;
    const bool valid = col < N;
    const float *Bcol = B + col;
    // rowPtrs may be a row-panel view: entry numbers are absolute, the panel's first entry is rowPtrs[0]
    const uint32_t first = __ldg(rowPtrs);
    const uint32_t end = first + nnz;
    const bool last = (w + 1 == warps);
    const uint32_t e0 = first + w * perWarp;
    const uint32_t e1 = last ? end : min(end, e0 + perWarp);
    const size_t slot = ((size_t)blockIdx.y * warps + w) * 2;
    float4 *cv = reinterpret_cast<float4 *>(carryVal + slot * 128) + lane;
    int32_t headRow = -1, tailRow = -1;

    // r = first row that STARTS at or after e0 (rowPtrs[M] = end >= e0, so r <= M)
    auto key = [&](uint32_t p) -> uint64_t { return (uint64_t)__ldg(rowPtrs + p); };
    uint32_t r = warp_lower_bound(M, (uint64_t)e0, key);
    // head: row r - 1 began in an earlier warp and reaches into this range
    if (r > 0 && e0 < e1 && __ldg(rowPtrs + r) > e0) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        accumulate<VEC>(colIdxs, vals, e0, min(__ldg(rowPtrs + r), e1), Bcol, ldb, valid, lane, col, N, acc);
        headRow = (int32_t)(r - 1);
        cv[0] = acc;
    }
    // rows that start inside [e0, e1); the last warp also takes the empty rows at the very end
    for (; r < M; ++r) {
        const uint32_t rs = __ldg(rowPtrs + r);
        if (rs >= e1 && !(last && rs == end)) break;
        const uint32_t re = __ldg(rowPtrs + r + 1);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        accumulate<VEC>(colIdxs, vals, rs, min(re, e1), Bcol, ldb, valid, lane, col, N, acc);
        if (re <= e1) {
            if (valid) store4<VEC>(C + (size_t)r * ldc + col, col, N, acc);
        } else {                                                   // cut: continues in the next warp(s)
            tailRow = (int32_t)r;
            cv[32] = acc;                                          // slot k = 1
            break;
        }
    }
    if (lane == 0) { carryRow[slot] = headRow; carryRow[slot + 1] = tailRow; }
}

// one warp per (tile, warp of the first kernel): if that warp left a tail, add the heads of the following warps in entry order
template <bool VEC>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
csr_nnzsplit_fixup_kernel(uint32_t warps, uint32_t N, float *__restrict__ C, size_t ldc,
                          const int32_t *__restrict__ carryRow, const float *__restrict__ carryVal) {
    const uint32_t lane = lane_id();
    const uint32_t w = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (w >= warps) return;
    const uint32_t col = blockIdx.y * 128u + lane * 4u;
    const int32_t *cr = carryRow + (size_t)blockIdx.y * warps * 2;
    const float4 *cv = reinterpret_cast<const float4 *>(carryVal + (size_t)blockIdx.y * warps * 2 * 128) + lane;
    const int32_t r = cr[2 * w + 1];
    if (r < 0 || col >= N) return;
    float4 sum = cv[((size_t)2 * w + 1) * 32];
    for (uint32_t v = w + 1; v < warps && cr[2 * v] == r; ++v) {
        const float4 p = cv[(size_t)2 * v * 32];
        sum.x += p.x; sum.y += p.y; sum.z += p.z; sum.w += p.w;
    }
    store4<VEC>(C + (size_t)r * ldc + col, col, N, sum);
}

} // namespace split

size_t spmm_csr_split_workspace(uint32_t nnz, uint32_t N) {
    const split::Plan p = split::make_plan(nnz, N);
    return (size_t)p.tiles * p.warps * 2 * (128 + 1) * 4;
}

int spmm_csr_split(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals, uint32_t M, uint32_t K, uint32_t nnz,
                   const float *B, uint32_t N, size_t ldb, float *C, size_t ldc, void *ws, size_t wsBytes, cudaStream_t st) {
    (void)K;
    const bool vok = (N % 4 == 0) && (ldb % 4 == 0) && (ldc % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    const split::Plan p = split::make_plan(nnz, N);
    const size_t need = spmm_csr_split_workspace(nnz, N);
    if (!ws || wsBytes < need)
        return set_error(CUSPMM_ERR_WORKSPACE, "CSR variant 6 needs %zu bytes of workspace, got %zu", need, wsBytes);
    if (reinterpret_cast<uintptr_t>(ws) & 15) return set_error(CUSPMM_ERR_INVALID, "the workspace must be 16-byte aligned");
    float *carryVal = static_cast<float *>(ws);
    int32_t *carryRow = reinterpret_cast<int32_t *>(carryVal + (size_t)p.tiles * p.warps * 2 * 128);
    const dim3 grid((p.warps + split::kWarpsPerBlock - 1) / split::kWarpsPerBlock, p.tiles);
    const int threads = split::kWarpsPerBlock * 32;
    if (vok) split::csr_nnzsplit_kernel<true><<<grid, threads, 0, st>>>(rowPtrs, colIdxs, vals, M, nnz, p.warps, p.perWarp, B, N, ldb, C, ldc, carryRow, carryVal);
    else split::csr_nnzsplit_kernel<false><<<grid, threads, 0, st>>>(rowPtrs, colIdxs, vals, M, nnz, p.warps, p.perWarp, B, N, ldb, C, ldc, carryRow, carryVal);
    CUSPMM_LAUNCH_CHECK("csr_nnzsplit_kernel");
    if (vok) split::csr_nnzsplit_fixup_kernel<true><<<grid, threads, 0, st>>>(p.warps, N, C, ldc, carryRow, carryVal);
    else split::csr_nnzsplit_fixup_kernel<false><<<grid, threads, 0, st>>>(p.warps, N, C, ldc, carryRow, carryVal);
    CUSPMM_LAUNCH_CHECK("csr_nnzsplit_fixup_kernel");
    return CUSPMM_OK;
}

} // namespace cuspmm_b200
