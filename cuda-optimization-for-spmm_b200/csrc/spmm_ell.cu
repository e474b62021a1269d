// spmm_ell.cu -- ELL SpMM for sm_100a on a sliced, slot-major ("column-major") layout.
//
// Replaces spmmELLK1/K2 (src/spmm/ell/spmm_ell_k{1,2}.cu: one thread per slot of the
// column-ELL doing N global atomicAdds into the host-resident CPU result).  Layout
// (DESIGN.md "Sliced ELL"): slices of 32 rows; slice s is W_s slots wide; entry j of row
// s*32+i lives at slicePtrs[s] + j*32 + i, so the 32 rows' j-th entries are one coalesced
// 128-byte line.  One CTA per slice (x column tile): the CTA copies the slice's (col,val)
// slots into shared memory with fully coalesced loads, CH slots at a time; warp w then
// walks rows 4w..4w+3 reading its entries as shared-memory broadcasts and the B rows with
// 128-bit loads.  No atomics; per C element the terms are added in ascending column
// order with fp32 FMA, which is the order spmmELLCpu produces (src/spmm/ell/spmm_ell.cpp:15-28).
#include "common.cuh"

namespace cuspmm_b200 {

constexpr int kSliceH = 32;
constexpr int kEllChunk = 32;   // slots staged per round: 32*32*8 B = 8 KB of shared memory

template <int U>
__global__ void __launch_bounds__(256)
sell_vec_kernel(const uint32_t *__restrict__ slicePtrs, const uint32_t *__restrict__ colIdxs,
                const float *__restrict__ vals, uint32_t M,
                const float *__restrict__ B, uint32_t N, size_t ldb, float *__restrict__ C, size_t ldc) {
    __shared__ uint32_t s_col[kEllChunk * kSliceH];
    __shared__ float s_val[kEllChunk * kSliceH];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t slice = blockIdx.x;
    const uint32_t base = __ldg(slicePtrs + slice);
    const uint32_t width = (__ldg(slicePtrs + slice + 1) - base) / kSliceH;
    const uint32_t col0 = blockIdx.y * (128u * U) + lane * 4u;
    bool valid[U];
#pragma unroll
    for (int u = 0; u < U; ++u) valid[u] = (col0 + u * 128u) < N;

    float4 acc[4][U];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int u = 0; u < U; ++u) acc[i][u] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (uint32_t j0 = 0; j0 < width; j0 += kEllChunk) {
        const uint32_t nslots = min((uint32_t)kEllChunk, width - j0);
        __syncthreads();     // previous round fully consumed
        for (uint32_t t = threadIdx.x; t < nslots * kSliceH; t += blockDim.x) {
            s_col[t] = ld_stream(colIdxs + base + (size_t)j0 * kSliceH + t);
            s_val[t] = ld_stream(vals + base + (size_t)j0 * kSliceH + t);
        }
        __syncthreads();
        for (uint32_t j = 0; j < nslots; ++j) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t c = s_col[j * kSliceH + warp * 4 + i];     // broadcast read
                if (c != kPad) {
                    const float v = s_val[j * kSliceH + warp * 4 + i];
                    const float *brow = B + (size_t)c * ldb + col0;
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (valid[u]) fma4(acc[i][u], v, __ldg(reinterpret_cast<const float4 *>(brow + u * 128)));
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t r = slice * kSliceH + warp * 4 + i;
        if (r < M) {
            float *crow = C + (size_t)r * ldc + col0;
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (valid[u]) __stcs(reinterpret_cast<float4 *>(crow + u * 128), acc[i][u]);
        }
    }
}

// any N / alignment: lane = column, 4 columns tiles of 32 per lane
template <int U>
__global__ void __launch_bounds__(256)
sell_scalar_kernel(const uint32_t *__restrict__ slicePtrs, const uint32_t *__restrict__ colIdxs,
                   const float *__restrict__ vals, uint32_t M,
                   const float *__restrict__ B, uint32_t N, size_t ldb, float *__restrict__ C, size_t ldc) {
    __shared__ uint32_t s_col[kEllChunk * kSliceH];
    __shared__ float s_val[kEllChunk * kSliceH];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t slice = blockIdx.x;
    const uint32_t base = __ldg(slicePtrs + slice);
    const uint32_t width = (__ldg(slicePtrs + slice + 1) - base) / kSliceH;
    const uint32_t col0 = blockIdx.y * (32u * U) + lane;
    float acc[4][U];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int u = 0; u < U; ++u) acc[i][u] = 0.f;
    for (uint32_t j0 = 0; j0 < width; j0 += kEllChunk) {
        const uint32_t nslots = min((uint32_t)kEllChunk, width - j0);
        __syncthreads();
        for (uint32_t t = threadIdx.x; t < nslots * kSliceH; t += blockDim.x) {
            s_col[t] = ld_stream(colIdxs + base + (size_t)j0 * kSliceH + t);
            s_val[t] = ld_stream(vals + base + (size_t)j0 * kSliceH + t);
        }
        __syncthreads();
        for (uint32_t j = 0; j < nslots; ++j) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t c = s_col[j * kSliceH + warp * 4 + i];
                if (c != kPad) {
                    const float v = s_val[j * kSliceH + warp * 4 + i];
                    const float *brow = B + (size_t)c * ldb;
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (col0 + u * 32u < N) acc[i][u] = fmaf(v, __ldg(brow + col0 + u * 32u), acc[i][u]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t r = slice * kSliceH + warp * 4 + i;
        if (r < M)
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (col0 + u * 32u < N) C[(size_t)r * ldc + col0 + u * 32u] = acc[i][u];
    }
}

int spmm_sell_rows_dispatch(const uint32_t *, const uint32_t *, const float *, uint32_t, uint32_t, uint32_t,
                            const float *, uint32_t, size_t, float *, size_t, int, cudaStream_t);

// variants: 0 auto (the CSR selector applied to the slot count), 1 row kernels (warp / sub-warp per row,
// nnz-balanced) reading the sliced layout, 2 staged TMA kernel, 3 slice per CTA (A staged through smem),
// 4 staged with the dual operand path, 5 every B read from tensor memory (CSR variant 7), 6 tensor cores (CSR variant 8 reading
// the sliced layout)
static int spmm_sell_dispatch(const uint32_t *slicePtrs, const uint32_t *colIdxs, const float *vals,
                              uint32_t M, uint32_t K, uint32_t sliceH, uint32_t numSlots, const float *B, uint32_t N,
                              size_t ldb, float *C, size_t ldc, int variant, cudaStream_t st) {
    CUSPMM_REQUIRE(variant >= 0 && variant <= CUSPMM_ELL_NUM_VARIANTS, "ELL variant %d does not exist", variant);
    CUSPMM_REQUIRE(sliceH == kSliceH, "sliced ELL kernels are built for slices of %d rows (got %u)", kSliceH, sliceH);
    CUSPMM_REQUIRE(ldb >= N && ldc >= N, "ldb/ldc must be >= N");
    if (M == 0 || N == 0) return CUSPMM_OK;
    CUSPMM_REQUIRE(slicePtrs && B && C, "null operand pointer");
    const uint32_t slices = (M + kSliceH - 1) / kSliceH;
    const bool vok = (N % 4 == 0) && (ldb % 4 == 0) && (ldc % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    if (variant == 0) return spmm_sell_rows_dispatch(slicePtrs, colIdxs, vals, M, K, numSlots, B, N, ldb, C, ldc, 0, st);
    if (variant == 1) {   // row kernels; the selector chooses among warp / sub-warp / scalar, never staged
        int v = (!vok) ? 4 : ((N <= 512 || (double)numSlots / M < 96.0) ? 2 : 1);
        return spmm_sell_rows_dispatch(slicePtrs, colIdxs, vals, M, K, numSlots, B, N, ldb, C, ldc, v, st);
    }
    if (variant == 2) return spmm_sell_rows_dispatch(slicePtrs, colIdxs, vals, M, K, numSlots, B, N, ldb, C, ldc, 3, st);
    if (variant == 4) return spmm_sell_rows_dispatch(slicePtrs, colIdxs, vals, M, K, numSlots, B, N, ldb, C, ldc, 5, st);
    if (variant == 5) return spmm_sell_rows_dispatch(slicePtrs, colIdxs, vals, M, K, numSlots, B, N, ldb, C, ldc, 7, st);
    if (variant == 6) return spmm_sell_rows_dispatch(slicePtrs, colIdxs, vals, M, K, numSlots, B, N, ldb, C, ldc, 8, st);
    if (vok) {
        if (N > 256) sell_vec_kernel<4><<<dim3(slices, (N + 511) / 512), 256, 0, st>>>(slicePtrs, colIdxs, vals, M, B, N, ldb, C, ldc);
        else if (N > 128) sell_vec_kernel<2><<<dim3(slices, 1), 256, 0, st>>>(slicePtrs, colIdxs, vals, M, B, N, ldb, C, ldc);
        else sell_vec_kernel<1><<<dim3(slices, 1), 256, 0, st>>>(slicePtrs, colIdxs, vals, M, B, N, ldb, C, ldc);
        CUSPMM_LAUNCH_CHECK("sell_vec_kernel");
    } else {
        sell_scalar_kernel<4><<<dim3(slices, (N + 127) / 128), 256, 0, st>>>(slicePtrs, colIdxs, vals, M, B, N, ldb, C, ldc);
        CUSPMM_LAUNCH_CHECK("sell_scalar_kernel");
    }
    return CUSPMM_OK;
}

} // namespace cuspmm_b200

extern "C" int cuspmm_spmm_sell(const uint32_t *slicePtrs, const uint32_t *colIdxs, const float *vals,
                                uint32_t M, uint32_t K, uint32_t sliceH, uint32_t numSlots, const float *B, uint32_t N,
                                size_t ldb, float *C, size_t ldc, int variant, void *stream) {
    return cuspmm_b200::spmm_sell_dispatch(slicePtrs, colIdxs, vals, M, K, sliceH, numSlots, B, N, ldb, C, ldc, variant,
                                           cuspmm_b200::as_stream(stream));
}
