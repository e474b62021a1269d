// spmm_coo.cu -- COO SpMM for sm_100a.
//
// Replaces spmmCOOK1 (src/spmm/coo/spmm_coo_k1.cu:9-27: one thread per non-zero doing N
// serial global atomicAdds).  Here the (row, col)-sorted triplet array is cut into equal
// nnz chunks, one per warp, and each cut is moved forward to the next row boundary with a
// 32-ary search on rowIdxs, so every row is owned by exactly one warp: no atomics, no
// carries, no workspace, fixed summation order (file order, fp32 FMA).  The warp also
// writes the zero rows that have no entries, so C needs no memset (beta = 0 semantics).
//
// variant 2 builds CSR row pointers from rowIdxs on the device (convert.cu) and runs the
// staged CSR kernel, which is the better kernel for dense-ish matrices.
#include "common.cuh"

namespace cuspmm_b200 {

int spmm_csr_dispatch(const uint32_t *, const uint32_t *, const float *, uint32_t, uint32_t, uint32_t,
                      const float *, uint32_t, size_t, float *, size_t, int, cudaStream_t);
int coo_to_csr_rowptrs(const uint32_t *rowIdxs, uint32_t M, uint32_t nnz, uint32_t *rowPtrs, cudaStream_t st);
int csr_select_variant(uint32_t M, uint32_t K, uint64_t nnz, uint32_t N, bool vec_ok, bool sell = false);

// first index i in [from, nnz] whose row is > `row` (i.e. the start of the next row)
__device__ __forceinline__ uint32_t next_row_start(const uint32_t *__restrict__ rowIdxs, uint32_t nnz,
                                                   uint32_t from) {
    if (from == 0) return 0;
    if (from >= nnz) return nnz;
    const uint32_t row = __ldg(rowIdxs + from - 1);
    auto key = [&](uint32_t p) -> uint64_t { return (uint64_t)__ldg(rowIdxs + from + p); };
    return from + warp_lower_bound(nnz - from, (uint64_t)row + 1, key);
}

template <int U>
__global__ void __launch_bounds__(256)
coo_rowaligned_vec_kernel(const uint32_t *__restrict__ rowIdxs, const uint32_t *__restrict__ colIdxs,
                          const float *__restrict__ vals, uint32_t M, uint32_t nnz, uint32_t chunk,
                          const float *__restrict__ B, uint32_t N, size_t ldb,
                          float *__restrict__ C, size_t ldc) {
    const uint32_t lane = lane_id();
    const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t s_nom64 = (uint64_t)w * chunk;
    if (s_nom64 >= nnz && w != 0) return;
    const uint32_t s_nom = (uint32_t)s_nom64;
    const uint32_t e_nom = (uint64_t)s_nom + chunk >= nnz ? nnz : s_nom + chunk;
    const uint32_t a = next_row_start(rowIdxs, nnz, s_nom);
    const uint32_t e = next_row_start(rowIdxs, nnz, e_nom);
    if (a >= nnz && nnz != 0) return;                 // the previous warp already covers up to M
    uint32_t next_out = (a == 0) ? 0u : __ldg(rowIdxs + a - 1) + 1u;   // first row this warp must write
    const uint32_t rhi = (e >= nnz) ? M : __ldg(rowIdxs + e - 1) + 1u;

    const uint32_t col0 = blockIdx.y * (128u * U) + lane * 4u;
    bool valid[U];
#pragma unroll
    for (int u = 0; u < U; ++u) valid[u] = (col0 + u * 128u) < N;

    float4 acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    auto store_row = [&](uint32_t r, bool zero) {
        float *crow = C + (size_t)r * ldc + col0;
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (valid[u]) __stcs(reinterpret_cast<float4 *>(crow + u * 128), zero ? zero4 : acc[u]);
    };

    uint32_t cur = kPad;   // row being accumulated
    for (uint32_t base = a; base < e; base += 32) {
        const uint32_t idx = base + lane;
        uint32_t mr = 0, mc = 0;
        float mv = 0.f;
        if (idx < e) { mr = ld_stream(rowIdxs + idx); mc = ld_stream(colIdxs + idx); mv = ld_stream(vals + idx); }
        const int cnt = min(32u, e - base);
        for (int j = 0; j < cnt; ++j) {
            const uint32_t r = __shfl_sync(0xFFFFFFFFu, mr, j);
            if (r != cur) {
                if (cur != kPad) { store_row(cur, false); next_out = cur + 1; }
                for (; next_out < r; ++next_out) store_row(next_out, true);   // rows without entries
                cur = r;
#pragma unroll
                for (int u = 0; u < U; ++u) acc[u] = zero4;
            }
            const uint32_t c = __shfl_sync(0xFFFFFFFFu, mc, j);
            const float v = __shfl_sync(0xFFFFFFFFu, mv, j);
            const float *brow = B + (size_t)c * ldb + col0;
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (valid[u]) fma4(acc[u], v, __ldg(reinterpret_cast<const float4 *>(brow + u * 128)));
        }
    }
    if (cur != kPad) { store_row(cur, false); next_out = cur + 1; }
    for (; next_out < rhi; ++next_out) store_row(next_out, true);
}

// any N / alignment
template <int U>
__global__ void __launch_bounds__(256)
coo_rowaligned_scalar_kernel(const uint32_t *__restrict__ rowIdxs, const uint32_t *__restrict__ colIdxs,
                             const float *__restrict__ vals, uint32_t M, uint32_t nnz, uint32_t chunk,
                             const float *__restrict__ B, uint32_t N, size_t ldb,
                             float *__restrict__ C, size_t ldc) {
    const uint32_t lane = lane_id();
    const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint64_t s_nom64 = (uint64_t)w * chunk;
    if (s_nom64 >= nnz && w != 0) return;
    const uint32_t s_nom = (uint32_t)s_nom64;
    const uint32_t e_nom = (uint64_t)s_nom + chunk >= nnz ? nnz : s_nom + chunk;
    const uint32_t a = next_row_start(rowIdxs, nnz, s_nom);
    const uint32_t e = next_row_start(rowIdxs, nnz, e_nom);
    if (a >= nnz && nnz != 0) return;
    uint32_t next_out = (a == 0) ? 0u : __ldg(rowIdxs + a - 1) + 1u;
    const uint32_t rhi = (e >= nnz) ? M : __ldg(rowIdxs + e - 1) + 1u;
    const uint32_t col0 = blockIdx.y * (32u * U) + lane;

    float acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = 0.f;
    auto store_row = [&](uint32_t r, bool zero) {
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (col0 + u * 32u < N) C[(size_t)r * ldc + col0 + u * 32u] = zero ? 0.f : acc[u];
    };
    uint32_t cur = kPad;
    for (uint32_t base = a; base < e; base += 32) {
        const uint32_t idx = base + lane;
        uint32_t mr = 0, mc = 0;
        float mv = 0.f;
        if (idx < e) { mr = ld_stream(rowIdxs + idx); mc = ld_stream(colIdxs + idx); mv = ld_stream(vals + idx); }
        const int cnt = min(32u, e - base);
        for (int j = 0; j < cnt; ++j) {
            const uint32_t r = __shfl_sync(0xFFFFFFFFu, mr, j);
            if (r != cur) {
                if (cur != kPad) { store_row(cur, false); next_out = cur + 1; }
                for (; next_out < r; ++next_out) store_row(next_out, true);
                cur = r;
#pragma unroll
                for (int u = 0; u < U; ++u) acc[u] = 0.f;
            }
            const uint32_t c = __shfl_sync(0xFFFFFFFFu, mc, j);
            const float v = __shfl_sync(0xFFFFFFFFu, mv, j);
            const float *brow = B + (size_t)c * ldb;
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (col0 + u * 32u < N) acc[u] = fmaf(v, __ldg(brow + col0 + u * 32u), acc[u]);
        }
    }
    if (cur != kPad) { store_row(cur, false); next_out = cur + 1; }
    for (; next_out < rhi; ++next_out) store_row(next_out, true);
}

static int spmm_coo_dispatch(const uint32_t *rowIdxs, const uint32_t *colIdxs, const float *vals,
                             uint32_t M, uint32_t K, uint32_t nnz, const float *B, uint32_t N, size_t ldb,
                             float *C, size_t ldc, int variant, void *ws, size_t ws_bytes, cudaStream_t st) {
    CUSPMM_REQUIRE(variant >= 0 && variant <= CUSPMM_COO_NUM_VARIANTS, "COO variant %d does not exist", variant);
    CUSPMM_REQUIRE(ldb >= N && ldc >= N, "ldb/ldc must be >= N");
    if (M == 0 || N == 0) return CUSPMM_OK;
    CUSPMM_REQUIRE(B && C && (nnz == 0 || (rowIdxs && colIdxs && vals)), "null operand pointer");
    const bool vok = (N % 4 == 0) && (ldb % 4 == 0) && (ldc % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    if (variant == 0) {   // the CSR selector decides whether the staged kernel pays; it needs row pointers
        const bool have_ws = ws && ws_bytes >= (size_t)(M + 1) * 4;
        const int cv = csr_select_variant(M, K, nnz, N, vok);
        variant = (have_ws && (cv == 3 || cv == 5 || cv == 8)) ? 2 : 1;
    }
    if (variant == 2) {
        if (ws_bytes < (size_t)(M + 1) * 4 || !ws)
            return set_error(CUSPMM_ERR_WORKSPACE, "COO variant 2 needs %zu bytes of workspace, got %zu",
                             (size_t)(M + 1) * 4, ws_bytes);
        uint32_t *rowPtrs = static_cast<uint32_t *>(ws);
        int rc = coo_to_csr_rowptrs(rowIdxs, M, nnz, rowPtrs, st);
        if (rc) return rc;
        return spmm_csr_dispatch(rowPtrs, colIdxs, vals, M, K, nnz, B, N, ldb, C, ldc, 0, st);
    }
    // chunk: a multiple of 32 entries, enough warps to fill the machine
    uint64_t want = (uint64_t)sm_count() * 8 * 8;
    uint32_t chunk = (uint32_t)(((uint64_t)nnz + want - 1) / want);
    chunk = ((chunk + 31) / 32) * 32;
    if (chunk == 0) chunk = 32;
    uint32_t warps = (uint32_t)(((uint64_t)nnz + chunk - 1) / chunk);
    if (warps == 0) warps = 1;
    const uint32_t blocks = (warps + 7) / 8;
    if (vok) {
        if (N > 256) {
            dim3 grid(blocks, (N + 511) / 512);
            coo_rowaligned_vec_kernel<4><<<grid, 256, 0, st>>>(rowIdxs, colIdxs, vals, M, nnz, chunk, B, N, ldb, C, ldc);
        } else if (N > 128) {
            coo_rowaligned_vec_kernel<2><<<dim3(blocks, 1), 256, 0, st>>>(rowIdxs, colIdxs, vals, M, nnz, chunk, B, N, ldb, C, ldc);
        } else {
            coo_rowaligned_vec_kernel<1><<<dim3(blocks, 1), 256, 0, st>>>(rowIdxs, colIdxs, vals, M, nnz, chunk, B, N, ldb, C, ldc);
        }
        CUSPMM_LAUNCH_CHECK("coo_rowaligned_vec_kernel");
    } else {
        dim3 grid(blocks, (N + 127) / 128);
        coo_rowaligned_scalar_kernel<4><<<grid, 256, 0, st>>>(rowIdxs, colIdxs, vals, M, nnz, chunk, B, N, ldb, C, ldc);
        CUSPMM_LAUNCH_CHECK("coo_rowaligned_scalar_kernel");
    }
    return CUSPMM_OK;
}

} // namespace cuspmm_b200

extern "C" size_t cuspmm_spmm_coo_workspace(uint32_t M, uint32_t nnz, uint32_t N, int variant) {
    (void)nnz; (void)N;
    return (variant == 1) ? 0 : (size_t)(M + 1) * sizeof(uint32_t);
}

extern "C" int cuspmm_spmm_coo(const uint32_t *rowIdxs, const uint32_t *colIdxs, const float *vals,
                               uint32_t M, uint32_t K, uint32_t nnz, const float *B, uint32_t N, size_t ldb,
                               float *C, size_t ldc, int variant, void *ws, size_t ws_bytes, void *stream) {
    return cuspmm_b200::spmm_coo_dispatch(rowIdxs, colIdxs, vals, M, K, nnz, B, N, ldb, C, ldc, variant, ws,
                                          ws_bytes, cuspmm_b200::as_stream(stream));
}
