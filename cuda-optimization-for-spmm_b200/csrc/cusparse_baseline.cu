// cusparse_baseline.cu -- same-run vendor baseline (north_star (d)).
// Restates cusparseTest (src/engine/cusparse.cu:10-57) with the reference's algorithm choices
// (CUSPARSE_SPMM_CSR_ALG2, src/formats/sparse_csr.cu:183-185; CUSPARSE_SPMM_COO_ALG4,
// src/formats/sparse_coo.cu:98-100) but row-major B/C always, CUDA-event timing, and handle /
// descriptor / buffer creation and cusparseSpMM_preprocess outside the timed region.
#include "common.cuh"

#include <cusparse.h>
#include <float.h>

#define CUSPMM_CUSPARSE(call)                                                                      \
    do {                                                                                           \
        cusparseStatus_t s__ = (call);                                                             \
        if (s__ != CUSPARSE_STATUS_SUCCESS) {                                                      \
            rc = ::cuspmm_b200::set_error(CUSPMM_ERR_CUSPARSE, "%s failed: %s", #call,            \
                                          cusparseGetErrorString(s__));                            \
            goto done;                                                                             \
        }                                                                                          \
    } while (0)

extern "C" int cuspmm_cusparse_spmm(int fmt, const uint32_t *rowOrPtr, const uint32_t *colIdxs, const float *vals,
                                    uint32_t M, uint32_t K, uint32_t nnz, const float *B, uint32_t N, float *C,
                                    int alg, int warmup, int iters, float *avg_ms, float *min_ms) {
    using namespace cuspmm_b200;
    CUSPMM_REQUIRE(fmt == 0 || fmt == 1, "fmt must be 0 (CSR) or 1 (COO)");
    CUSPMM_REQUIRE(iters >= 1 && warmup >= 0, "bad iteration counts");
    int rc = CUSPMM_OK;
    cusparseHandle_t handle = nullptr;
    cusparseSpMatDescr_t matA = nullptr;
    cusparseDnMatDescr_t matB = nullptr, matC = nullptr;
    void *buffer = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    size_t bufBytes = 0;
    const float alpha = 1.0f, beta = 0.0f;
    cusparseSpMMAlg_t a = alg ? (cusparseSpMMAlg_t)alg : (fmt == 0 ? CUSPARSE_SPMM_CSR_ALG2 : CUSPARSE_SPMM_COO_ALG4);
    float total = 0.f, best = FLT_MAX;

    CUSPMM_CUSPARSE(cusparseCreate(&handle));
    if (fmt == 0)
        CUSPMM_CUSPARSE(cusparseCreateCsr(&matA, M, K, nnz, (void *)rowOrPtr, (void *)colIdxs, (void *)vals,
                                          CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_32F));
    else
        CUSPMM_CUSPARSE(cusparseCreateCoo(&matA, M, K, nnz, (void *)rowOrPtr, (void *)colIdxs, (void *)vals,
                                          CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO, CUDA_R_32F));
    CUSPMM_CUSPARSE(cusparseCreateDnMat(&matB, K, N, N, (void *)B, CUDA_R_32F, CUSPARSE_ORDER_ROW));
    CUSPMM_CUSPARSE(cusparseCreateDnMat(&matC, M, N, N, (void *)C, CUDA_R_32F, CUSPARSE_ORDER_ROW));
    CUSPMM_CUSPARSE(cusparseSpMM_bufferSize(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_NON_TRANSPOSE,
                                            &alpha, matA, matB, &beta, matC, CUDA_R_32F, a, &bufBytes));
    if (cudaMalloc(&buffer, bufBytes ? bufBytes : 1) != cudaSuccess) {
        rc = set_error(CUSPMM_ERR_CUDA, "cudaMalloc of the cuSPARSE buffer (%zu bytes) failed", bufBytes);
        goto done;
    }
    CUSPMM_CUSPARSE(cusparseSpMM_preprocess(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_NON_TRANSPOSE,
                                            &alpha, matA, matB, &beta, matC, CUDA_R_32F, a, buffer));
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < warmup + iters; ++i) {
        if (i >= warmup) cudaEventRecord(e0, 0);
        CUSPMM_CUSPARSE(cusparseSpMM(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_NON_TRANSPOSE,
                                     &alpha, matA, matB, &beta, matC, CUDA_R_32F, a, buffer));
        if (i >= warmup) {
            cudaEventRecord(e1, 0);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            total += ms;
            if (ms < best) best = ms;
        }
    }
    if (cudaDeviceSynchronize() != cudaSuccess) {
        rc = set_error(CUSPMM_ERR_CUDA, "cuSPARSE SpMM left a CUDA error: %s", cudaGetErrorString(cudaGetLastError()));
        goto done;
    }
    if (avg_ms) *avg_ms = total / iters;
    if (min_ms) *min_ms = best;
done:
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (buffer) cudaFree(buffer);
    if (matA) cusparseDestroySpMat(matA);
    if (matB) cusparseDestroyDnMat(matB);
    if (matC) cusparseDestroyDnMat(matC);
    if (handle) cusparseDestroy(handle);
    return rc;
}

// cuSPARSE BSR SpMM (fp32 blocks): the baseline the reference wires up but never runs
// (cusparseCreateBsr in src/formats/sparse_bsr.cu:139-155, SUPPORT_CUSPARSE = false in engine_bsr.hpp:24).
extern "C" int cuspmm_cusparse_spmm_bsr(const uint32_t *blockRowPtrs, const uint32_t *blockColIdxs, const float *blocks,
                                        uint32_t numBlockRows, uint32_t numBlockCols, uint32_t numBlocks, uint32_t bs,
                                        const float *B, uint32_t N, float *C, int warmup, int iters, float *avg_ms,
                                        float *min_ms) {
    using namespace cuspmm_b200;
    CUSPMM_REQUIRE(iters >= 1 && warmup >= 0 && bs >= 1, "bad arguments");
    int rc = CUSPMM_OK;
    cusparseHandle_t handle = nullptr;
    cusparseSpMatDescr_t matA = nullptr;
    cusparseDnMatDescr_t matB = nullptr, matC = nullptr;
    void *buffer = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    size_t bufBytes = 0;
    const float alpha = 1.0f, beta = 0.0f;
    const cusparseSpMMAlg_t a = CUSPARSE_SPMM_ALG_DEFAULT;
    float total = 0.f, best = FLT_MAX;
    const int64_t M = (int64_t)numBlockRows * bs, K = (int64_t)numBlockCols * bs;

    CUSPMM_CUSPARSE(cusparseCreate(&handle));
    CUSPMM_CUSPARSE(cusparseCreateBsr(&matA, numBlockRows, numBlockCols, numBlocks, bs, bs, (void *)blockRowPtrs,
                                      (void *)blockColIdxs, (void *)blocks, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I,
                                      CUSPARSE_INDEX_BASE_ZERO, CUDA_R_32F, CUSPARSE_ORDER_ROW));
    CUSPMM_CUSPARSE(cusparseCreateDnMat(&matB, K, N, N, (void *)B, CUDA_R_32F, CUSPARSE_ORDER_ROW));
    CUSPMM_CUSPARSE(cusparseCreateDnMat(&matC, M, N, N, (void *)C, CUDA_R_32F, CUSPARSE_ORDER_ROW));
    CUSPMM_CUSPARSE(cusparseSpMM_bufferSize(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_NON_TRANSPOSE,
                                            &alpha, matA, matB, &beta, matC, CUDA_R_32F, a, &bufBytes));
    if (cudaMalloc(&buffer, bufBytes ? bufBytes : 1) != cudaSuccess) {
        rc = set_error(CUSPMM_ERR_CUDA, "cudaMalloc of the cuSPARSE buffer (%zu bytes) failed", bufBytes);
        goto done;
    }
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < warmup + iters; ++i) {
        if (i >= warmup) cudaEventRecord(e0, 0);
        CUSPMM_CUSPARSE(cusparseSpMM(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_NON_TRANSPOSE,
                                     &alpha, matA, matB, &beta, matC, CUDA_R_32F, a, buffer));
        if (i >= warmup) {
            cudaEventRecord(e1, 0);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            total += ms;
            if (ms < best) best = ms;
        }
    }
    if (cudaDeviceSynchronize() != cudaSuccess) {
        rc = set_error(CUSPMM_ERR_CUDA, "cuSPARSE BSR SpMM left a CUDA error: %s", cudaGetErrorString(cudaGetLastError()));
        goto done;
    }
    if (avg_ms) *avg_ms = total / iters;
    if (min_ms) *min_ms = best;
done:
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (buffer) cudaFree(buffer);
    if (matA) cusparseDestroySpMat(matA);
    if (matB) cusparseDestroyDnMat(matB);
    if (matC) cusparseDestroyDnMat(matC);
    if (handle) cusparseDestroy(handle);
    return rc;
}

// ------------------------------------------------------------------ cuSPARSE Blocked-ELL baseline
// the ELL descriptor the reference leaves unimplemented (src/formats/sparse_ell.cu:92-105 throws).
// cuSPARSE's only ELL flavour is Blocked-ELL: every block row stores the same number of bs x bs blocks (ellCols / bs),
// column index -1 = padding.  The BSR operand is re-laid on the device (pad to the longest block row, values as a dense
// (rows x ellCols) row-major array), which is outside the timed region like the descriptor / buffer / preprocess work.
namespace cuspmm_b200 {
__global__ void blockedell_width_kernel(const uint32_t *__restrict__ blockRowPtrs, uint32_t numBlockRows, unsigned int *maxw) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t w = r < numBlockRows ? blockRowPtrs[r + 1] - blockRowPtrs[r] : 0u;
    w = __reduce_max_sync(0xFFFFFFFFu, w);
    if (lane_id() == 0 && w) atomicMax(maxw, w);
}
// one thread per (block row, slot, element): ellVals[(R * bs + i) * ellCols + s * bs + c] = blocks[b][i][c] (0 in padding slots)
__global__ void blockedell_fill_kernel(const uint32_t *__restrict__ blockRowPtrs, const uint32_t *__restrict__ blockColIdxs,
                                       const float *__restrict__ blocks, uint32_t numBlockRows, uint32_t bs, uint32_t width,
                                       int *__restrict__ ellCols, float *__restrict__ ellVals) {
    const uint64_t per = (uint64_t)bs * bs;
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)numBlockRows * width * per) return;
    const uint32_t e = (uint32_t)(idx % per), s = (uint32_t)((idx / per) % width), R = (uint32_t)(idx / (per * width));
    const uint32_t i = e / bs, c = e % bs;
    const uint32_t b0 = blockRowPtrs[R], cnt = blockRowPtrs[R + 1] - b0;
    const bool live = s < cnt;
    if (e == 0) ellCols[(size_t)R * width + s] = live ? (int)blockColIdxs[b0 + s] : -1;
    ellVals[((size_t)R * bs + i) * ((size_t)width * bs) + (size_t)s * bs + c] = live ? blocks[(size_t)(b0 + s) * per + e] : 0.f;
}
} // namespace cuspmm_b200

extern "C" int cuspmm_cusparse_spmm_blockedell(const uint32_t *blockRowPtrs, const uint32_t *blockColIdxs, const float *blocks,
                                               uint32_t numBlockRows, uint32_t numBlockCols, uint32_t numBlocks, uint32_t bs,
                                               const float *B, uint32_t N, float *C, int warmup, int iters, float *avg_ms,
                                               float *min_ms, uint32_t *ell_width_blocks) {
    using namespace cuspmm_b200;
    CUSPMM_REQUIRE(iters >= 1 && warmup >= 0 && bs >= 1 && blockRowPtrs, "bad arguments");
    (void)numBlocks;
    int rc = CUSPMM_OK;
    cusparseHandle_t handle = nullptr;
    cusparseSpMatDescr_t matA = nullptr;
    cusparseDnMatDescr_t matB = nullptr, matC = nullptr;
    void *buffer = nullptr;
    unsigned int *dMax = nullptr;
    int *ellCols = nullptr;
    float *ellVals = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    size_t bufBytes = 0;
    const float alpha = 1.0f, beta = 0.0f;
    const cusparseSpMMAlg_t a = CUSPARSE_SPMM_BLOCKED_ELL_ALG1;
    float total = 0.f, best = FLT_MAX;
    const int64_t M = (int64_t)numBlockRows * bs, K = (int64_t)numBlockCols * bs;
    unsigned int width = 0;

    if (cudaMalloc(&dMax, 4) != cudaSuccess || cudaMemset(dMax, 0, 4) != cudaSuccess) {
        rc = set_error(CUSPMM_ERR_CUDA, "cudaMalloc failed");
        goto done;
    }
    blockedell_width_kernel<<<(numBlockRows + 255) / 256, 256>>>(blockRowPtrs, numBlockRows, dMax);
    if (cudaMemcpy(&width, dMax, 4, cudaMemcpyDeviceToHost) != cudaSuccess) {
        rc = set_error(CUSPMM_ERR_CUDA, "Blocked-ELL width kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
        goto done;
    }
    if (width == 0) width = 1;
    if (ell_width_blocks) *ell_width_blocks = width;
    {
        const size_t nvals = (size_t)numBlockRows * width * bs * bs;
        if (cudaMalloc(&ellCols, (size_t)numBlockRows * width * sizeof(int)) != cudaSuccess ||
            cudaMalloc(&ellVals, nvals * sizeof(float)) != cudaSuccess) {
            rc = set_error(CUSPMM_ERR_CUDA, "cudaMalloc of the Blocked-ELL operand (%zu values) failed", nvals);
            goto done;
        }
        blockedell_fill_kernel<<<(unsigned)((nvals + 255) / 256), 256>>>(blockRowPtrs, blockColIdxs, blocks, numBlockRows, bs, width,
                                                                        ellCols, ellVals);
        if (cudaDeviceSynchronize() != cudaSuccess) {
            rc = set_error(CUSPMM_ERR_CUDA, "Blocked-ELL fill kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
            goto done;
        }
        count_launch(2);
    }
    CUSPMM_CUSPARSE(cusparseCreate(&handle));
    CUSPMM_CUSPARSE(cusparseCreateBlockedEll(&matA, M, K, bs, (int64_t)width * bs, ellCols, ellVals, CUSPARSE_INDEX_32I,
                                             CUSPARSE_INDEX_BASE_ZERO, CUDA_R_32F));
    CUSPMM_CUSPARSE(cusparseCreateDnMat(&matB, K, N, N, (void *)B, CUDA_R_32F, CUSPARSE_ORDER_ROW));
    CUSPMM_CUSPARSE(cusparseCreateDnMat(&matC, M, N, N, (void *)C, CUDA_R_32F, CUSPARSE_ORDER_ROW));
    CUSPMM_CUSPARSE(cusparseSpMM_bufferSize(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_NON_TRANSPOSE,
                                            &alpha, matA, matB, &beta, matC, CUDA_R_32F, a, &bufBytes));
    if (cudaMalloc(&buffer, bufBytes ? bufBytes : 1) != cudaSuccess) {
        rc = set_error(CUSPMM_ERR_CUDA, "cudaMalloc of the cuSPARSE buffer (%zu bytes) failed", bufBytes);
        goto done;
    }
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < warmup + iters; ++i) {
        if (i >= warmup) cudaEventRecord(e0, 0);
        CUSPMM_CUSPARSE(cusparseSpMM(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_NON_TRANSPOSE,
                                     &alpha, matA, matB, &beta, matC, CUDA_R_32F, a, buffer));
        if (i >= warmup) {
            cudaEventRecord(e1, 0);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            total += ms;
            if (ms < best) best = ms;
        }
    }
    if (cudaDeviceSynchronize() != cudaSuccess) {
        rc = set_error(CUSPMM_ERR_CUDA, "cuSPARSE Blocked-ELL SpMM left a CUDA error: %s", cudaGetErrorString(cudaGetLastError()));
        goto done;
    }
    if (avg_ms) *avg_ms = total / iters;
    if (min_ms) *min_ms = best;
done:
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (buffer) cudaFree(buffer);
    if (dMax) cudaFree(dMax);
    if (ellCols) cudaFree(ellCols);
    if (ellVals) cudaFree(ellVals);
    if (matA) cusparseDestroySpMat(matA);
    if (matB) cusparseDestroyDnMat(matB);
    if (matC) cusparseDestroyDnMat(matC);
    if (handle) cusparseDestroy(handle);
    return rc;
}
