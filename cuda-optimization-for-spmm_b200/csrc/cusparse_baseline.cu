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
