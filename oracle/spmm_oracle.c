/*
 * oracle/spmm_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's four host SpMM functions and the four
 * toDense() converters.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may call into this file; the product
 * path (cuda-optimization-for-spmm_b200/csrc) never does.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference).  Arithmetic is kept operation-for-operation identical:
 *   - CSR : fp32 product, fp64 running sum in CSR order, one rounding to fp32
 *           on store                       (src/spmm/csr/spmm_csr.cpp:15-27)
 *   - COO / ELL / BSR : fp32 product, fp32 "+=" into a zeroed C, traversal
 *           order exactly as written       (src/spmm/coo/spmm_coo.cpp:16-24,
 *            src/spmm/ell/spmm_ell.cpp:15-28, src/spmm/bsr/spmm_bsr.cpp:17-39)
 * Build with -ffp-contract=off and without -mfma so no product is fused into
 * the following add (the reference is built by CMake with no optimisation
 * flags on x86-64, i.e. unfused SSE2 arithmetic).
 *
 * Parity pinning: tests/test_oracle_golden.py checks these functions against
 * the reference's committed fixtures (data/small_10x10, data/small_32x32:
 * result.expect, coo.out, coo_cuda.out) and against oracle/_ref (the
 * reference's own spmm_*.cpp compiled from /root/reference) bit for bit.
 *
 * The *_omp variants run the same per-row arithmetic with OpenMP over rows
 * (each C row is still produced by one thread in the same order, so results
 * are bit-identical to the serial ones); they exist only so that the CPU
 * baseline can be quoted on all host cores.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define EXPORT __attribute__((visibility("default")))

EXPORT int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ CSR --- */
/* src/spmm/csr/spmm_csr.cpp:15-27.  C is overwritten (":25  mc->data[..] = acc"). */
static inline void csr_row(uint32_t r, uint32_t N, const uint32_t *rowPtrs,
                           const uint32_t *colIdxs, const float *vals,
                           const float *B, float *C) {
    uint32_t row_start = rowPtrs[r];
    uint32_t row_end = rowPtrs[r + 1];
    for (uint32_t c = 0; c < N; c++) {
        double acc = 0.f;
        for (uint32_t i = row_start; i < row_end; i++) {
            uint32_t k = colIdxs[i];
            float prod = vals[i] * B[(size_t)k * N + c]; /* float*float, rounded to fp32 */
            acc += prod;                                  /* widened, summed in fp64 */
        }
        C[(size_t)r * N + c] = (float)acc;
    }
}

EXPORT void oracle_spmm_csr(uint32_t M, uint32_t K, uint32_t N,
                            const uint32_t *rowPtrs, const uint32_t *colIdxs,
                            const float *vals, const float *B, float *C) {
    (void)K;
    for (uint32_t r = 0; r < M; r++) csr_row(r, N, rowPtrs, colIdxs, vals, B, C);
}

/* Same rows [r0, r1) only: bounded CPU-baseline samples of a large workload. */
EXPORT void oracle_spmm_csr_rows(uint32_t r0, uint32_t r1, uint32_t N,
                                 const uint32_t *rowPtrs, const uint32_t *colIdxs,
                                 const float *vals, const float *B, float *C) {
    for (uint32_t r = r0; r < r1; r++) csr_row(r, N, rowPtrs, colIdxs, vals, B, C);
}

EXPORT void oracle_spmm_csr_rows_omp(uint32_t r0, uint32_t r1, uint32_t N,
                                     const uint32_t *rowPtrs, const uint32_t *colIdxs,
                                     const float *vals, const float *B, float *C) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t r = r0; r < (int64_t)r1; r++)
        csr_row((uint32_t)r, N, rowPtrs, colIdxs, vals, B, C);
}

/* ------------------------------------------------------------------ COO --- */
/* src/spmm/coo/spmm_coo.cpp:16-24.  Needs a zeroed C; accumulates in fp32 in
 * file order (AccT is unused by the reference). */
EXPORT void oracle_spmm_coo(uint32_t M, uint32_t K, uint32_t N, uint32_t nnz,
                            const uint32_t *rowIdxs, const uint32_t *colIdxs,
                            const float *vals, const float *B, float *C) {
    (void)M; (void)K;
    for (uint32_t idx = 0; idx < nnz; idx++) {
        uint32_t r = rowIdxs[idx];
        uint32_t c = colIdxs[idx];
        float value = vals[idx];
        float *crow = C + (size_t)r * N;
        const float *brow = B + (size_t)c * N;
        for (uint32_t j = 0; j < N; j++) {
            float prod = value * brow[j];
            crow[j] = crow[j] + prod;
        }
    }
}

/* ------------------------------------------------------------------ ELL --- */
/* src/spmm/ell/spmm_ell.cpp:15-28.  Column-ELL: rowIdxs/vals are
 * [numCols x maxColNnz]; padding row index is -1, parsed into uint32 and read
 * back through "int row" (:17), skipped when negative (:20). */
EXPORT void oracle_spmm_ell(uint32_t M, uint32_t K, uint32_t N, uint32_t maxColNnz,
                            const uint32_t *rowIdxs, const float *vals,
                            const float *B, float *C) {
    (void)M;
    for (uint32_t col = 0; col < K; col++) {
        for (uint32_t s = 0; s < maxColNnz; s++) {
            int row = (int)rowIdxs[(size_t)col * maxColNnz + s];
            float value = vals[(size_t)col * maxColNnz + s];
            if (row >= 0) {
                float *crow = C + (size_t)row * N;
                const float *brow = B + (size_t)col * N;
                for (uint32_t j = 0; j < N; j++) {
                    float prod = value * brow[j];
                    crow[j] = crow[j] + prod;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------ BSR --- */
/* src/spmm/bsr/spmm_bsr.cpp:17-39.  Blocks are row-major with stride
 * blockColSize (:33); zeros stored inside a block are multiplied too. */
static inline void bsr_block_row(uint32_t blockRow, uint32_t N, uint32_t br, uint32_t bc,
                                 const uint32_t *blockRowPtrs, const uint32_t *blockColIdxs,
                                 const float *blocks, const float *B, float *C) {
    uint32_t start = blockRowPtrs[blockRow];
    uint32_t end = blockRowPtrs[blockRow + 1];
    for (uint32_t b = start; b < end; b++) {
        uint32_t blockCol = blockColIdxs[b];
        const float *blk = blocks + (size_t)br * bc * b;
        uint32_t r0 = blockRow * br;
        uint32_t c0 = blockCol * bc;
        for (uint32_t ar = r0; ar < r0 + br; ar++) {
            for (uint32_t ac = c0; ac < c0 + bc; ac++) {
                float a = blk[(ar - r0) * bc + (ac - c0)];
                float *crow = C + (size_t)ar * N;
                const float *brow = B + (size_t)ac * N;
                for (uint32_t j = 0; j < N; j++) {
                    float prod = a * brow[j];
                    crow[j] = crow[j] + prod;
                }
            }
        }
    }
}

EXPORT void oracle_spmm_bsr(uint32_t numBlockRows, uint32_t N, uint32_t br, uint32_t bc,
                            const uint32_t *blockRowPtrs, const uint32_t *blockColIdxs,
                            const float *blocks, const float *B, float *C) {
    for (uint32_t R = 0; R < numBlockRows; R++)
        bsr_block_row(R, N, br, bc, blockRowPtrs, blockColIdxs, blocks, B, C);
}

EXPORT void oracle_spmm_bsr_omp(uint32_t numBlockRows, uint32_t N, uint32_t br, uint32_t bc,
                                const uint32_t *blockRowPtrs, const uint32_t *blockColIdxs,
                                const float *blocks, const float *B, float *C) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t R = 0; R < (int64_t)numBlockRows; R++)
        bsr_block_row((uint32_t)R, N, br, bc, blockRowPtrs, blockColIdxs, blocks, B, C);
}

/* -------------------------------------------------------------- toDense --- */
/* src/formats/sparse_csr.cu:163-180 */
EXPORT void oracle_csr_to_dense(uint32_t M, uint32_t K, const uint32_t *rowPtrs,
                                const uint32_t *colIdxs, const float *vals, float *D) {
    memset(D, 0, (size_t)M * K * sizeof(float));
    for (uint32_t r = 0; r < M; r++)
        for (uint32_t i = rowPtrs[r]; i < rowPtrs[r + 1]; i++)
            D[(size_t)r * K + colIdxs[i]] = vals[i];
}

/* src/formats/sparse_coo.cu:153-168 */
EXPORT void oracle_coo_to_dense(uint32_t M, uint32_t K, uint32_t nnz, const uint32_t *rowIdxs,
                                const uint32_t *colIdxs, const float *vals, float *D) {
    memset(D, 0, (size_t)M * K * sizeof(float));
    for (uint32_t i = 0; i < nnz; i++)
        D[(size_t)rowIdxs[i] * K + colIdxs[i]] = vals[i];
}

/* src/formats/sparse_ell.cu:161-178 */
EXPORT void oracle_ell_to_dense(uint32_t M, uint32_t K, uint32_t maxColNnz,
                                const uint32_t *rowIdxs, const float *vals, float *D) {
    memset(D, 0, (size_t)M * K * sizeof(float));
    for (uint32_t col = 0; col < K; col++) {
        size_t base = (size_t)col * maxColNnz;
        for (uint32_t s = 0; s < maxColNnz; s++) {
            int row = (int)rowIdxs[base + s];
            if (row >= 0) D[(size_t)row * K + col] = vals[base + s];
        }
    }
}

/* src/formats/sparse_bsr.cu:297-326.  The reference indexes the block with
 * stride blockRowSize (:318) where spmmBSRCpu uses blockColSize; they agree
 * for the square blocks the reference instantiates.  This restatement uses
 * blockColSize, i.e. the layout the SpMM itself reads. */
EXPORT void oracle_bsr_to_dense(uint32_t M, uint32_t K, uint32_t br, uint32_t bc,
                                const uint32_t *blockRowPtrs, const uint32_t *blockColIdxs,
                                const float *blocks, float *D) {
    memset(D, 0, (size_t)M * K * sizeof(float));
    uint32_t numBlockRows = M / br;
    for (uint32_t R = 0; R < numBlockRows; R++)
        for (uint32_t b = blockRowPtrs[R]; b < blockRowPtrs[R + 1]; b++) {
            const float *blk = blocks + (size_t)br * bc * b;
            uint32_t r0 = R * br, c0 = blockColIdxs[b] * bc;
            for (uint32_t r = 0; r < br; r++)
                for (uint32_t c = 0; c < bc; c++)
                    D[(size_t)(r0 + r) * K + (c0 + c)] = blk[r * bc + c];
        }
}

/* ------------------------------------------------------------- checking --- */
/* torch::allclose as the reference's wrappers call it
 * (src/spmm/csr/spmm_csr_k1.cu:75-78, include/utils.hpp:10-11):
 * |a - b| <= atol + rtol * |b| for every element (NaNs never close). */
EXPORT int oracle_allclose(const float *a, const float *b, size_t n, float rtol, float atol) {
    for (size_t i = 0; i < n; i++) {
        float d = a[i] - b[i];
        if (d < 0) d = -d;
        float t = b[i] < 0 ? -b[i] : b[i];
        if (!(d <= atol + rtol * t)) return 0;
    }
    return 1;
}

/* (|A|.|B|) row panel [r0, r1): the component-wise error denominator used by
 * the parity tests (max_ij |C - Cref|_ij / (|A||B|)_ij). */
EXPORT void oracle_absprod_csr_rows(uint32_t r0, uint32_t r1, uint32_t N,
                                    const uint32_t *rowPtrs, const uint32_t *colIdxs,
                                    const float *vals, const float *B, double *S) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t r = r0; r < (int64_t)r1; r++) {
        double *srow = S + (size_t)(r - r0) * N;
        for (uint32_t c = 0; c < N; c++) srow[c] = 0.0;
        for (uint32_t i = rowPtrs[r]; i < rowPtrs[r + 1]; i++) {
            double a = vals[i] < 0 ? -(double)vals[i] : (double)vals[i];
            const float *brow = B + (size_t)colIdxs[i] * N;
            for (uint32_t c = 0; c < N; c++) {
                double b = brow[c] < 0 ? -(double)brow[c] : (double)brow[c];
                srow[c] += a * b;
            }
        }
    }
}

/* -------------------------------------------------------------- parsing --- */
/* The reference's readers pull numbers with operator>> (src/formats/dense.cu:27-34,
 * sparse_csr.cu:30-50, sparse_coo.cu:34-36, sparse_ell.cu:39-51, sparse_bsr.cu:44-60).
 * For float that is a correctly rounded decimal->binary32 conversion (strtof);
 * for uint32_t a leading '-' negates in the unsigned type, so the ELL padding
 * "-1" becomes 0xFFFFFFFF.  These helpers give the Python side of the oracle
 * the same conversions (Python's float() would round through binary64 first). */
#include <stdlib.h>
#include <ctype.h>

EXPORT size_t oracle_parse_f32(const char *text, float *out, size_t maxn) {
    size_t n = 0;
    const char *p = text;
    while (n < maxn) {
        char *end;
        float v = strtof(p, &end);
        if (end == p) break;
        out[n++] = v;
        p = end;
    }
    return n;
}

EXPORT size_t oracle_parse_u32(const char *text, uint32_t *out, size_t maxn) {
    size_t n = 0;
    const char *p = text;
    while (n < maxn) {
        char *end;
        while (isspace((unsigned char)*p)) p++;
        if (!*p) break;
        int neg = (*p == '-');
        unsigned long long v = strtoull(neg ? p + 1 : p, &end, 10);
        if (end == (neg ? p + 1 : p)) break;
        out[n++] = neg ? (uint32_t)(0u - (uint32_t)v) : (uint32_t)v;
        p = end;
    }
    return n;
}
