/*
 * oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Glue that lets the reference's OWN host SpMM translation units
 *     /root/reference/src/spmm/csr/spmm_csr.cpp   (spmmCSRCpu)
 *     /root/reference/src/spmm/coo/spmm_coo.cpp   (spmmCOOCpu)
 *     /root/reference/src/spmm/ell/spmm_ell.cpp   (spmmELLCpu)
 *     /root/reference/src/spmm/bsr/spmm_bsr.cpp   (spmmBSRCpu)
 * be compiled, unmodified and from where they lie, into oracle/_ref/libref_spmm.so
 * (recipe: oracle/Makefile, target `ref`).  No reference source is copied.
 *
 * Those four files only touch public fields of the storage classes declared in
 * /root/reference/include/formats/*.hpp.  The classes' member functions live in
 * src/formats/*.cu and allocate with cudaMallocHost, which needs a CUDA driver
 * even for the CPU path, so they are NOT linked.  Instead this shim supplies
 * host-only stand-ins for the handful of members the linker needs
 * (constructors, destructors, the two cuSPARSE virtuals, and
 * DenseMatrix::toOrdering, which is only reached for COL_MAJOR operands and
 * aborts here) and views the caller's buffers through the reference's classes.
 *
 * The exported C functions below take exactly the arrays the oracle takes, so
 * tests can compare oracle/spmm_oracle.c against the reference's own compiled
 * arithmetic bit for bit, and bench.py can time the reference's own loops.
 */
#include "formats/dense.hpp"
#include "formats/sparse_bsr.hpp"
#include "formats/sparse_coo.hpp"
#include "formats/sparse_csr.hpp"
#include "formats/sparse_ell.hpp"

#include <cstdint>
#include <cstdlib>

namespace cuspmm {

// ---- declarations of the reference's templates (include/engine/engine_*.hpp) ----
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmCSRCpu(SparseMatrixCSR<DT, MT> *, DenseMatrix<DT, MT> *, DenseMatrix<DT, MT> *);
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmCOOCpu(SparseMatrixCOO<DT, MT> *, DenseMatrix<DT, MT> *, DenseMatrix<DT, MT> *);
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmELLCpu(SparseMatrixELL<DT, MT> *, DenseMatrix<DT, MT> *, DenseMatrix<DT, MT> *);
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmBSRCpu(SparseMatrixBSR<DT, MT> *, DenseMatrix<DT, MT> *, DenseMatrix<DT, MT> *);

// ---- host-only stand-ins for members defined in src/formats/*.cu ----
template <> DenseMatrix<float, uint32_t>::~DenseMatrix() {}
template <> bool DenseMatrix<float, uint32_t>::toOrdering(ORDERING) {
    std::fprintf(stderr, "ref_shim: toOrdering() is not available without CUDA; pass ROW_MAJOR B\n");
    std::abort();
}

template <> SparseMatrixCSR<float, uint32_t>::SparseMatrixCSR() : SparseMatrix<float, uint32_t>() {
    rowPtrs = nullptr; colIdxs = nullptr;
}
template <> SparseMatrixCSR<float, uint32_t>::~SparseMatrixCSR() {}
template <> void SparseMatrixCSR<float, uint32_t>::setCusparseSpMatDesc(cusparseSpMatDescr_t *) { std::abort(); }
template <> cusparseSpMMAlg_t SparseMatrixCSR<float, uint32_t>::getCusparseAlg() { return CUSPARSE_SPMM_CSR_ALG2; }

template <> SparseMatrixCOO<float, uint32_t>::SparseMatrixCOO() : SparseMatrix<float, uint32_t>() {
    rowIdxs = nullptr; colIdxs = nullptr;
}
template <> SparseMatrixCOO<float, uint32_t>::~SparseMatrixCOO() {}
template <> void SparseMatrixCOO<float, uint32_t>::setCusparseSpMatDesc(cusparseSpMatDescr_t *) { std::abort(); }
template <> cusparseSpMMAlg_t SparseMatrixCOO<float, uint32_t>::getCusparseAlg() { return CUSPARSE_SPMM_COO_ALG4; }

template <> SparseMatrixELL<float, uint32_t>::SparseMatrixELL() : SparseMatrix<float, uint32_t>() {
    rowIdxs = nullptr; maxColNnz = 0;
}
template <> SparseMatrixELL<float, uint32_t>::~SparseMatrixELL() {}
template <> void SparseMatrixELL<float, uint32_t>::setCusparseSpMatDesc(cusparseSpMatDescr_t *) { std::abort(); }
template <> cusparseSpMMAlg_t SparseMatrixELL<float, uint32_t>::getCusparseAlg() { return CUSPARSE_SPMM_ALG_DEFAULT; }

template <> SparseMatrixBSR<float, uint32_t>::SparseMatrixBSR() : SparseMatrix<float, uint32_t>() {
    blockRowSize = blockColSize = numBlocks = numBlockRows = numElements = 0;
    blockRowPtrs = nullptr; blockColIdxs = nullptr;
}
template <> SparseMatrixBSR<float, uint32_t>::~SparseMatrixBSR() {}
template <> void SparseMatrixBSR<float, uint32_t>::setCusparseSpMatDesc(cusparseSpMatDescr_t *) { std::abort(); }
template <> cusparseSpMMAlg_t SparseMatrixBSR<float, uint32_t>::getCusparseAlg() { return CUSPARSE_SPMM_ALG_DEFAULT; }

} // namespace cuspmm

using namespace cuspmm;
typedef DenseMatrix<float, uint32_t> Dn;

static void viewDense(Dn &d, uint32_t rows, uint32_t cols, float *p) {
    d.numRows = rows; d.numCols = cols; d.onDevice = false;
    d.ordering = ORDERING::ROW_MAJOR; d.data = p;
}

extern "C" {

__attribute__((visibility("default")))
void ref_spmm_csr(uint32_t M, uint32_t K, uint32_t N, uint32_t nnz, const uint32_t *rowPtrs,
                  const uint32_t *colIdxs, const float *vals, const float *B, float *C) {
    SparseMatrixCSR<float, uint32_t> a;
    a.numRows = M; a.numCols = K; a.numNonZero = nnz; a.onDevice = false;
    a.rowPtrs = const_cast<uint32_t *>(rowPtrs); a.colIdxs = const_cast<uint32_t *>(colIdxs);
    a.data = const_cast<float *>(vals);
    Dn b, c; viewDense(b, K, N, const_cast<float *>(B)); viewDense(c, M, N, C);
    spmmCSRCpu<float, uint32_t, double>(&a, &b, &c);
    a.rowPtrs = a.colIdxs = nullptr; a.data = nullptr; b.data = c.data = nullptr;
}

__attribute__((visibility("default")))
void ref_spmm_coo(uint32_t M, uint32_t K, uint32_t N, uint32_t nnz, const uint32_t *rowIdxs,
                  const uint32_t *colIdxs, const float *vals, const float *B, float *C) {
    SparseMatrixCOO<float, uint32_t> a;
    a.numRows = M; a.numCols = K; a.numNonZero = nnz; a.onDevice = false;
    a.rowIdxs = const_cast<uint32_t *>(rowIdxs); a.colIdxs = const_cast<uint32_t *>(colIdxs);
    a.data = const_cast<float *>(vals);
    Dn b, c; viewDense(b, K, N, const_cast<float *>(B)); viewDense(c, M, N, C);
    spmmCOOCpu<float, uint32_t, double>(&a, &b, &c);
    a.rowIdxs = a.colIdxs = nullptr; a.data = nullptr; b.data = c.data = nullptr;
}

__attribute__((visibility("default")))
void ref_spmm_ell(uint32_t M, uint32_t K, uint32_t N, uint32_t nnz, uint32_t maxColNnz,
                  const uint32_t *rowIdxs, const float *vals, const float *B, float *C) {
    SparseMatrixELL<float, uint32_t> a;
    a.numRows = M; a.numCols = K; a.numNonZero = nnz; a.onDevice = false; a.maxColNnz = maxColNnz;
    a.rowIdxs = const_cast<uint32_t *>(rowIdxs); a.data = const_cast<float *>(vals);
    Dn b, c; viewDense(b, K, N, const_cast<float *>(B)); viewDense(c, M, N, C);
    spmmELLCpu<float, uint32_t, double>(&a, &b, &c);
    a.rowIdxs = nullptr; a.data = nullptr; b.data = c.data = nullptr;
}

__attribute__((visibility("default")))
void ref_spmm_bsr(uint32_t M, uint32_t K, uint32_t N, uint32_t br, uint32_t bc, uint32_t numBlocks,
                  const uint32_t *blockRowPtrs, const uint32_t *blockColIdxs, const float *blocks,
                  const float *B, float *C) {
    SparseMatrixBSR<float, uint32_t> a;
    a.numRows = M; a.numCols = K; a.onDevice = false;
    a.blockRowSize = br; a.blockColSize = bc; a.numBlocks = numBlocks;
    a.numBlockRows = M / br;                       // sparse_bsr.cu:34
    a.numElements = numBlocks * br * bc; a.numNonZero = a.numElements;
    a.blockRowPtrs = const_cast<uint32_t *>(blockRowPtrs);
    a.blockColIdxs = const_cast<uint32_t *>(blockColIdxs);
    a.data = const_cast<float *>(blocks);
    Dn b, c; viewDense(b, K, N, const_cast<float *>(B)); viewDense(c, M, N, C);
    spmmBSRCpu<float, uint32_t, double>(&a, &b, &c);
    a.blockRowPtrs = a.blockColIdxs = nullptr; a.data = nullptr; b.data = c.data = nullptr;
}

} // extern "C"
