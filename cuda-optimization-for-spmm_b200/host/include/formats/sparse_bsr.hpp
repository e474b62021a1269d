// formats/sparse_bsr.hpp -- BSR storage (reference: include/formats/sparse_bsr.hpp:13-57).
// blockRowPtrs[numBlockRows+1], blockColIdxs[numBlocks], data[numBlocks*br*bc] row-major inside a
// block; numBlockRows = numRows / blockRowSize (src/formats/sparse_bsr.cu:34).  New: fromCSR()
// builds the BSR on the DEVICE (the reference's fromDense() throws, sparse_bsr.cu:259).
#pragma once

#include "commons.hpp"
#include "cuda_utils.hpp"
#include "formats/dense.hpp"
#include "formats/matrix.hpp"
#include "formats/sparse_csr.hpp"

namespace cuspmm {

template <typename _dataT, typename _metaT>
class SparseMatrixBSR : public SparseMatrix<_dataT, _metaT> {
  public:
    using DT = _dataT;
    using MT = _metaT;
    MT blockRowSize = 0;
    MT blockColSize = 0;
    MT numBlocks = 0;
    MT *blockRowPtrs = nullptr;
    MT *blockColIdxs = nullptr;
    MT numBlockRows = 0;
    MT numElements = 0;

    SparseMatrixBSR() = default;
    explicit SparseMatrixBSR(std::string filePath);
    SparseMatrixBSR(MT numRows, MT numCols, MT numNonZero, MT blockRowSize, MT blockColSize, MT numBlocks, bool onDevice);
    SparseMatrixBSR(SparseMatrixBSR<DT, MT> *target, bool onDevice);
    ~SparseMatrixBSR() override;

    void setCusparseSpMatDesc(cusparseSpMatDescr_t *matDescP) override;
    cusparseSpMMAlg_t getCusparseAlg() override;
    bool copyData(SparseMatrixBSR<DT, MT> *source, bool onDevice);
    SparseMatrixBSR<DT, MT> *copy2Device();
    SparseMatrixBSR<DT, MT> *copy2Host();
    void assertCheck();
    void assertSameShape(SparseMatrixBSR<DT, MT> *target);
    bool allocateSpace(bool onDevice);
    // reference signature kept; implemented (host) instead of throwing
    SparseMatrixBSR<DT, MT> *fromDense(DenseMatrix<DT, MT> *dense, MT blockRowSize, MT blockColSize);
    // device conversion: `csr` must be on the device; rows/cols are zero-padded to block multiples
    static SparseMatrixBSR<DT, MT> *fromCSR(SparseMatrixCSR<DT, MT> *csr, MT blockRowSize, MT blockColSize);
    DenseMatrix<DT, MT> *toDense();
};

}  // namespace cuspmm
