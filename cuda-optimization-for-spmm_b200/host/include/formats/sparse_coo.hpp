// formats/sparse_coo.hpp -- COO storage (reference: include/formats/sparse_coo.hpp:12-39).
// rowIdxs / colIdxs / data [nnz], sorted by (row, col) as convert_mtx.py:181-185 writes them.
#pragma once

#include "commons.hpp"
#include "cuda_utils.hpp"
#include "formats/dense.hpp"
#include "formats/matrix.hpp"

namespace cuspmm {

template <typename _dataT, typename _metaT>
class SparseMatrixCOO : public SparseMatrix<_dataT, _metaT> {
  public:
    using DT = _dataT;
    using MT = _metaT;
    MT *rowIdxs = nullptr;
    MT *colIdxs = nullptr;

    SparseMatrixCOO() = default;
    explicit SparseMatrixCOO(std::string filePath);
    SparseMatrixCOO(MT numRows, MT numCols, MT numNonZero, bool onDevice);
    ~SparseMatrixCOO() override;

    void setCusparseSpMatDesc(cusparseSpMatDescr_t *matDescP) override;
    cusparseSpMMAlg_t getCusparseAlg() override;
    SparseMatrixCOO<DT, MT> *copy2Device();
    bool allocateSpace(bool onDevice);
    DenseMatrix<DT, MT> *toDense();
};

}  // namespace cuspmm
