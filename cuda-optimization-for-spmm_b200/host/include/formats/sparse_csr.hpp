// formats/sparse_csr.hpp -- CSR storage (reference: include/formats/sparse_csr.hpp:12-39).
// Public fields rowPtrs[numRows+1], colIdxs[nnz], data[nnz]; file ctor reads the 4-line text
// format (src/formats/sparse_csr.cu:12-51).  Index arrays are copied with sizeof(MT) (the
// reference uses sizeof(DT), :119-126, harmless only because both are 4 bytes).
#pragma once

#include "commons.hpp"
#include "cuda_utils.hpp"
#include "formats/dense.hpp"
#include "formats/matrix.hpp"

namespace cuspmm {

template <typename _dataT, typename _metaT>
class SparseMatrixCSR : public SparseMatrix<_dataT, _metaT> {
  public:
    using DT = _dataT;
    using MT = _metaT;
    MT *rowPtrs = nullptr;
    MT *colIdxs = nullptr;

    SparseMatrixCSR() = default;
    explicit SparseMatrixCSR(std::string filePath);
    SparseMatrixCSR(MT numRows, MT numCols, MT numNonZero, bool onDevice);
    ~SparseMatrixCSR() override;

    void setCusparseSpMatDesc(cusparseSpMatDescr_t *matDescP) override;
    cusparseSpMMAlg_t getCusparseAlg() override;
    SparseMatrixCSR<DT, MT> *copy2Device();
    SparseMatrixCSR<DT, MT> *copy2Host();
    bool allocateSpace(bool onDevice);
    DenseMatrix<DT, MT> *toDense();

    template <typename U, typename M>
    friend std::ostream &operator<<(std::ostream &out, SparseMatrixCSR<U, M> &m);
};

}  // namespace cuspmm
