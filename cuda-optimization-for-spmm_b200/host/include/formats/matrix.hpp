// formats/matrix.hpp -- base classes of the storage formats.
// Mirrors include/formats/matrix.hpp:10-47 of the reference: enum ORDERING, Matrix<DT,MT>
// {numRows, numCols, onDevice}, SparseMatrix<DT,MT> {numNonZero, data, cuSPARSE virtuals}.
#pragma once

#include <cstdint>
#include <cstring>

#include "cuda_utils.hpp"
#include "spmm_cusparse.hpp"

namespace cuspmm {

enum ORDERING {
    ROW_MAJOR,
    COL_MAJOR,
};

template <typename _dataT, typename _metaT>
class Matrix {
  public:
    using DT = _dataT;
    using MT = _metaT;
    MT numRows = 0, numCols = 0;
    bool onDevice = false;
    Matrix() = default;
    virtual ~Matrix() = default;
};

template <typename _dataT, typename _metaT>
class SparseMatrix : public Matrix<_dataT, _metaT> {
  public:
    using DT = _dataT;
    using MT = _metaT;
    MT numNonZero = 0;
    DT *data = nullptr;
    SparseMatrix() = default;

    virtual void setCusparseSpMatDesc(cusparseSpMatDescr_t *matDescP) = 0;
    virtual cusparseSpMMAlg_t getCusparseAlg() = 0;
};

// pinned-host / device allocation shared by every format (reference: allocateSpace() of each class:
// cudaMallocHost or cudaMalloc followed by a memset to 0)
template <typename T>
inline T *allocZeroed(size_t count, bool onDevice) {
    T *p = nullptr;
    const size_t bytes = (count ? count : 1) * sizeof(T);
    if (onDevice) {
        cudaCheckError(cudaMalloc(&p, bytes));
        cudaCheckError(cudaMemset(p, 0, bytes));
    } else {
        cudaCheckError(cudaMallocHost(&p, bytes));
        std::memset(p, 0, bytes);
    }
    return p;
}
template <typename T>
inline void freeSpaceOf(T *&p, bool onDevice) {
    if (!p) return;
    if (onDevice) { cudaCheckError(cudaFree(p)); } else { cudaCheckError(cudaFreeHost(p)); }
    p = nullptr;
}

}  // namespace cuspmm
