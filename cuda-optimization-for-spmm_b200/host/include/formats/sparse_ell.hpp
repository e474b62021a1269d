// formats/sparse_ell.hpp -- ELL storage.
// SparseMatrixELL is the reference's column-ELL (include/formats/sparse_ell.hpp:12-37):
// rowIdxs / data are [numCols x maxColNnz], padding row index -1 (0xFFFFFFFF in MT), read from
// the *_rowind.ell / *_values_colmajor.ell pair (src/formats/sparse_ell.cu:13-55).
// SlicedELL is the engine's native device layout (DESIGN.md "Sliced ELL"), produced on the device
// by SparseMatrixELL::toSliced() (column-ELL -> CSR -> sliced ELL through the C ABI).
#pragma once

#include "commons.hpp"
#include "cuda_utils.hpp"
#include "formats/dense.hpp"
#include "formats/matrix.hpp"

namespace cuspmm {

template <typename DT, typename MT>
struct SlicedELL {      // device resident
    MT numRows = 0, numCols = 0, numNonZero = 0, sliceHeight = 32, numSlices = 0, numSlots = 0;
    MT *slicePtrs = nullptr, *colIdxs = nullptr;
    DT *data = nullptr;
    ~SlicedELL() {
        if (slicePtrs) cudaFree(slicePtrs);
        if (colIdxs) cudaFree(colIdxs);
        if (data) cudaFree(data);
    }
};

template <typename _dataT, typename _metaT>
class SparseMatrixELL : public SparseMatrix<_dataT, _metaT> {
  public:
    using DT = _dataT;
    using MT = _metaT;
    MT *rowIdxs = nullptr;
    MT maxColNnz = 0;

    SparseMatrixELL() = default;
    SparseMatrixELL(std::string rowindPath, std::string valuesPath);
    SparseMatrixELL(MT numRows, MT numCols, MT numNonZero, MT maxColNnz, bool onDevice);
    ~SparseMatrixELL() override;

    void setCusparseSpMatDesc(cusparseSpMatDescr_t *matDescP) override;
    cusparseSpMMAlg_t getCusparseAlg() override;
    bool allocateSpace(bool onDevice);
    SparseMatrixELL<DT, MT> *copy2Device();
    DenseMatrix<DT, MT> *toDense();
    // device conversion (this must be on the device): column-ELL -> sliced ELL
    SlicedELL<DT, MT> *toSliced();
};

}  // namespace cuspmm
