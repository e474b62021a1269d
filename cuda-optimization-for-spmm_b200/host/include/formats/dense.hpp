// formats/dense.hpp -- DenseMatrix<DT,MT>: B and C of the SpMM.
// Same public interface as the reference's include/formats/dense.hpp:18-52 (fields `data`,
// `ordering`; file / shape / copy constructors; copyData, copy2Device, copy2Host, toOrdering,
// save2File, allocateSpace, freeSpace, setCusparseDnMatDesc).  Differences that matter:
//   * sizes are computed in size_t (the reference multiplies uint32 first, src/formats/dense.cu:146,238);
//   * toOrdering() of a DEVICE matrix transposes on the device (cuspmm_transpose_f32) instead of
//     D2H -> host double loop -> H2D (src/formats/dense.cu:140-191).
#pragma once

#include "commons.hpp"
#include "cuda_utils.hpp"
#include "formats/matrix.hpp"

namespace cuspmm {

template <typename _dataT, typename _metaT>
class DenseMatrix : public Matrix<_dataT, _metaT> {
  public:
    using DT = _dataT;
    using MT = _metaT;
    DT *data = nullptr;
    ORDERING ordering = ORDERING::ROW_MAJOR;

    DenseMatrix() = default;
    explicit DenseMatrix(std::string filePath);
    DenseMatrix(MT numRows, MT numCols, bool onDevice, ORDERING ordering = ORDERING::ROW_MAJOR);
    DenseMatrix(DenseMatrix<DT, MT> *source, bool onDevice);
    ~DenseMatrix() override { freeSpace(); }
    DenseMatrix(const DenseMatrix &) = delete;
    DenseMatrix &operator=(const DenseMatrix &) = delete;

    size_t numElements() const { return (size_t)this->numRows * (size_t)this->numCols; }
    bool copyData(DenseMatrix<DT, MT> *source);
    void setCusparseDnMatDesc(cusparseDnMatDescr_t *matDescP);
    void assertSameShape(DenseMatrix<DT, MT> *target);
    DenseMatrix<DT, MT> *copy2Device();
    DenseMatrix<DT, MT> *copy2Host();
    bool toOrdering(ORDERING newOrdering);
    bool save2File(std::string filePath);
    bool allocateSpace(bool onDevice);
    bool freeSpace();
};

}  // namespace cuspmm
