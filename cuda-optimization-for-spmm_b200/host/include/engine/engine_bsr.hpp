// engine/engine_bsr.hpp -- EngineBSR: kernel-number dispatch for the BSR path.
// Same interface as the reference's include/engine/engine_bsr.hpp (spmmBSRCpu, spmmBSRWrapper<k>,
// EngineBSR{MataT, MatbT, SUPPORT_CUSPARSE, fmt, dirPath, seqTime, logSeq, report, runKernel}).
// Every GPU wrapper multiplies through the C ABI (include/cuspmm_b200.h); there is no CPU fallback.
#pragma once

#include "commons.hpp"
#include "engine/cusparse.hpp"
#include "engine/engine_base.hpp"
#include "formats/dense.hpp"
#include "formats/sparse_bsr.hpp"
#include "spmm_cusparse.hpp"

namespace cuspmm {

// kernel 0: the reference's host SpMM (in-process checker)
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmBSRCpu(SparseMatrixBSR<DT, MT> *ma, DenseMatrix<DT, MT> *mb, DenseMatrix<DT, MT> *mc);
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmBSRWrapper1(SparseMatrixBSR<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref);
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmBSRWrapper2(SparseMatrixBSR<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref);
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmBSRWrapper3(SparseMatrixBSR<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref);

template <typename DT, typename MT, typename AccT>
class EngineBSR : public EngineBase {
  public:
    using MataT = SparseMatrixBSR<DT, MT>;
    using MatbT = DenseMatrix<DT, MT>;

    bool SUPPORT_CUSPARSE = false;
    std::string fmt;
    std::string dirPath;
    double seqTime = 1.f;

    explicit EngineBSR(std::string dirPath) {
        this->numKernels = 3;
        this->dirPath = dirPath;
        this->fmt = "BSR";
    }

    void logSeq(double seq) { this->seqTime = seq; }

    void report(MataT *a, MatbT *b, int num, double pro, double kernel, double epilog, bool correct) {
        reportTime(this->dirPath, a->numRows, a->numCols, a->numNonZero, this->fmt, b->ordering, num, pro, kernel, epilog, correct);
    }

    void *runKernel(int num, void *_ma, void *_mb, void *_mc) override {
        auto ma = reinterpret_cast<MataT *>(_ma);
        auto mb = reinterpret_cast<MatbT *>(_mb);
        auto mc = reinterpret_cast<MatbT *>(_mc);
        if (num == 0) return spmmBSRCpu<DT, MT, AccT>(ma, mb, mc);
        if (num == 1) return spmmBSRWrapper1<DT, MT, AccT>(ma, mb, mc);
        if (num == 2) return spmmBSRWrapper2<DT, MT, AccT>(ma, mb, mc);
        if (num == 3) return spmmBSRWrapper3<DT, MT, AccT>(ma, mb, mc);
        if (num == -1) return spmmBSRWrapper1<DT, MT, AccT>(ma, mb, mc);
        throw std::runtime_error("Not implemented");
    }
};

}  // namespace cuspmm
