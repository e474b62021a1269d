// engine/engine_ell.hpp -- EngineELL: kernel-number dispatch for the ELL path.
// Same interface as the reference's include/engine/engine_ell.hpp (spmmELLCpu, spmmELLWrapper<k>,
// EngineELL{MataT, MatbT, SUPPORT_CUSPARSE, fmt, dirPath, seqTime, logSeq, report, runKernel}).
// Every GPU wrapper multiplies through the C ABI (include/cuspmm_b200.h); there is no CPU fallback.
#pragma once

#include "commons.hpp"
#include "engine/cusparse.hpp"
#include "engine/engine_base.hpp"
#include "formats/dense.hpp"
#include "formats/sparse_ell.hpp"
#include "spmm_cusparse.hpp"

namespace cuspmm {

// kernel 0: the reference's host SpMM (in-process checker)
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmELLCpu(SparseMatrixELL<DT, MT> *ma, DenseMatrix<DT, MT> *mb, DenseMatrix<DT, MT> *mc);
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmELLWrapper1(SparseMatrixELL<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref);
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmELLWrapper2(SparseMatrixELL<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref);
// 3 slice per CTA, 4 staged + tensor-memory operand path, 5 every B read from tensor memory (C-ABI ELL variants 3..5)
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmELLWrapper3(SparseMatrixELL<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref);
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmELLWrapper4(SparseMatrixELL<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref);
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmELLWrapper5(SparseMatrixELL<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref);
// additive: tensor cores (the CSR tensor kernel reading the sliced layout)
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmELLWrapper6(SparseMatrixELL<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref);

template <typename DT, typename MT, typename AccT>
class EngineELL : public EngineBase {
  public:
    using MataT = SparseMatrixELL<DT, MT>;
    using MatbT = DenseMatrix<DT, MT>;

    bool SUPPORT_CUSPARSE = false;
    std::string fmt;
    std::string dirPath;
    double seqTime = 1.f;

    explicit EngineELL(std::string dirPath) {
        this->numKernels = CUSPMM_ELL_NUM_VARIANTS;   // 6: every C-ABI ELL variant (the reference ships Wrapper2 but sets 1, engine_ell.hpp:32)
        this->dirPath = dirPath;
        this->fmt = "ELL";
    }

    void logSeq(double seq) { this->seqTime = seq; }

    void report(MataT *a, MatbT *b, int num, double pro, double kernel, double epilog, bool correct) {
        reportTime(this->dirPath, a->numRows, a->numCols, a->numNonZero, this->fmt, b->ordering, num, pro, kernel, epilog, correct);
    }

    void *runKernel(int num, void *_ma, void *_mb, void *_mc) override {
        auto ma = reinterpret_cast<MataT *>(_ma);
        auto mb = reinterpret_cast<MatbT *>(_mb);
        auto mc = reinterpret_cast<MatbT *>(_mc);
        if (num == 0) return spmmELLCpu<DT, MT, AccT>(ma, mb, mc);
        if (num == 1) return spmmELLWrapper1<DT, MT, AccT>(ma, mb, mc);
        if (num == 2) return spmmELLWrapper2<DT, MT, AccT>(ma, mb, mc);
        if (num == 3) return spmmELLWrapper3<DT, MT, AccT>(ma, mb, mc);
        if (num == 4) return spmmELLWrapper4<DT, MT, AccT>(ma, mb, mc);
        if (num == 5) return spmmELLWrapper5<DT, MT, AccT>(ma, mb, mc);
        if (num == 6) return spmmELLWrapper6<DT, MT, AccT>(ma, mb, mc);
        if (num == -1) return spmmELLWrapper1<DT, MT, AccT>(ma, mb, mc);
        throw std::runtime_error("Not implemented");
    }
};

}  // namespace cuspmm
