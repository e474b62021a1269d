// engine/engine_csr.hpp -- EngineCSR: kernel-number dispatch for the CSR path.
// Same interface as the reference's include/engine/engine_csr.hpp (spmmCSRCpu, spmmCSRWrapper<k>,
// EngineCSR{MataT, MatbT, SUPPORT_CUSPARSE, fmt, dirPath, seqTime, logSeq, report, runKernel}).
// Every GPU wrapper multiplies through the C ABI (include/cuspmm_b200.h); there is no CPU fallback.
#pragma once

#include "commons.hpp"
#include "engine/cusparse.hpp"
#include "engine/engine_base.hpp"
#include "formats/dense.hpp"
#include "formats/sparse_csr.hpp"
#include "spmm_cusparse.hpp"

namespace cuspmm {

// kernel 0: the reference's host SpMM (in-process checker)
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmCSRCpu(SparseMatrixCSR<DT, MT> *ma, DenseMatrix<DT, MT> *mb, DenseMatrix<DT, MT> *mc);
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmCSRWrapper1(SparseMatrixCSR<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref);
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmCSRWrapper2(SparseMatrixCSR<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref);
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmCSRWrapper3(SparseMatrixCSR<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref);
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmCSRWrapper4(SparseMatrixCSR<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref);
// additive (the reference stops at 4): the staged kernel with part of every B chunk in tensor memory
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmCSRWrapper5(SparseMatrixCSR<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref);
// additive: equal nnz ranges per warp, rows cut at range boundaries, ordered carry fix-up (few / skewed rows)
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmCSRWrapper6(SparseMatrixCSR<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref);
// additive: every B read from tensor memory (columns split over the TMEM lane quarters)
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmCSRWrapper7(SparseMatrixCSR<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref);
// additive: tensor cores -- tiles of A made dense in shared memory, tcgen05.mma with a tf32 + bf16 three-product split
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmCSRWrapper8(SparseMatrixCSR<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref);

template <typename DT, typename MT, typename AccT>
class EngineCSR : public EngineBase {
  public:
    using MataT = SparseMatrixCSR<DT, MT>;
    using MatbT = DenseMatrix<DT, MT>;

    bool SUPPORT_CUSPARSE = true;
    std::string fmt;
    std::string dirPath;
    double seqTime = 1.f;

    explicit EngineCSR(std::string dirPath) {
        this->numKernels = CUSPMM_CSR_NUM_VARIANTS;   // 8: the reference's four slots + the four additive kernels
        this->dirPath = dirPath;
        this->fmt = "CSR";
    }

    void logSeq(double seq) { this->seqTime = seq; }

    void report(MataT *a, MatbT *b, int num, double pro, double kernel, double epilog, bool correct) {
        reportTime(this->dirPath, a->numRows, a->numCols, a->numNonZero, this->fmt, b->ordering, num, pro, kernel, epilog, correct);
    }

    void *runKernel(int num, void *_ma, void *_mb, void *_mc) override {
        auto ma = reinterpret_cast<MataT *>(_ma);
        auto mb = reinterpret_cast<MatbT *>(_mb);
        auto mc = reinterpret_cast<MatbT *>(_mc);
        if (num == 0) return spmmCSRCpu<DT, MT, AccT>(ma, mb, mc);
        if (num == 1) return spmmCSRWrapper1<DT, MT, AccT>(ma, mb, mc);
        if (num == 2) return spmmCSRWrapper2<DT, MT, AccT>(ma, mb, mc);
        if (num == 3) return spmmCSRWrapper3<DT, MT, AccT>(ma, mb, mc);
        if (num == 4) return spmmCSRWrapper4<DT, MT, AccT>(ma, mb, mc);
        if (num == 5) return spmmCSRWrapper5<DT, MT, AccT>(ma, mb, mc);
        if (num == 6) return spmmCSRWrapper6<DT, MT, AccT>(ma, mb, mc);
        if (num == 7) return spmmCSRWrapper7<DT, MT, AccT>(ma, mb, mc);
        if (num == 8) return spmmCSRWrapper8<DT, MT, AccT>(ma, mb, mc);
        if (num == -1) return spmmCSRWrapper3<DT, MT, AccT>(ma, mb, mc);
        throw std::runtime_error("Not implemented");
    }
};

}  // namespace cuspmm
