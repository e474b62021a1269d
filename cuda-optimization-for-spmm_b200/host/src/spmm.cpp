// spmm.cpp -- spmm<FMT>Cpu (kernel 0, the reference's in-process checker) and the GPU wrappers
// spmm<FMT>Wrapper<k> of the host layer.
//
// Wrapper contract = the reference's (e.g. src/spmm/csr/spmm_csr_k3.cu:59-105): a and b live on the
// device, `ref` is the HOST result of kernel 0; the wrapper allocates a device C ("prolog"), runs the
// kernel, copies C back ("epilog"), compares with ref using allclose(REL_TOL, ABS_TOL), prints one
// record and returns the device C (caller owns it).  What differs: the kernel is called through the
// C ABI, it is timed on the device with CUDA events over g_opts.iters launches after g_opts.warmup
// (cudaKernelTimeMs = average), B is never transposed on the host, and a variant that cannot run a
// shape prints a record with correct = 0 and returns nullptr (the reference does that for K4 only,
// spmm_csr_k4.cu:97-101).
//
// The CPU functions follow the reference's arithmetic exactly (CSR: fp32 product summed in AccT =
// double, src/spmm/csr/spmm_csr.cpp:15-27; COO/ELL/BSR: fp32 += into the zeroed C in storage order,
// spmm_coo.cpp:16-24, spmm_ell.cpp:15-28, spmm_bsr.cpp:17-39).  They are the checker, never a
// fallback: nothing here runs them in place of a GPU kernel.
#include "engine.hpp"

#include <functional>

namespace cuspmm {

// ------------------------------------------------------------------------------- kernel 0
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmCSRCpu(SparseMatrixCSR<DT, MT> *ma, DenseMatrix<DT, MT> *mb, DenseMatrix<DT, MT> *mc) {
    assert(!ma->onDevice && !mb->onDevice);
    if (mb->ordering == ORDERING::COL_MAJOR) mb->toOrdering(ORDERING::ROW_MAJOR);
    const size_t N = mb->numCols;
    for (MT r = 0; r < ma->numRows; ++r) {
        const MT lo = ma->rowPtrs[r], hi = ma->rowPtrs[r + 1];
        for (size_t c = 0; c < N; ++c) {
            AccT acc = 0;
            for (MT i = lo; i < hi; ++i) {
                const DT prod = ma->data[i] * mb->data[(size_t)ma->colIdxs[i] * N + c];
                acc += prod;
            }
            mc->data[(size_t)r * N + c] = (DT)acc;
        }
    }
    return mc;
}

template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmCOOCpu(SparseMatrixCOO<DT, MT> *ma, DenseMatrix<DT, MT> *mb, DenseMatrix<DT, MT> *mc) {
    assert(!ma->onDevice && !mb->onDevice);
    if (mb->ordering == ORDERING::COL_MAJOR) mb->toOrdering(ORDERING::ROW_MAJOR);
    const size_t N = mb->numCols;
    for (size_t i = 0; i < ma->numNonZero; ++i) {
        DT *crow = mc->data + (size_t)ma->rowIdxs[i] * N;
        const DT *brow = mb->data + (size_t)ma->colIdxs[i] * N;
        const DT v = ma->data[i];
        for (size_t j = 0; j < N; ++j) crow[j] += v * brow[j];
    }
    return mc;
}

template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmELLCpu(SparseMatrixELL<DT, MT> *ma, DenseMatrix<DT, MT> *mb, DenseMatrix<DT, MT> *mc) {
    assert(!ma->onDevice && !mb->onDevice);
    if (mb->ordering == ORDERING::COL_MAJOR) mb->toOrdering(ORDERING::ROW_MAJOR);
    const size_t N = mb->numCols;
    for (size_t col = 0; col < ma->numCols; ++col)
        for (size_t s = 0; s < ma->maxColNnz; ++s) {
            const int row = (int)ma->rowIdxs[col * ma->maxColNnz + s];
            if (row < 0) continue;
            const DT v = ma->data[col * ma->maxColNnz + s];
            DT *crow = mc->data + (size_t)row * N;
            const DT *brow = mb->data + col * N;
            for (size_t j = 0; j < N; ++j) crow[j] += v * brow[j];
        }
    return mc;
}

template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmBSRCpu(SparseMatrixBSR<DT, MT> *ma, DenseMatrix<DT, MT> *mb, DenseMatrix<DT, MT> *mc) {
    assert(!ma->onDevice && !mb->onDevice);
    if (mb->ordering == ORDERING::COL_MAJOR) mb->toOrdering(ORDERING::ROW_MAJOR);
    const size_t N = mb->numCols;
    const MT br = ma->blockRowSize, bc = ma->blockColSize;
    for (MT R = 0; R < ma->numBlockRows; ++R)
        for (MT b = ma->blockRowPtrs[R]; b < ma->blockRowPtrs[R + 1]; ++b) {
            const DT *blk = ma->data + (size_t)b * br * bc;
            for (MT i = 0; i < br; ++i)
                for (MT j = 0; j < bc; ++j) {
                    const DT v = blk[(size_t)i * bc + j];
                    DT *crow = mc->data + (size_t)(R * br + i) * N;
                    const DT *brow = mb->data + (size_t)(ma->blockColIdxs[b] * bc + j) * N;
                    for (size_t n = 0; n < N; ++n) crow[n] += v * brow[n];
                }
        }
    return mc;
}

// ------------------------------------------------------------------------------- shared wrapper body
namespace {
using Clock = std::chrono::high_resolution_clock;
inline double msSince(Clock::time_point t0) {
    return std::chrono::duration_cast<std::chrono::microseconds>(Clock::now() - t0).count() / 1000.0;
}

struct KernelSpec {
    std::string format, name;
    int kernelNum = 0;
    uint32_t M = 0, K = 0, nnz = 0;     // of A as the record reports it
    uint32_t outRows = 0;               // rows of C (BSR pads M up to a block multiple)
    double algBytes = 0;                // SURVEY.md section 8d formula for this format
    double flops = 0;
};

// prologFn: one-time device-side preparation counted as "prolog" (conversions, plans); returns false
// when the variant cannot run this shape.  launchFn(C): one multiply into device C, returns C-ABI status.
template <typename DT, typename MT>
DenseMatrix<DT, MT> *runWrapper(const KernelSpec &ks, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref,
                                const std::function<bool()> &prologFn,
                                const std::function<int(DenseMatrix<DT, MT> *)> &launchFn) {
    if (b->ordering == ORDERING::COL_MAJOR) b->toOrdering(ORDERING::ROW_MAJOR);   // on the device
    assert(b->onDevice);
    RecordExtra ex;
    ex.kernelName = ks.name;

    auto t1 = Clock::now();
    auto *c = new DenseMatrix<DT, MT>(ks.outRows, b->numCols, true, ORDERING::ROW_MAJOR);
    const bool ok = prologFn ? prologFn() : true;
    cudaCheckError(cudaDeviceSynchronize());
    const double pro = msSince(t1);
    if (!ok) {   // cannot run this shape: record with correct = 0, like the reference's K4 bail-out
        reportTime(testcase, ks.M, ks.K, ks.nnz, ks.format, b->ordering, ks.kernelNum, pro, 0, 0, false, &ex);
        delete c;
        return nullptr;
    }

    for (int i = 0; i < g_opts.warmup; ++i) cuspmmCheck(launchFn(c));
    cudaEvent_t e0, e1;
    cudaCheckError(cudaEventCreate(&e0));
    cudaCheckError(cudaEventCreate(&e1));
    const int iters = g_opts.iters > 0 ? g_opts.iters : 1;
    cudaCheckError(cudaEventRecord(e0, 0));
    for (int i = 0; i < iters; ++i) cuspmmCheck(launchFn(c));
    cudaCheckError(cudaEventRecord(e1, 0));
    cudaCheckError(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaCheckError(cudaEventElapsedTime(&ms, e0, e1));
    cudaCheckError(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double kernel = ms / iters;

    auto t3 = Clock::now();
    DenseMatrix<DT, MT> *res = c->copy2Host();
    const double epi = msSince(t3);

    const size_t n = std::min(res->numElements(), ref->numElements());
    const bool correct = allClose(res->data, ref->data, n, REL_TOL, ABS_TOL);
    double maxAbs = 0;
    for (size_t i = 0; i < n; ++i) maxAbs = std::max(maxAbs, std::fabs((double)res->data[i] - (double)ref->data[i]));
    delete res;

    ex.gflops = ks.flops / (kernel * 1e-3) / 1e9;
    ex.algBytes = ks.algBytes;
    ex.hbmGBs = ks.algBytes / (kernel * 1e-3) / 1e9;
    ex.hbmFrac = ex.hbmGBs / hbmPeakGBs();
    ex.maxAbsErr = maxAbs;     // max |C - Cref| (absolute), the quantity allclose bounds
    reportTime(testcase, ks.M, ks.K, ks.nnz, ks.format, b->ordering, ks.kernelNum, pro, kernel, epi, correct, &ex);
    return c;
}
}  // namespace

// ------------------------------------------------------------------------------- CSR wrappers
template <typename DT, typename MT, typename AccT>
static DenseMatrix<DT, MT> *csrWrapper(int k, const char *name, SparseMatrixCSR<DT, MT> *a, DenseMatrix<DT, MT> *b,
                                       DenseMatrix<DT, MT> *ref) {
    static_assert(std::is_same_v<DT, float> && std::is_same_v<MT, uint32_t>, "GPU kernels exist for <float, uint32_t>");
    assert(a->onDevice && b->onDevice);
    KernelSpec ks;
    ks.format = "CSR"; ks.name = name; ks.kernelNum = k;
    ks.M = a->numRows; ks.K = a->numCols; ks.nnz = a->numNonZero; ks.outRows = a->numRows;
    const double N = b->numCols;
    ks.algBytes = 8.0 * ks.nnz + 4.0 * (ks.M + 1.0) + 4.0 * ks.K * N + 4.0 * ks.M * N;
    ks.flops = 2.0 * ks.nnz * N;
    int probe = CUSPMM_OK;
    // variant 6 cuts rows between warps and needs carry slots: caller-provided workspace (0 bytes for the other variants)
    const size_t wsBytes = cuspmm_spmm_csr_workspace(a->numRows, a->numCols, a->numNonZero, b->numCols, k);
    void *ws = nullptr;
    if (wsBytes) cudaCheckError(cudaMalloc(&ws, wsBytes));
    auto launch = [&](DenseMatrix<DT, MT> *c) {
        return cuspmm_spmm_csr_ws(a->rowPtrs, a->colIdxs, a->data, a->numRows, a->numCols, a->numNonZero, b->data, b->numCols,
                                  b->numCols, c->data, c->numCols, k, ws, wsBytes, nullptr);
    };
    // shape support is decided by the library: a dry call into a scratch C tells us (status 1/3 = cannot run)
    auto prolog = [&]() {
        DenseMatrix<DT, MT> scratch(a->numRows, b->numCols, true);
        probe = launch(&scratch);
        return probe == CUSPMM_OK;
    };
    auto *res = runWrapper<DT, MT>(ks, b, ref, prolog, launch);
    if (ws) cudaCheckError(cudaFree(ws));
    return res;
}

#define CSR_WRAPPER(k, name)                                                                                            \
    template <typename DT, typename MT, typename AccT>                                                                  \
    DenseMatrix<DT, MT> *spmmCSRWrapper##k(SparseMatrixCSR<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref) { \
        return csrWrapper<DT, MT, AccT>(k, name, a, b, ref);                                                            \
    }
CSR_WRAPPER(1, "csr_rowsplit_vec")
CSR_WRAPPER(2, "csr_subwarp_vec")
CSR_WRAPPER(3, "csr_staged_tma")
CSR_WRAPPER(4, "csr_rowsplit_scalar")
CSR_WRAPPER(5, "csr_staged_tma_tmem")
CSR_WRAPPER(6, "csr_nnz_split_ordered_carry")
CSR_WRAPPER(7, "csr_all_tmem_quad")
CSR_WRAPPER(8, "csr_tensor_split")

// ------------------------------------------------------------------------------- COO wrappers
template <typename DT, typename MT, typename AccT>
static DenseMatrix<DT, MT> *cooWrapper(int k, const char *name, SparseMatrixCOO<DT, MT> *a, DenseMatrix<DT, MT> *b,
                                       DenseMatrix<DT, MT> *ref) {
    assert(a->onDevice && b->onDevice);
    KernelSpec ks;
    ks.format = "COO"; ks.name = name; ks.kernelNum = k;
    ks.M = a->numRows; ks.K = a->numCols; ks.nnz = a->numNonZero; ks.outRows = a->numRows;
    const double N = b->numCols;
    ks.algBytes = 12.0 * ks.nnz + 4.0 * ks.K * N + 4.0 * ks.M * N;
    ks.flops = 2.0 * ks.nnz * N;
    const size_t wsBytes = cuspmm_spmm_coo_workspace(a->numRows, a->numNonZero, b->numCols, k);
    void *ws = nullptr;
    auto launch = [&](DenseMatrix<DT, MT> *c) {
        return cuspmm_spmm_coo(a->rowIdxs, a->colIdxs, a->data, a->numRows, a->numCols, a->numNonZero, b->data, b->numCols,
                               b->numCols, c->data, c->numCols, k, ws, wsBytes, nullptr);
    };
    auto prolog = [&]() {
        if (wsBytes) cudaCheckError(cudaMalloc(&ws, wsBytes));
        DenseMatrix<DT, MT> scratch(a->numRows, b->numCols, true);
        return launch(&scratch) == CUSPMM_OK;
    };
    auto *c = runWrapper<DT, MT>(ks, b, ref, prolog, launch);
    if (ws) cudaFree(ws);
    return c;
}
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmCOOWrapper1(SparseMatrixCOO<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref) {
    return cooWrapper<DT, MT, AccT>(1, "coo_rowaligned_vec", a, b, ref);
}
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmCOOWrapper2(SparseMatrixCOO<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref) {
    return cooWrapper<DT, MT, AccT>(2, "coo_rowptr_then_csr_selector", a, b, ref);
}

// ------------------------------------------------------------------------------- ELL wrappers
template <typename DT, typename MT, typename AccT>
static DenseMatrix<DT, MT> *ellWrapper(int k, const char *name, SparseMatrixELL<DT, MT> *a, DenseMatrix<DT, MT> *b,
                                       DenseMatrix<DT, MT> *ref) {
    assert(a->onDevice && b->onDevice);
    KernelSpec ks;
    ks.format = "ELL"; ks.name = name; ks.kernelNum = k;
    ks.M = a->numRows; ks.K = a->numCols; ks.nnz = a->numNonZero; ks.outRows = a->numRows;
    const double N = b->numCols;
    ks.flops = 2.0 * ks.nnz * N;
    SlicedELL<DT, MT> *s = nullptr;
    auto launch = [&](DenseMatrix<DT, MT> *c) {
        return cuspmm_spmm_sell(s->slicePtrs, s->colIdxs, s->data, a->numRows, a->numCols, 32, s->numSlots, b->data, b->numCols,
                                b->numCols, c->data, c->numCols, k, nullptr);
    };
    auto prolog = [&]() {      // column-ELL (the reference's storage) -> sliced ELL, on the device
        s = a->toSliced();
        ks.algBytes = 8.0 * s->numSlots + 4.0 * (s->numSlices + 1.0) + 4.0 * ks.K * N + 4.0 * ks.M * N;
        DenseMatrix<DT, MT> scratch(a->numRows, b->numCols, true);
        return launch(&scratch) == CUSPMM_OK;
    };
    auto *c = runWrapper<DT, MT>(ks, b, ref, prolog, launch);
    delete s;
    return c;
}
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmELLWrapper1(SparseMatrixELL<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref) {
    return ellWrapper<DT, MT, AccT>(1, "sell32_row_kernels", a, b, ref);
}
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmELLWrapper2(SparseMatrixELL<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref) {
    return ellWrapper<DT, MT, AccT>(2, "sell32_staged_tma", a, b, ref);
}
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmELLWrapper3(SparseMatrixELL<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref) {
    return ellWrapper<DT, MT, AccT>(3, "sell32_slice_per_cta", a, b, ref);
}
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmELLWrapper4(SparseMatrixELL<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref) {
    return ellWrapper<DT, MT, AccT>(4, "sell32_staged_tma_tmem", a, b, ref);
}
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmELLWrapper5(SparseMatrixELL<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref) {
    return ellWrapper<DT, MT, AccT>(5, "sell32_all_tmem_quad", a, b, ref);
}
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmELLWrapper6(SparseMatrixELL<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref) {
    return ellWrapper<DT, MT, AccT>(6, "sell32_tensor_split", a, b, ref);
}

// ------------------------------------------------------------------------------- BSR wrappers
template <typename DT, typename MT>
static KernelSpec bsrSpec(int k, const char *name, SparseMatrixBSR<DT, MT> *a, DenseMatrix<DT, MT> *b, double elemBytes) {
    KernelSpec ks;
    ks.format = "BSR"; ks.name = name; ks.kernelNum = k;
    ks.M = a->numRows; ks.K = a->numCols; ks.nnz = a->numNonZero; ks.outRows = a->numBlockRows * a->blockRowSize;
    const double N = b->numCols;
    ks.algBytes = (double)a->numElements * elemBytes + 4.0 * a->numBlocks + 4.0 * (a->numBlockRows + 1.0) +
                  elemBytes * ks.K * N + 4.0 * ks.M * N;
    ks.flops = 2.0 * a->numElements * N;      // executed flops (zeros inside stored blocks included)
    return ks;
}

template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmBSRWrapper1(SparseMatrixBSR<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref) {
    assert(a->onDevice && b->onDevice);
    KernelSpec ks = bsrSpec(1, "bsr_f32_simt", a, b, 4.0);
    auto launch = [&](DenseMatrix<DT, MT> *c) {
        return cuspmm_spmm_bsr_f32(a->blockRowPtrs, a->blockColIdxs, a->data, a->numBlockRows, a->blockRowSize, a->blockColSize,
                                   a->numCols, b->data, b->numCols, b->numCols, c->data, c->numCols, nullptr);
    };
    return runWrapper<DT, MT>(ks, b, ref, nullptr, launch);
}

template <typename DT, typename MT, typename AccT>
static DenseMatrix<DT, MT> *bsrTcWrapper(int k, const char *name, cuspmmBlockType type, SparseMatrixBSR<DT, MT> *a,
                                         DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref) {
    assert(a->onDevice && b->onDevice);
    KernelSpec ks = bsrSpec(k, name, a, b, 2.0);
    cuspmmBsrTcPlan plan = nullptr;
    auto prolog = [&]() {
        if (a->blockRowSize != a->blockColSize || (a->blockRowSize != 16 && a->blockRowSize != 32)) return false;
        if (b->numRows < a->numCols) return false;   // B must cover the (padded) K of the BSR matrix
        cuspmmCheck(cuspmm_bsr_tc_plan_create(&plan, a->blockRowPtrs, a->blockColIdxs, a->data, a->numBlockRows, a->numBlocks,
                                              a->blockRowSize, a->numCols, b->numCols, type, nullptr));
        cuspmmCheck(cuspmm_bsr_tc_prepare_B(plan, b->data, b->numCols, b->numCols, nullptr));
        return true;
    };
    auto launch = [&](DenseMatrix<DT, MT> *c) { return cuspmm_bsr_tc_run(plan, c->data, c->numCols, nullptr); };
    auto *c = runWrapper<DT, MT>(ks, b, ref, prolog, launch);
    if (plan) cuspmm_bsr_tc_plan_destroy(plan);
    return c;
}
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmBSRWrapper2(SparseMatrixBSR<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref) {
    return bsrTcWrapper<DT, MT, AccT>(2, "bsr_tcgen05_bf16", CUSPMM_BLK_BF16, a, b, ref);
}
template <typename DT, typename MT, typename AccT>
DenseMatrix<DT, MT> *spmmBSRWrapper3(SparseMatrixBSR<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *ref) {
    return bsrTcWrapper<DT, MT, AccT>(3, "bsr_tcgen05_fp16", CUSPMM_BLK_FP16, a, b, ref);
}

// ------------------------------------------------------------------------------- instantiations
// (the reference instantiates <float, uint32_t, double> only: e.g. src/spmm/csr/spmm_csr.cpp:32)
using F = float; using U = uint32_t; using A = double;
using Dn = DenseMatrix<F, U>;
template Dn *spmmCSRCpu<F, U, A>(SparseMatrixCSR<F, U> *, Dn *, Dn *);
template Dn *spmmCOOCpu<F, U, A>(SparseMatrixCOO<F, U> *, Dn *, Dn *);
template Dn *spmmELLCpu<F, U, A>(SparseMatrixELL<F, U> *, Dn *, Dn *);
template Dn *spmmBSRCpu<F, U, A>(SparseMatrixBSR<F, U> *, Dn *, Dn *);
template Dn *spmmCSRWrapper1<F, U, A>(SparseMatrixCSR<F, U> *, Dn *, Dn *);
template Dn *spmmCSRWrapper2<F, U, A>(SparseMatrixCSR<F, U> *, Dn *, Dn *);
template Dn *spmmCSRWrapper3<F, U, A>(SparseMatrixCSR<F, U> *, Dn *, Dn *);
template Dn *spmmCSRWrapper4<F, U, A>(SparseMatrixCSR<F, U> *, Dn *, Dn *);
template Dn *spmmCSRWrapper5<F, U, A>(SparseMatrixCSR<F, U> *, Dn *, Dn *);
template Dn *spmmCSRWrapper6<F, U, A>(SparseMatrixCSR<F, U> *, Dn *, Dn *);
template Dn *spmmCSRWrapper7<F, U, A>(SparseMatrixCSR<F, U> *, Dn *, Dn *);
template Dn *spmmCSRWrapper8<F, U, A>(SparseMatrixCSR<F, U> *, Dn *, Dn *);
template Dn *spmmCOOWrapper1<F, U, A>(SparseMatrixCOO<F, U> *, Dn *, Dn *);
template Dn *spmmCOOWrapper2<F, U, A>(SparseMatrixCOO<F, U> *, Dn *, Dn *);
template Dn *spmmELLWrapper1<F, U, A>(SparseMatrixELL<F, U> *, Dn *, Dn *);
template Dn *spmmELLWrapper2<F, U, A>(SparseMatrixELL<F, U> *, Dn *, Dn *);
template Dn *spmmELLWrapper3<F, U, A>(SparseMatrixELL<F, U> *, Dn *, Dn *);
template Dn *spmmELLWrapper4<F, U, A>(SparseMatrixELL<F, U> *, Dn *, Dn *);
template Dn *spmmELLWrapper5<F, U, A>(SparseMatrixELL<F, U> *, Dn *, Dn *);
template Dn *spmmELLWrapper6<F, U, A>(SparseMatrixELL<F, U> *, Dn *, Dn *);
template Dn *spmmBSRWrapper1<F, U, A>(SparseMatrixBSR<F, U> *, Dn *, Dn *);
template Dn *spmmBSRWrapper2<F, U, A>(SparseMatrixBSR<F, U> *, Dn *, Dn *);
template Dn *spmmBSRWrapper3<F, U, A>(SparseMatrixBSR<F, U> *, Dn *, Dn *);

}  // namespace cuspmm
