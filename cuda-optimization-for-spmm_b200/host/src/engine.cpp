// engine.cpp -- runEngine + the multi-GPU row-panel run + cusparseTest.
// Orchestration mirrors src/engine/engine.cpp:17-61 of the reference (host C, H2D of A/B/C, CPU
// kernel 0, GPU kernels 1..numKernels checked against it, cuSPARSE when supported) with the leaks
// fixed (the reference frees only c and dc, :59-60) and two additions from the north star: device-side
// timing inside the wrappers and, with --gpus N, the balanced row-panel multi-GPU run of every format.
#include "engine.hpp"

#include <cmath>
#include <fstream>
#include <functional>
#include <memory>
#include <sstream>
#include <vector>

namespace cuspmm {

RunOptions g_opts;

namespace {
double g_hbmPeak = 0;
std::string g_hbmPeakSource;
void resolveHbmPeak() {
    if (g_hbmPeak > 0) return;
    if (g_opts.hbmPeakGBs > 0) { g_hbmPeak = g_opts.hbmPeakGBs; g_hbmPeakSource = "--hbm-peak"; return; }
    if (const char *e = getenv("CUSPMM_HBM_PEAK_GBS")) {
        const double v = atof(e);
        if (v > 0) { g_hbmPeak = v; g_hbmPeakSource = "CUSPMM_HBM_PEAK_GBS"; return; }
    }
    std::vector<std::string> paths;
    if (const char *e = getenv("CUSPMM_MEASURED_PEAKS")) paths.push_back(e);
    for (const char *pre : {"", "../", "../../", "../../../"}) paths.push_back(std::string(pre) + "MEASURED_PEAKS.json");
    for (const auto &p : paths) {
        std::ifstream in(p);
        if (!in) continue;
        std::stringstream ss;
        ss << in.rdbuf();
        const std::string text = ss.str();
        const size_t k = text.find("\"hbm_gbs\"");
        if (k == std::string::npos) continue;
        const size_t c = text.find(':', k);
        if (c == std::string::npos) continue;
        const double v = atof(text.c_str() + c + 1);
        if (v > 0) { g_hbmPeak = v; g_hbmPeakSource = "measured (" + p + " hbm_gbs)"; return; }
    }
    g_hbmPeak = 6650.0;
    g_hbmPeakSource = "fallback (B200_PROFILING.md)";
}
}  // namespace
double hbmPeakGBs() { resolveHbmPeak(); return g_hbmPeak; }
const char *hbmPeakSource() { resolveHbmPeak(); return g_hbmPeakSource.c_str(); }

namespace {
using Clock = std::chrono::high_resolution_clock;

// Every format: A split into balanced row panels over g_opts.nGpus devices (cuspmm_mgpu_*): CSR / COO by non-zeros at row
// boundaries, ELL by slots at slice boundaries, BSR by blocks at block-row boundaries.  The record's cudaKernelTimeMs is
// the device time (max over the GPUs, CUDA events); e2eTotalTimeMs is the wall clock of create + set_B + one run + get_C,
// i.e. host operands in, host C out.
struct MgpuJob {
    std::string name;
    uint32_t outRows = 0;
    double algBytes = 0, flops = 0;
    int variant = 0;
    std::function<int(cuspmmMgpuPlan *, int, const int *)> create;
};

void runMgpuJob(const MgpuJob &job, uint32_t M, uint32_t K, uint32_t nnz, DenseMatrix<float, uint32_t> *b,
                DenseMatrix<float, uint32_t> *ref, const std::string &fmt) {
    const int n = g_opts.nGpus;
    RecordExtra ex;
    ex.nGpus = n;
    ex.kernelName = job.name + (g_opts.gather ? "+peer_gather" : "");
    std::vector<int> devs(n);
    for (int g = 0; g < n; ++g) devs[g] = g_opts.device + g;
    auto t0 = Clock::now();
    cuspmmMgpuPlan plan = nullptr;
    cuspmmCheck(job.create(&plan, n, devs.data()));
    cuspmmCheck(cuspmm_mgpu_set_B(plan, b->data, b->numCols));
    const double pro = std::chrono::duration_cast<std::chrono::microseconds>(Clock::now() - t0).count() / 1000.0;
    float ms = 0.f, first = 0.f;
    auto tr = Clock::now();
    cuspmmCheck(cuspmm_mgpu_run(plan, job.variant, g_opts.gather, 1, &first));
    const double oneRun = std::chrono::duration_cast<std::chrono::microseconds>(Clock::now() - tr).count() / 1000.0;
    if (g_opts.warmup > 1) cuspmmCheck(cuspmm_mgpu_run(plan, job.variant, g_opts.gather, g_opts.warmup - 1, &ms));
    cuspmmCheck(cuspmm_mgpu_run(plan, job.variant, g_opts.gather, g_opts.iters > 0 ? g_opts.iters : 1, &ms));
    auto t1 = Clock::now();
    DenseMatrix<float, uint32_t> res(job.outRows, b->numCols, false);
    cuspmmCheck(cuspmm_mgpu_get_C(plan, res.data));
    const double epi = std::chrono::duration_cast<std::chrono::microseconds>(Clock::now() - t1).count() / 1000.0;
    std::vector<uint32_t> counts(n);
    cuspmmCheck(cuspmm_mgpu_get_counts(plan, counts.data()));
    uint64_t total = 0;
    uint32_t worst = 0;
    for (uint32_t c : counts) { total += c; worst = std::max(worst, c); }
    ex.imbalance = total ? (double)worst * n / (double)total : 1.0;
    cuspmmCheck(cuspmm_mgpu_destroy(plan));
    cudaCheckError(cudaSetDevice(g_opts.device));
    const size_t cmp = std::min(res.numElements(), ref->numElements());
    const bool correct = allClose(res.data, ref->data, cmp, REL_TOL, ABS_TOL);
    double maxAbs = 0;
    for (size_t i = 0; i < cmp; ++i) maxAbs = std::max(maxAbs, std::fabs((double)res.data[i] - (double)ref->data[i]));
    ex.maxAbsErr = maxAbs;
    ex.algBytes = job.algBytes;
    ex.gflops = job.flops / (ms * 1e-3) / 1e9;
    ex.hbmGBs = ex.algBytes / (ms * 1e-3) / 1e9;
    ex.hbmFrac = ex.hbmGBs / (hbmPeakGBs() * n);
    ex.e2eMs = pro + oneRun + epi;
    reportTime(testcase, M, K, nnz, fmt, b->ordering, 100 + n, pro, ms, epi, correct, &ex);
}

template <typename MaT, typename MbT>
void runMultiGpu(MaT *, MaT *, MbT *, MbT *, const std::string &) {}

using Dn = DenseMatrix<float, uint32_t>;

template <>
void runMultiGpu(SparseMatrixCSR<float, uint32_t> *a, SparseMatrixCSR<float, uint32_t> *, Dn *b, Dn *ref, const std::string &fmt) {
    if (g_opts.nGpus <= 1) return;
    const double N = b->numCols;
    MgpuJob job;
    job.name = "mgpu_row_panels_csr";
    job.outRows = a->numRows;
    job.algBytes = 8.0 * a->numNonZero + 4.0 * (a->numRows + 1.0) + 4.0 * a->numCols * N * g_opts.nGpus + 4.0 * a->numRows * N;
    job.flops = 2.0 * a->numNonZero * N;
    job.create = [&](cuspmmMgpuPlan *pl, int n, const int *devs) {
        return cuspmm_mgpu_create_csr(pl, n, devs, a->rowPtrs, a->colIdxs, a->data, a->numRows, a->numCols, a->numNonZero, b->numCols);
    };
    runMgpuJob(job, a->numRows, a->numCols, a->numNonZero, b, ref, fmt);
}

template <>
void runMultiGpu(SparseMatrixCOO<float, uint32_t> *a, SparseMatrixCOO<float, uint32_t> *, Dn *b, Dn *ref, const std::string &fmt) {
    if (g_opts.nGpus <= 1) return;
    const double N = b->numCols;
    MgpuJob job;
    job.name = "mgpu_row_panels_coo";
    job.outRows = a->numRows;
    job.algBytes = 12.0 * a->numNonZero + 4.0 * a->numCols * N * g_opts.nGpus + 4.0 * a->numRows * N;
    job.flops = 2.0 * a->numNonZero * N;
    job.create = [&](cuspmmMgpuPlan *pl, int n, const int *devs) {
        return cuspmm_mgpu_create_coo(pl, n, devs, a->rowIdxs, a->colIdxs, a->data, a->numRows, a->numCols, a->numNonZero, b->numCols);
    };
    runMgpuJob(job, a->numRows, a->numCols, a->numNonZero, b, ref, fmt);
}

// ELL: the host holds the reference's column-ELL; the engine's sliced layout is produced on the device (da) and read back
// once, then split at slice boundaries
template <>
void runMultiGpu(SparseMatrixELL<float, uint32_t> *a, SparseMatrixELL<float, uint32_t> *da, Dn *b, Dn *ref, const std::string &fmt) {
    if (g_opts.nGpus <= 1) return;
    const double N = b->numCols;
    std::unique_ptr<SlicedELL<float, uint32_t>> s(da->toSliced());
    std::vector<uint32_t> ptrs(s->numSlices + 1), cols(std::max<uint32_t>(s->numSlots, 1));
    std::vector<float> vals(std::max<uint32_t>(s->numSlots, 1));
    cudaCheckError(cudaMemcpy(ptrs.data(), s->slicePtrs, ptrs.size() * 4, cudaMemcpyDeviceToHost));
    if (s->numSlots) {
        cudaCheckError(cudaMemcpy(cols.data(), s->colIdxs, (size_t)s->numSlots * 4, cudaMemcpyDeviceToHost));
        cudaCheckError(cudaMemcpy(vals.data(), s->data, (size_t)s->numSlots * 4, cudaMemcpyDeviceToHost));
    }
    MgpuJob job;
    job.name = "mgpu_slice_panels_sell32";
    job.outRows = a->numRows;
    job.algBytes = 8.0 * s->numSlots + 4.0 * (s->numSlices + 1.0) + 4.0 * a->numCols * N * g_opts.nGpus + 4.0 * a->numRows * N;
    job.flops = 2.0 * a->numNonZero * N;
    const uint32_t slots = s->numSlots;
    job.create = [&](cuspmmMgpuPlan *pl, int n, const int *devs) {
        return cuspmm_mgpu_create_sell(pl, n, devs, ptrs.data(), cols.data(), vals.data(), a->numRows, a->numCols, 32, slots, b->numCols);
    };
    runMgpuJob(job, a->numRows, a->numCols, a->numNonZero, b, ref, fmt);
}

// BSR: block rows balanced by blocks; fp32 kernels (bit-for-bit spmmBSRCpu's order), and for 16x16 / 32x32 blocks also the
// tcgen05 bf16 plan per panel
template <>
void runMultiGpu(SparseMatrixBSR<float, uint32_t> *a, SparseMatrixBSR<float, uint32_t> *, Dn *b, Dn *ref, const std::string &fmt) {
    if (g_opts.nGpus <= 1) return;
    if (b->numRows < a->numCols) return;       // B must cover the (padded) K of the BSR matrix
    const double N = b->numCols;
    const bool tcOk = a->blockRowSize == a->blockColSize && (a->blockRowSize == 16 || a->blockRowSize == 32);
    for (int variant : {1, 2}) {
        if (variant == 2 && !tcOk) continue;
        const double e = variant == 1 ? 4.0 : 2.0;
        MgpuJob job;
        job.name = variant == 1 ? "mgpu_blockrow_panels_bsr_f32" : "mgpu_blockrow_panels_bsr_tcgen05_bf16";
        job.variant = variant;
        job.outRows = a->numBlockRows * a->blockRowSize;
        job.algBytes = (double)a->numElements * e + 4.0 * a->numBlocks + 4.0 * (a->numBlockRows + 1.0) + e * a->numCols * N * g_opts.nGpus +
                       4.0 * job.outRows * N;
        job.flops = 2.0 * a->numElements * N;
        job.create = [&](cuspmmMgpuPlan *pl, int n, const int *devs) {
            return cuspmm_mgpu_create_bsr(pl, n, devs, a->blockRowPtrs, a->blockColIdxs, a->data, a->numBlockRows, a->blockRowSize,
                                          a->blockColSize, a->numCols, b->numCols);
        };
        if (variant == 2) {      // the tensor-core result is compared with the reference's tolerances all the same
            runMgpuJob(job, a->numRows, a->numCols, a->numNonZero, b, ref, fmt);
        } else {
            runMgpuJob(job, a->numRows, a->numCols, a->numNonZero, b, ref, fmt);
        }
    }
}
}  // namespace

template <typename EngT>
void runEngine(EngT *engine, typename EngT::MataT *a, typename EngT::MatbT *b, float abs_tol, float rel_tol, bool skipSeq) {
    (void)abs_tol; (void)rel_tol;   // unused by the reference as well (the wrappers use REL_TOL / ABS_TOL)
    using ma_t = typename EngT::MataT;
    using mb_t = typename EngT::MatbT;
    mb_t *c = new mb_t(a->numRows, b->numCols, false, ORDERING::ROW_MAJOR);

    // 1. move to device
    ma_t *da = a->copy2Device();
    mb_t *db = b->copy2Device();

    // 2. CPU kernel 0 (the checker); its time goes into the cudaKernelTimeMs slot of record 0 (engine.cpp:36-37)
    auto seqStart = Clock::now();
    mb_t *cpuRes = c;
    if (!skipSeq) cpuRes = reinterpret_cast<mb_t *>(engine->runKernel(0, a, b, c));
    const double seqMs = std::chrono::duration_cast<std::chrono::microseconds>(Clock::now() - seqStart).count() / 1000.0;
    engine->logSeq(seqMs);
    reportTime(testcase, a->numRows, a->numCols, a->numNonZero, engine->fmt, b->ordering, 0, 0, seqMs, 0, 1);

    // 3. GPU kernels, each checked against the CPU result
    for (int i = 1; i <= engine->numKernels; ++i) {
        if (g_opts.onlyKernel && g_opts.onlyKernel != i) continue;
        auto *kRes = reinterpret_cast<mb_t *>(engine->runKernel(i, da, db, cpuRes));
        delete kRes;
    }

    // 4. cuSPARSE, same run, checked too (the reference passes correct = 1 unchecked, engine.cpp:54-55)
    if (engine->SUPPORT_CUSPARSE && !g_opts.onlyKernel) {
        mb_t *dc = new mb_t(a->numRows, b->numCols, true, ORDERING::ROW_MAJOR);
        long pro = 0, kernel = 0, epi = 0;
        cusparseTest<typename ma_t::DT, typename ma_t::MT>(da, db, dc, pro, kernel, epi);
        auto t1 = Clock::now();
        mb_t *tmp = dc->copy2Host();
        epi += std::chrono::duration_cast<std::chrono::microseconds>(Clock::now() - t1).count();
        const bool correct = allClose(tmp->data, cpuRes->data, tmp->numElements(), REL_TOL, ABS_TOL);
        RecordExtra ex;
        ex.kernelName = "cusparseSpMM";
        ex.gflops = 2.0 * a->numNonZero * (double)b->numCols / (kernel * 1e-6) / 1e9;
        reportTime(testcase, a->numRows, a->numCols, a->numNonZero, engine->fmt, db->ordering, -1, pro / 1000.0, kernel / 1000.0,
                   epi / 1000.0, correct, &ex);
        delete tmp;
        delete dc;
    }

    // 5. multi-GPU row panels (every format)
    runMultiGpu(a, da, b, cpuRes, engine->fmt);

    delete da;
    delete db;
    delete c;
}

// cusparseTest: the reference's call sequence (src/engine/cusparse.cu:19-54) with the kernel timed by CUDA
// events over g_opts.iters launches and cusparseSpMM_preprocess in the prolog.
template <typename DT, typename MT>
DenseMatrix<DT, MT> *cusparseTest(SparseMatrix<DT, MT> *a, DenseMatrix<DT, MT> *b, DenseMatrix<DT, MT> *c, long &pro, long &kernel,
                                  long &epi) {
    cusparseHandle_t handle;
    cusparseSpMatDescr_t matA;
    cusparseDnMatDescr_t matB, matC;
    auto t1 = Clock::now();
    CHECK_CUSPARSE(cusparseCreate(&handle));
    a->setCusparseSpMatDesc(&matA);
    b->setCusparseDnMatDesc(&matB);
    c->setCusparseDnMatDesc(&matC);
    const float alpha = 1.0f, beta = 0.f;
    void *dBuffer = nullptr;
    size_t bufferSize = 0;
    const cusparseSpMMAlg_t alg = a->getCusparseAlg();
    const cudaDataType ct = std::is_same_v<DT, float> ? CUDA_R_32F : CUDA_R_64F;
    CHECK_CUSPARSE(cusparseSpMM_bufferSize(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, matA,
                                           matB, &beta, matC, ct, alg, &bufferSize));
    cudaCheckError(cudaMalloc(&dBuffer, bufferSize ? bufferSize : 1));
    CHECK_CUSPARSE(cusparseSpMM_preprocess(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, matA,
                                           matB, &beta, matC, ct, alg, dBuffer));
    for (int i = 0; i < g_opts.warmup; ++i)
        CHECK_CUSPARSE(cusparseSpMM(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, matA, matB,
                                    &beta, matC, ct, alg, dBuffer));
    cudaCheckError(cudaDeviceSynchronize());
    auto t2 = Clock::now();
    cudaEvent_t e0, e1;
    cudaCheckError(cudaEventCreate(&e0));
    cudaCheckError(cudaEventCreate(&e1));
    const int iters = g_opts.iters > 0 ? g_opts.iters : 1;
    cudaCheckError(cudaEventRecord(e0, 0));
    for (int i = 0; i < iters; ++i)
        CHECK_CUSPARSE(cusparseSpMM(handle, CUSPARSE_OPERATION_NON_TRANSPOSE, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, matA, matB,
                                    &beta, matC, ct, alg, dBuffer));
    cudaCheckError(cudaEventRecord(e1, 0));
    cudaCheckError(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaCheckError(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    auto t3 = Clock::now();
    CHECK_CUSPARSE(cusparseDestroySpMat(matA));
    CHECK_CUSPARSE(cusparseDestroyDnMat(matB));
    CHECK_CUSPARSE(cusparseDestroyDnMat(matC));
    CHECK_CUSPARSE(cusparseDestroy(handle));
    cudaCheckError(cudaFree(dBuffer));
    auto t4 = Clock::now();
    pro = std::chrono::duration_cast<std::chrono::microseconds>(t2 - t1).count();
    kernel = (long)(ms / iters * 1000.0);
    epi = std::chrono::duration_cast<std::chrono::microseconds>(t4 - t3).count();
    return c;
}

template DenseMatrix<float, uint32_t> *cusparseTest(SparseMatrix<float, uint32_t> *, DenseMatrix<float, uint32_t> *,
                                                    DenseMatrix<float, uint32_t> *, long &, long &, long &);

#define ENG_INST(fmt)                                                                                                      \
    template void runEngine<Engine##fmt<float, uint32_t, double>>(Engine##fmt<float, uint32_t, double> *,                  \
                                                                   Engine##fmt<float, uint32_t, double>::MataT *,           \
                                                                   Engine##fmt<float, uint32_t, double>::MatbT *, float, float, bool);
ENG_INST(BSR)
ENG_INST(COO)
ENG_INST(CSR)
ENG_INST(ELL)

}  // namespace cuspmm
