// host_api.cu -- host-buffer entry points (all four formats) and the multi-GPU row-panel engine (all four formats).
//
//  * cuspmm_spmm_{csr,coo,sell,bsr}_host: what runEngine + spmm<FMT>Wrapper<k> do end to end in the reference
//    (src/engine/engine.cpp:20-44: H2D of A, B, C; kernel; D2H of C), but pipelined: A is cut into balanced row panels
//    (rows / slices / block rows) and panel p+1's H2D overlaps panel p's kernel and panel p-1's D2H on three streams.
//    Device buffers are cached per device (grow-only, released by cuspmm_host_pipeline_release), so the "prolog"
//    cudaMalloc + memset of the reference (spmm_csr_k3.cu:69-73) is gone.
//  * cuspmm_mgpu_*: north_star (c) -- A split into balanced row panels over the GPUs of one node (CSR: split points from
//    the DEVICE partitioner; COO: the same at row boundaries; sliced ELL: slice boundaries balanced by slots; BSR: block
//    rows balanced by blocks), B replicated over NVLink with peer copies, each GPU multiplies its panel on its own stream;
//    with gather the kernels store their C rows directly into GPU 0's C through peer memory (compute and "collective"
//    are one kernel: the stores travel over NVLink while the panel is still being computed).
#include "common.cuh"

#include <algorithm>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

namespace cuspmm_b200 {

int spmm_csr_dispatch(const uint32_t *, const uint32_t *, const float *, uint32_t, uint32_t, uint32_t,
                      const float *, uint32_t, size_t, float *, size_t, int, cudaStream_t);
int bsr_tc_plan_create_impl(cuspmmBsrTcPlan *out, const uint32_t *blockRowPtrs, const uint32_t *blockColIdxs, const float *blocks,
                            uint32_t numBlockRows, uint32_t numBlocks, uint32_t bs, uint32_t K, uint32_t maxN, cuspmmBlockType type,
                            void *stream, void *blocksQ_ext, void *Bq_ext);      // spmm_bsr_tc.cu
size_t bsr_tc_blocks_bytes(uint32_t numBlocks, uint32_t bs);
size_t bsr_tc_B_bytes(uint32_t K, uint32_t bs, uint32_t maxN);
int coo_panel_rowptrs(const uint32_t *rowIdxs, uint32_t rowBase, uint32_t rows, uint32_t cnt, uint32_t idxBase,
                      uint32_t *rowPtrs, cudaStream_t st);      // convert.cu

__global__ void rebase_kernel(uint32_t *p, uint32_t n, uint32_t base) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] -= base;
}

struct GrowBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

constexpr int kMaxPanels = 16;

// One pipeline (three streams, events, cached device buffers) per device ordinal; used by one call at a time.
struct HostPipe {
    GrowBuf a0, a1, a2, a3, B, C;          // format arrays (pointers / indices / values / workspace), B, C
    GrowBuf tcBlocks, tcB;                 // 16-bit re-tiled operands of the tensor-core BSR plan (lent to the plan of the call)
    cudaStream_t up = nullptr, run = nullptr, down = nullptr;
    cudaEvent_t upDone[kMaxPanels] = {}, runDone[kMaxPanels] = {};
    cudaEvent_t t0 = nullptr, t1 = nullptr, bDone = nullptr, bReady = nullptr;
    cuspmmBsrTcPlan tc = nullptr;          // tensor-core BSR plan of the call in flight (built after the upload, destroyed before return)
    std::mutex mu;
    bool init = false;
    void destroy() {
        if (tc) { cuspmm_bsr_tc_plan_destroy(tc); tc = nullptr; }
        a0.release(); a1.release(); a2.release(); a3.release(); B.release(); C.release(); tcBlocks.release(); tcB.release();
        if (init) {
            cudaStreamDestroy(up); cudaStreamDestroy(run); cudaStreamDestroy(down);
            for (int i = 0; i < kMaxPanels; ++i) { cudaEventDestroy(upDone[i]); cudaEventDestroy(runDone[i]); }
            cudaEventDestroy(t0); cudaEventDestroy(t1); cudaEventDestroy(bDone); cudaEventDestroy(bReady);
            init = false;
        }
    }
};
static std::mutex g_pipes_mu;
static std::map<int, std::unique_ptr<HostPipe>> g_pipes;       // any device ordinal

static HostPipe *pipe_for(int dev) {
    std::lock_guard<std::mutex> lock(g_pipes_mu);
    auto &slot = g_pipes[dev];
    if (!slot) slot.reset(new HostPipe());
    return slot.get();
}

static int pipe_init(HostPipe &hp) {
    if (hp.init) return CUSPMM_OK;
    CUSPMM_CUDA(cudaStreamCreateWithFlags(&hp.up, cudaStreamNonBlocking));
    CUSPMM_CUDA(cudaStreamCreateWithFlags(&hp.run, cudaStreamNonBlocking));
    CUSPMM_CUDA(cudaStreamCreateWithFlags(&hp.down, cudaStreamNonBlocking));
    for (int i = 0; i < kMaxPanels; ++i) {
        CUSPMM_CUDA(cudaEventCreateWithFlags(&hp.upDone[i], cudaEventDisableTiming));
        CUSPMM_CUDA(cudaEventCreateWithFlags(&hp.runDone[i], cudaEventDisableTiming));
    }
    CUSPMM_CUDA(cudaEventCreate(&hp.t0));
    CUSPMM_CUDA(cudaEventCreate(&hp.t1));
    CUSPMM_CUDA(cudaEventCreateWithFlags(&hp.bDone, cudaEventDisableTiming));
    CUSPMM_CUDA(cudaEventCreateWithFlags(&hp.bReady, cudaEventDisableTiming));
    hp.init = true;
    return CUSPMM_OK;
}

// Whatever way a host-buffer call ends, nothing may still be reading the caller's host buffers or the cached device
// buffers when the pipeline mutex is released: synchronise the three streams on every exit path.
struct PipeDrain {
    HostPipe &hp;
    ~PipeDrain() {
        if (!hp.init) return;
        cudaStreamSynchronize(hp.up);
        cudaStreamSynchronize(hp.run);
        cudaStreamSynchronize(hp.down);
    }
};

// Number of panels of the host pipeline: the staged kernels stream all of B once per ~60-row CTA, so a panel must still
// give every SM a full CTA (>= ~6400 rows: 16 thin panels made the kernels 3x slower than the copies); within that, more
// panels = finer overlap of H2D(p+1) / kernel(p) / D2H(p-1).  Tapering makes the early panels bigger, so one more is free.
uint32_t host_panel_count(uint64_t aBytes, uint32_t rows) {
    uint32_t parts = (uint32_t)std::min<uint64_t>(kMaxPanels, std::max<uint64_t>(1, aBytes / (32u << 20)));
    parts = std::min(parts, std::max(1u, (rows + 3200u) / 6400u));
    if (parts >= 3) ++parts;
    return std::min<uint32_t>(parts, kMaxPanels);
}

// Split points of a prefix array ptr[0..n] (row pointers, slice pointers, block-row pointers) into `parts` contiguous
// ranges balanced by the prefix value.  taper = false: equal shares (the multi-GPU rule, the same as the device
// partitioner).  taper = true: shares shrink linearly towards the end (the host pipeline: the copies are the bottleneck,
// so what matters is how much kernel + D2H is left once the last H2D lands -- a small last panel keeps that tail short).
std::vector<uint32_t> prefix_splits(const uint32_t *ptr, uint32_t n, uint32_t parts, bool taper) {
    std::vector<uint32_t> s(parts + 1, 0);
    const uint64_t total = ptr[n] - ptr[0];
    const uint64_t wsum = (uint64_t)parts * (parts + 1) / 2;      // weights parts, parts-1, ..., 1
    uint64_t wacc = 0;
    for (uint32_t g = 1; g < parts; ++g) {
        wacc += parts - (g - 1);
        const uint64_t t = ptr[0] + (taper ? (total * wacc) / wsum : ((uint64_t)g * total) / parts);
        s[g] = (uint32_t)(std::lower_bound(ptr, ptr + n + 1, (uint32_t)t) - ptr);
        if (s[g] > n) s[g] = n;
        if (s[g] < s[g - 1]) s[g] = s[g - 1];
    }
    s[parts] = n;
    return s;
}

// COO: entry-index split points at ROW boundaries (entries sorted by row).  idx[g] = first entry of the first row that
// starts at or after the g-th share; rows[g] = that row.
static void coo_splits(const uint32_t *rowIdxs, uint32_t M, uint32_t nnz, uint32_t parts, bool taper,
                       std::vector<uint32_t> &rows, std::vector<uint32_t> &idx) {
    rows.assign(parts + 1, 0);
    idx.assign(parts + 1, 0);
    const uint64_t wsum = (uint64_t)parts * (parts + 1) / 2;
    uint64_t wacc = 0;
    for (uint32_t g = 1; g < parts; ++g) {
        wacc += parts - (g - 1);
        const uint64_t t = taper ? ((uint64_t)nnz * wacc) / wsum : ((uint64_t)g * nnz) / parts;
        uint32_t r = t < nnz ? rowIdxs[t] : M;
        if (t > 0 && t < nnz && rowIdxs[t - 1] == r) ++r;          // t falls inside row r: the panel starts at the next row
        if (r > M) r = M;
        if (r < rows[g - 1]) r = rows[g - 1];
        rows[g] = r;
        idx[g] = (uint32_t)(std::lower_bound(rowIdxs, rowIdxs + nnz, r) - rowIdxs);
    }
    rows[parts] = M;
    idx[parts] = nnz;
}

// The panel loop shared by the four formats.  upload(p) enqueues the H2D of panel p's share of A on hp.up; launch(p)
// enqueues its kernel(s) on hp.run writing rows [rows[p], rows[p+1]) of dC.
template <class Upload, class Launch>
static int run_panels(HostPipe &hp, uint32_t parts, const std::vector<uint32_t> &rows, uint32_t N, float *C, float *dC,
                      Upload upload, Launch launch) {
    for (uint32_t p = 0; p < parts; ++p) {
        const uint32_t r0 = rows[p], r1 = rows[p + 1];
        int rc = upload(p);
        if (rc) return rc;
        CUSPMM_CUDA(cudaEventRecord(hp.upDone[p], hp.up));
        if (r1 == r0) continue;
        CUSPMM_CUDA(cudaStreamWaitEvent(hp.run, hp.upDone[p], 0));
        rc = launch(p);
        if (rc) return rc;
        CUSPMM_CUDA(cudaEventRecord(hp.runDone[p], hp.run));
        CUSPMM_CUDA(cudaStreamWaitEvent(hp.down, hp.runDone[p], 0));
        CUSPMM_CUDA(cudaMemcpyAsync(C + (size_t)r0 * N, dC + (size_t)r0 * N, (size_t)(r1 - r0) * N * 4,
                                    cudaMemcpyDeviceToHost, hp.down));
    }
    return CUSPMM_OK;
}

// common prologue / epilogue of the host-buffer calls
struct HostCall {
    HostPipe *hp = nullptr;
    std::unique_lock<std::mutex> lock;
};
static int host_begin(HostCall &hc) {
    int dev = 0;
    CUSPMM_CUDA(cudaGetDevice(&dev));
    hc.hp = pipe_for(dev);
    hc.lock = std::unique_lock<std::mutex>(hc.hp->mu);
    return pipe_init(*hc.hp);
}
// B: host pointer (uploaded on hp.up, first in the queue) or device pointer (the run stream waits for `b_ready_stream`)
static int stage_B(HostPipe &hp, const float *B, int b_on_device, void *b_ready_stream, uint32_t K, uint32_t N, size_t ldb,
                   const float **dB, size_t *dldb) {
    if (b_on_device) {
        CUSPMM_CUDA(cudaEventRecord(hp.bReady, as_stream(b_ready_stream)));
        CUSPMM_CUDA(cudaStreamWaitEvent(hp.run, hp.bReady, 0));
        *dB = B;
        *dldb = ldb;
        return CUSPMM_OK;
    }
    CUSPMM_CUDA(hp.B.reserve((size_t)K * N * 4));
    if (ldb == N) CUSPMM_CUDA(cudaMemcpyAsync(hp.B.p, B, (size_t)K * N * 4, cudaMemcpyHostToDevice, hp.up));
    else CUSPMM_CUDA(cudaMemcpy2DAsync(hp.B.p, (size_t)N * 4, B, ldb * 4, (size_t)N * 4, K, cudaMemcpyHostToDevice, hp.up));
    *dB = static_cast<const float *>(hp.B.p);
    *dldb = N;
    return CUSPMM_OK;
}
static int host_end(HostPipe &hp, float *device_ms) {
    CUSPMM_CUDA(cudaEventRecord(hp.bDone, hp.up));
    CUSPMM_CUDA(cudaStreamWaitEvent(hp.down, hp.bDone, 0));
    CUSPMM_CUDA(cudaEventRecord(hp.t1, hp.down));
    CUSPMM_CUDA(cudaStreamSynchronize(hp.down));
    CUSPMM_CUDA(cudaStreamSynchronize(hp.run));
    if (device_ms) CUSPMM_CUDA(cudaEventElapsedTime(device_ms, hp.t0, hp.t1));
    return CUSPMM_OK;
}

static int csr_host_impl(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals, uint32_t M, uint32_t K,
                         uint32_t nnz, const float *B, int b_on_device, size_t ldb, void *b_ready_stream, uint32_t N,
                         float *C, int variant, float *device_ms) {
    CUSPMM_REQUIRE(rowPtrs && B && C && (nnz == 0 || (colIdxs && vals)), "null operand pointer");
    CUSPMM_REQUIRE(ldb >= N, "ldb (%zu) must be >= N (%u)", ldb, N);
    if (M == 0 || N == 0) return CUSPMM_OK;
    HostCall hc;
    int rc = host_begin(hc);
    if (rc) return rc;
    HostPipe &hp = *hc.hp;
    PipeDrain drain{hp};
    CUSPMM_CUDA(hp.a0.reserve((size_t)(M + 1) * 4));
    CUSPMM_CUDA(hp.a1.reserve((size_t)std::max(nnz, 1u) * 4));
    CUSPMM_CUDA(hp.a2.reserve((size_t)std::max(nnz, 1u) * 4));
    CUSPMM_CUDA(hp.C.reserve((size_t)M * N * 4));
    uint32_t *dRow = static_cast<uint32_t *>(hp.a0.p), *dCol = static_cast<uint32_t *>(hp.a1.p);
    float *dVal = static_cast<float *>(hp.a2.p), *dC = static_cast<float *>(hp.C.p);

    const uint32_t parts = host_panel_count((uint64_t)nnz * 8, M);
    const std::vector<uint32_t> sp = prefix_splits(rowPtrs, M, parts, /*taper=*/true);

    CUSPMM_CUDA(cudaEventRecord(hp.t0, hp.up));
    CUSPMM_CUDA(cudaMemcpyAsync(dRow, rowPtrs, (size_t)(M + 1) * 4, cudaMemcpyHostToDevice, hp.up));
    const float *dB = nullptr;
    size_t dldb = 0;
    rc = stage_B(hp, B, b_on_device, b_ready_stream, K, N, ldb, &dB, &dldb);
    if (rc) return rc;
    rc = run_panels(hp, parts, sp, N, C, dC,
        [&](uint32_t p) -> int {
            const uint32_t i0 = rowPtrs[sp[p]], i1 = rowPtrs[sp[p + 1]];
            if (i1 > i0) {
                CUSPMM_CUDA(cudaMemcpyAsync(dCol + i0, colIdxs + i0, (size_t)(i1 - i0) * 4, cudaMemcpyHostToDevice, hp.up));
                CUSPMM_CUDA(cudaMemcpyAsync(dVal + i0, vals + i0, (size_t)(i1 - i0) * 4, cudaMemcpyHostToDevice, hp.up));
            }
            return CUSPMM_OK;
        },
        [&](uint32_t p) -> int {
            const uint32_t r0 = sp[p], r1 = sp[p + 1];
            return spmm_csr_dispatch(dRow + r0, dCol, dVal, r1 - r0, K, rowPtrs[r1] - rowPtrs[r0], dB, N, dldb,
                                     dC + (size_t)r0 * N, N, variant, hp.run);
        });
    if (rc) return rc;
    return host_end(hp, device_ms);
}

} // namespace cuspmm_b200

using namespace cuspmm_b200;

extern "C" int cuspmm_spmm_csr_host(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals,
                                    uint32_t M, uint32_t K, uint32_t nnz, const float *B, uint32_t N, float *C,
                                    int variant, float *device_ms) {
    return csr_host_impl(rowPtrs, colIdxs, vals, M, K, nnz, B, 0, N, nullptr, N, C, variant, device_ms);
}

extern "C" int cuspmm_spmm_csr_host_devB(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals,
                                         uint32_t M, uint32_t K, uint32_t nnz, const float *B_dev, size_t ldb,
                                         void *b_ready_stream, uint32_t N, float *C, int variant, float *device_ms) {
    return csr_host_impl(rowPtrs, colIdxs, vals, M, K, nnz, B_dev, 1, ldb, b_ready_stream, N, C, variant, device_ms);
}

extern "C" int cuspmm_spmm_coo_host(const uint32_t *rowIdxs, const uint32_t *colIdxs, const float *vals,
                                    uint32_t M, uint32_t K, uint32_t nnz, const float *B, uint32_t N, float *C,
                                    int variant, float *device_ms) {
    CUSPMM_REQUIRE(B && C && (nnz == 0 || (rowIdxs && colIdxs && vals)), "null operand pointer");
    CUSPMM_REQUIRE(variant >= 0 && variant <= CUSPMM_COO_NUM_VARIANTS, "COO variant %d does not exist", variant);
    if (M == 0 || N == 0) return CUSPMM_OK;
    HostCall hc;
    int rc = host_begin(hc);
    if (rc) return rc;
    HostPipe &hp = *hc.hp;
    PipeDrain drain{hp};
    CUSPMM_CUDA(hp.a0.reserve((size_t)std::max(nnz, 1u) * 4));          // rowIdxs
    CUSPMM_CUDA(hp.a1.reserve((size_t)std::max(nnz, 1u) * 4));
    CUSPMM_CUDA(hp.a2.reserve((size_t)std::max(nnz, 1u) * 4));
    CUSPMM_CUDA(hp.a3.reserve((size_t)(M + 1) * 4));                    // row pointers built on the device, panel by panel
    CUSPMM_CUDA(hp.C.reserve((size_t)M * N * 4));
    uint32_t *dRowIdx = static_cast<uint32_t *>(hp.a0.p), *dCol = static_cast<uint32_t *>(hp.a1.p);
    uint32_t *dRowPtr = static_cast<uint32_t *>(hp.a3.p);
    float *dVal = static_cast<float *>(hp.a2.p), *dC = static_cast<float *>(hp.C.p);

    // variant 1 (the row-aligned COO kernel) addresses C by absolute row and zero-fills the rows between its entries, so it
    // runs once over the whole matrix; variants 0 / 2 go panel by panel through device-built row pointers
    const uint32_t parts = variant == 1 ? 1u : host_panel_count((uint64_t)nnz * 12, M);
    std::vector<uint32_t> rows, idx;
    coo_splits(rowIdxs, M, nnz, parts, /*taper=*/true, rows, idx);

    CUSPMM_CUDA(cudaEventRecord(hp.t0, hp.up));
    const float *dB = nullptr;
    size_t dldb = 0;
    rc = stage_B(hp, B, 0, nullptr, K, N, N, &dB, &dldb);
    if (rc) return rc;
    rc = run_panels(hp, parts, rows, N, C, dC,
        [&](uint32_t p) -> int {
            const uint32_t i0 = idx[p], i1 = idx[p + 1];
            if (i1 > i0) {
                CUSPMM_CUDA(cudaMemcpyAsync(dRowIdx + i0, rowIdxs + i0, (size_t)(i1 - i0) * 4, cudaMemcpyHostToDevice, hp.up));
                CUSPMM_CUDA(cudaMemcpyAsync(dCol + i0, colIdxs + i0, (size_t)(i1 - i0) * 4, cudaMemcpyHostToDevice, hp.up));
                CUSPMM_CUDA(cudaMemcpyAsync(dVal + i0, vals + i0, (size_t)(i1 - i0) * 4, cudaMemcpyHostToDevice, hp.up));
            }
            return CUSPMM_OK;
        },
        [&](uint32_t p) -> int {
            const uint32_t r0 = rows[p], r1 = rows[p + 1], i0 = idx[p], i1 = idx[p + 1];
            if (variant == 1)
                return cuspmm_spmm_coo(dRowIdx, dCol, dVal, M, K, nnz, dB, N, dldb, dC, N, 1, nullptr, 0, hp.run);
            int rc2 = coo_panel_rowptrs(dRowIdx + i0, r0, r1 - r0, i1 - i0, i0, dRowPtr + r0, hp.run);
            if (rc2) return rc2;
            return spmm_csr_dispatch(dRowPtr + r0, dCol, dVal, r1 - r0, K, i1 - i0, dB, N, dldb, dC + (size_t)r0 * N, N, 0, hp.run);
        });
    if (rc) return rc;
    return host_end(hp, device_ms);
}

extern "C" int cuspmm_spmm_sell_host(const uint32_t *slicePtrs, const uint32_t *colIdxs, const float *vals,
                                     uint32_t M, uint32_t K, uint32_t sliceH, uint32_t numSlots, const float *B,
                                     uint32_t N, float *C, int variant, float *device_ms) {
    CUSPMM_REQUIRE(slicePtrs && B && C && (numSlots == 0 || (colIdxs && vals)), "null operand pointer");
    CUSPMM_REQUIRE(sliceH == 32, "sliced ELL kernels are built for slices of 32 rows (got %u)", sliceH);
    if (M == 0 || N == 0) return CUSPMM_OK;
    HostCall hc;
    int rc = host_begin(hc);
    if (rc) return rc;
    HostPipe &hp = *hc.hp;
    PipeDrain drain{hp};
    const uint32_t slices = (M + 31) / 32;
    CUSPMM_CUDA(hp.a0.reserve((size_t)(slices + 1) * 4));
    CUSPMM_CUDA(hp.a1.reserve((size_t)std::max(numSlots, 1u) * 4));
    CUSPMM_CUDA(hp.a2.reserve((size_t)std::max(numSlots, 1u) * 4));
    CUSPMM_CUDA(hp.C.reserve((size_t)M * N * 4));
    uint32_t *dPtr = static_cast<uint32_t *>(hp.a0.p), *dCol = static_cast<uint32_t *>(hp.a1.p);
    float *dVal = static_cast<float *>(hp.a2.p), *dC = static_cast<float *>(hp.C.p);

    const uint32_t parts = host_panel_count((uint64_t)numSlots * 8, M);
    const std::vector<uint32_t> ss = prefix_splits(slicePtrs, slices, parts, /*taper=*/true);     // slice boundaries
    std::vector<uint32_t> rows(parts + 1);
    for (uint32_t p = 0; p <= parts; ++p) rows[p] = std::min<uint64_t>(M, (uint64_t)ss[p] * 32);

    CUSPMM_CUDA(cudaEventRecord(hp.t0, hp.up));
    CUSPMM_CUDA(cudaMemcpyAsync(dPtr, slicePtrs, (size_t)(slices + 1) * 4, cudaMemcpyHostToDevice, hp.up));
    const float *dB = nullptr;
    size_t dldb = 0;
    rc = stage_B(hp, B, 0, nullptr, K, N, N, &dB, &dldb);
    if (rc) return rc;
    rc = run_panels(hp, parts, rows, N, C, dC,
        [&](uint32_t p) -> int {
            const uint32_t i0 = slicePtrs[ss[p]], i1 = slicePtrs[ss[p + 1]];
            if (i1 > i0) {
                CUSPMM_CUDA(cudaMemcpyAsync(dCol + i0, colIdxs + i0, (size_t)(i1 - i0) * 4, cudaMemcpyHostToDevice, hp.up));
                CUSPMM_CUDA(cudaMemcpyAsync(dVal + i0, vals + i0, (size_t)(i1 - i0) * 4, cudaMemcpyHostToDevice, hp.up));
            }
            return CUSPMM_OK;
        },
        [&](uint32_t p) -> int {
            const uint32_t r0 = rows[p], r1 = rows[p + 1];
            return cuspmm_spmm_sell(dPtr + ss[p], dCol, dVal, r1 - r0, K, 32, slicePtrs[ss[p + 1]] - slicePtrs[ss[p]], dB, N, dldb,
                                    dC + (size_t)r0 * N, N, variant, hp.run);
        });
    if (rc) return rc;
    return host_end(hp, device_ms);
}

extern "C" int cuspmm_spmm_bsr_host(const uint32_t *blockRowPtrs, const uint32_t *blockColIdxs, const float *blocks,
                                    uint32_t numBlockRows, uint32_t br, uint32_t bc, uint32_t K, const float *B, uint32_t N,
                                    float *C, int variant, float *device_ms) {
    CUSPMM_REQUIRE(blockRowPtrs && B && C && br > 0 && bc > 0, "bad arguments");
    CUSPMM_REQUIRE(variant >= 0 && variant <= CUSPMM_BSR_NUM_VARIANTS, "BSR variant %d does not exist", variant);
    const uint64_t M64 = (uint64_t)numBlockRows * br;
    CUSPMM_REQUIRE(M64 <= 0xFFFFFFFFull, "numBlockRows * br overflows uint32");
    const uint32_t M = (uint32_t)M64;
    if (M == 0 || N == 0) return CUSPMM_OK;
    if (variant == 0) variant = 1;                 // fp32, bit-for-bit the reference's order; 2 / 3 = tensor cores (rounded operands)
    const bool tc = variant >= 2;
    if (tc && !(br == bc && (br == 16 || br == 32)))
        return set_error(CUSPMM_ERR_UNSUPPORTED, "tensor-core BSR supports 16x16 and 32x32 blocks (got %ux%u)", br, bc);
    const uint32_t nb = blockRowPtrs[numBlockRows];
    CUSPMM_REQUIRE(nb == 0 || (blockColIdxs && blocks), "null operand pointer");
    HostCall hc;
    int rc = host_begin(hc);
    if (rc) return rc;
    HostPipe &hp = *hc.hp;
    PipeDrain drain{hp};
    const size_t be = (size_t)br * bc;
    CUSPMM_CUDA(hp.a0.reserve((size_t)(numBlockRows + 1) * 4));
    CUSPMM_CUDA(hp.a1.reserve((size_t)std::max(nb, 1u) * 4));
    CUSPMM_CUDA(hp.a2.reserve((size_t)std::max(nb, 1u) * be * 4));
    CUSPMM_CUDA(hp.C.reserve((size_t)M * N * 4));
    if (tc) {          // the plan's 16-bit operands live in cached buffers too: no cudaMalloc / cudaFree per call
        CUSPMM_CUDA(hp.tcBlocks.reserve(bsr_tc_blocks_bytes(nb, br)));
        CUSPMM_CUDA(hp.tcB.reserve(bsr_tc_B_bytes(K, br, N)));
    }
    uint32_t *dPtr = static_cast<uint32_t *>(hp.a0.p), *dCol = static_cast<uint32_t *>(hp.a1.p);
    float *dBlk = static_cast<float *>(hp.a2.p), *dC = static_cast<float *>(hp.C.p);

    // tensor cores: the plan re-tiles all blocks at once, so the multiply is one launch after the last upload
    const uint32_t parts = tc ? 1u : host_panel_count((uint64_t)nb * be * 4, M);
    const std::vector<uint32_t> bs = prefix_splits(blockRowPtrs, numBlockRows, parts, /*taper=*/true);
    std::vector<uint32_t> rows(parts + 1);
    for (uint32_t p = 0; p <= parts; ++p) rows[p] = bs[p] * br;

    CUSPMM_CUDA(cudaEventRecord(hp.t0, hp.up));
    CUSPMM_CUDA(cudaMemcpyAsync(dPtr, blockRowPtrs, (size_t)(numBlockRows + 1) * 4, cudaMemcpyHostToDevice, hp.up));
    const float *dB = nullptr;
    size_t dldb = 0;
    rc = stage_B(hp, B, 0, nullptr, K, N, N, &dB, &dldb);
    if (rc) return rc;
    rc = run_panels(hp, parts, rows, N, C, dC,
        [&](uint32_t p) -> int {
            const uint32_t i0 = blockRowPtrs[bs[p]], i1 = blockRowPtrs[bs[p + 1]];
            if (i1 > i0) {
                CUSPMM_CUDA(cudaMemcpyAsync(dCol + i0, blockColIdxs + i0, (size_t)(i1 - i0) * 4, cudaMemcpyHostToDevice, hp.up));
                CUSPMM_CUDA(cudaMemcpyAsync(dBlk + (size_t)i0 * be, blocks + (size_t)i0 * be, (size_t)(i1 - i0) * be * 4,
                                            cudaMemcpyHostToDevice, hp.up));
            }
            return CUSPMM_OK;
        },
        [&](uint32_t p) -> int {
            if (!tc)
                return cuspmm_spmm_bsr_f32(dPtr + bs[p], dCol, dBlk, bs[p + 1] - bs[p], br, bc, K, dB, N, dldb,
                                           dC + (size_t)rows[p] * N, N, hp.run);
            int rc2 = bsr_tc_plan_create_impl(&hp.tc, dPtr, dCol, dBlk, numBlockRows, nb, br, K, N,
                                              variant == 2 ? CUSPMM_BLK_BF16 : CUSPMM_BLK_FP16, hp.run, hp.tcBlocks.p, hp.tcB.p);
            if (rc2) return rc2;
            rc2 = cuspmm_bsr_tc_prepare_B(hp.tc, dB, N, dldb, hp.run);
            if (rc2) return rc2;
            return cuspmm_bsr_tc_run(hp.tc, dC, N, hp.run);
        });
    if (rc == CUSPMM_OK) rc = host_end(hp, device_ms);
    if (hp.tc) {                       // the plan borrows the staging buffers: it does not outlive the call
        cudaStreamSynchronize(hp.run);
        cuspmm_bsr_tc_plan_destroy(hp.tc);
        hp.tc = nullptr;
    }
    return rc;
}

namespace cuspmm_b200 { void tc_pool_trim(int dev); }       // spmm_csr_tc.cu: the pools of the tiled copies of B

extern "C" int cuspmm_host_pipeline_release(int device) {
    cuspmm_b200::tc_pool_trim(device);
    std::lock_guard<std::mutex> lock(g_pipes_mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto it = g_pipes.begin(); it != g_pipes.end();) {
        if (device >= 0 && it->first != device) { ++it; continue; }
        std::lock_guard<std::mutex> plock(it->second->mu);          // waits for a call in flight on that device
        cudaSetDevice(it->first);
        it->second->destroy();
        ++it;
    }
    cudaSetDevice(cur);
    return CUSPMM_OK;
}

// =============================================================================== multi-GPU
enum MgpuFmt { MG_CSR = 0, MG_COO = 1, MG_SELL = 2, MG_BSR = 3 };

struct MgpuPanel {
    int dev = 0;
    uint32_t r0 = 0, r1 = 0;          // C rows of the panel
    uint32_t u0 = 0, u1 = 0;          // units: rows (CSR / COO), slices (ELL), block rows (BSR)
    uint32_t cnt = 0;                 // non-zeros / slots / blocks of the panel
    uint32_t *ptrs = nullptr;         // rebased rowPtrs / slicePtrs / blockRowPtrs (CSR, ELL, BSR); COO: rebased rowIdxs
    uint32_t *cols = nullptr;
    float *vals = nullptr, *B = nullptr, *C = nullptr;
    void *ws = nullptr;               // COO: (rows + 1) * 4 bytes
    size_t wsBytes = 0;
    cuspmmBsrTcPlan tc = nullptr;     // BSR tensor-core plan of the panel (built on first use)
    int tcType = -1;
    bool tcB = false;                 // the plan holds the current B
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};

struct cuspmmMgpuPlan_s {
    std::vector<MgpuPanel> p;
    std::vector<uint32_t> splits;     // in C rows
    MgpuFmt fmt = MG_CSR;
    uint32_t M = 0, K = 0, maxN = 0, N = 0, br = 1, bc = 1;
    uint64_t total = 0;
    float *C0 = nullptr;              // full C on device p[0].dev (gather target)
    bool peer = false;
    bool lastGather = false;
};

namespace {

struct PlanGuard {            // an error return must not leak the half-built plan
    cuspmmMgpuPlan_s *&p;
    bool armed = true;
    ~PlanGuard() { if (armed && p) { cuspmm_mgpu_destroy(p); p = nullptr; } }
};

int mgpu_begin(cuspmmMgpuPlan_s *pl, int ngpus, const int *devices) {
    int count = 0;
    CUSPMM_CUDA(cudaGetDeviceCount(&count));
    CUSPMM_REQUIRE(ngpus <= count, "asked for %d GPUs, %d visible", ngpus, count);
    pl->p.resize(ngpus);
    for (int g = 0; g < ngpus; ++g) pl->p[g].dev = devices ? devices[g] : g;
    // peer access (all pairs that support it)
    pl->peer = true;
    for (int a = 0; a < ngpus; ++a)
        for (int b = 0; b < ngpus; ++b) {
            if (a == b) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, pl->p[a].dev, pl->p[b].dev);
            if (!can) { pl->peer = false; continue; }
            cudaSetDevice(pl->p[a].dev);
            cudaError_t e = cudaDeviceEnablePeerAccess(pl->p[b].dev, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) pl->peer = false;
            cudaGetLastError();
        }
    return CUSPMM_OK;
}

// per panel: stream, events, B and C buffers, the three format arrays (ptrs has nptr entries, cols cnt, vals cnt * valsPer)
int mgpu_alloc_panel(cuspmmMgpuPlan_s *pl, MgpuPanel &q, size_t nptr, size_t valsPer) {
    const uint32_t rows = q.r1 - q.r0;
    CUSPMM_CUDA(cudaSetDevice(q.dev));
    CUSPMM_CUDA(cudaStreamCreateWithFlags(&q.st, cudaStreamNonBlocking));
    CUSPMM_CUDA(cudaEventCreate(&q.e0));
    CUSPMM_CUDA(cudaEventCreate(&q.e1));
    CUSPMM_CUDA(cudaMalloc(&q.ptrs, std::max<size_t>(nptr, 1) * 4));
    CUSPMM_CUDA(cudaMalloc(&q.cols, (size_t)std::max(q.cnt, 1u) * 4));
    CUSPMM_CUDA(cudaMalloc(&q.vals, (size_t)std::max(q.cnt, 1u) * valsPer * 4));
    CUSPMM_CUDA(cudaMalloc(&q.B, (size_t)pl->K * pl->maxN * 4));
    CUSPMM_CUDA(cudaMalloc(&q.C, (size_t)std::max(rows, 1u) * pl->maxN * 4));
    return CUSPMM_OK;
}

int mgpu_finish(cuspmmMgpuPlan_s *pl) {
    CUSPMM_CUDA(cudaSetDevice(pl->p[0].dev));
    CUSPMM_CUDA(cudaMalloc(&pl->C0, (size_t)std::max(pl->M, 1u) * pl->maxN * 4));
    for (auto &q : pl->p) { CUSPMM_CUDA(cudaSetDevice(q.dev)); CUSPMM_CUDA(cudaStreamSynchronize(q.st)); }
    CUSPMM_CUDA(cudaSetDevice(pl->p[0].dev));
    return CUSPMM_OK;
}

int rebase(uint32_t *p, uint32_t n, uint32_t base, cudaStream_t st) {
    if (n == 0 || base == 0) return CUSPMM_OK;
    rebase_kernel<<<(n + 255) / 256, 256, 0, st>>>(p, n, base);
    CUSPMM_LAUNCH_CHECK("rebase_kernel");
    return CUSPMM_OK;
}

} // namespace

extern "C" int cuspmm_mgpu_create_csr(cuspmmMgpuPlan *out, int ngpus, const int *devices, const uint32_t *rowPtrs,
                                      const uint32_t *colIdxs, const float *vals, uint32_t M, uint32_t K, uint32_t nnz,
                                      uint32_t maxN) {
    CUSPMM_REQUIRE(out && ngpus >= 1 && rowPtrs && maxN >= 1, "bad arguments");
    auto *pl = new cuspmmMgpuPlan_s();
    PlanGuard guard{pl};
    pl->fmt = MG_CSR; pl->M = M; pl->K = K; pl->total = nnz; pl->maxN = maxN;
    int rc = mgpu_begin(pl, ngpus, devices);
    if (rc) return rc;
    // split points from the device partitioner (on the first GPU)
    pl->splits.assign(ngpus + 1, 0);
    {
        CUSPMM_CUDA(cudaSetDevice(pl->p[0].dev));
        uint32_t *dRow = nullptr;
        CUSPMM_CUDA(cudaMalloc(&dRow, (size_t)(M + 1) * 4));
        CUSPMM_CUDA(cudaMemcpy(dRow, rowPtrs, (size_t)(M + 1) * 4, cudaMemcpyHostToDevice));
        rc = cuspmm_partition_rows_by_nnz(dRow, M, nnz, (uint32_t)ngpus, pl->splits.data(), nullptr);
        cudaFree(dRow);
        if (rc) return rc;
    }
    for (int g = 0; g < ngpus; ++g) {
        MgpuPanel &q = pl->p[g];
        q.r0 = q.u0 = pl->splits[g]; q.r1 = q.u1 = pl->splits[g + 1];
        const uint32_t i0 = rowPtrs[q.r0], i1 = rowPtrs[q.r1], rows = q.r1 - q.r0;
        q.cnt = i1 - i0;
        rc = mgpu_alloc_panel(pl, q, rows + 1, 1);
        if (rc) return rc;
        CUSPMM_CUDA(cudaMemcpyAsync(q.ptrs, rowPtrs + q.r0, (size_t)(rows + 1) * 4, cudaMemcpyHostToDevice, q.st));
        if (q.cnt) {
            CUSPMM_CUDA(cudaMemcpyAsync(q.cols, colIdxs + i0, (size_t)q.cnt * 4, cudaMemcpyHostToDevice, q.st));
            CUSPMM_CUDA(cudaMemcpyAsync(q.vals, vals + i0, (size_t)q.cnt * 4, cudaMemcpyHostToDevice, q.st));
        }
        rc = rebase(q.ptrs, rows + 1, i0, q.st);
        if (rc) return rc;
    }
    rc = mgpu_finish(pl);
    if (rc) return rc;
    guard.armed = false;
    *out = pl;
    return CUSPMM_OK;
}

extern "C" int cuspmm_mgpu_create_coo(cuspmmMgpuPlan *out, int ngpus, const int *devices, const uint32_t *rowIdxs,
                                      const uint32_t *colIdxs, const float *vals, uint32_t M, uint32_t K, uint32_t nnz,
                                      uint32_t maxN) {
    CUSPMM_REQUIRE(out && ngpus >= 1 && maxN >= 1 && (nnz == 0 || (rowIdxs && colIdxs && vals)), "bad arguments");
    auto *pl = new cuspmmMgpuPlan_s();
    PlanGuard guard{pl};
    pl->fmt = MG_COO; pl->M = M; pl->K = K; pl->total = nnz; pl->maxN = maxN;
    int rc = mgpu_begin(pl, ngpus, devices);
    if (rc) return rc;
    std::vector<uint32_t> idx;
    coo_splits(rowIdxs, M, nnz, (uint32_t)ngpus, /*taper=*/false, pl->splits, idx);     // the CSR rule at row boundaries
    for (int g = 0; g < ngpus; ++g) {
        MgpuPanel &q = pl->p[g];
        q.r0 = q.u0 = pl->splits[g]; q.r1 = q.u1 = pl->splits[g + 1];
        const uint32_t i0 = idx[g], i1 = idx[g + 1], rows = q.r1 - q.r0;
        q.cnt = i1 - i0;
        rc = mgpu_alloc_panel(pl, q, q.cnt, 1);            // ptrs = the panel's rowIdxs, rebased to its first row
        if (rc) return rc;
        q.wsBytes = (size_t)(rows + 1) * 4;
        CUSPMM_CUDA(cudaMalloc(&q.ws, q.wsBytes));
        if (q.cnt) {
            CUSPMM_CUDA(cudaMemcpyAsync(q.ptrs, rowIdxs + i0, (size_t)q.cnt * 4, cudaMemcpyHostToDevice, q.st));
            CUSPMM_CUDA(cudaMemcpyAsync(q.cols, colIdxs + i0, (size_t)q.cnt * 4, cudaMemcpyHostToDevice, q.st));
            CUSPMM_CUDA(cudaMemcpyAsync(q.vals, vals + i0, (size_t)q.cnt * 4, cudaMemcpyHostToDevice, q.st));
        }
        rc = rebase(q.ptrs, q.cnt, q.r0, q.st);
        if (rc) return rc;
    }
    rc = mgpu_finish(pl);
    if (rc) return rc;
    guard.armed = false;
    *out = pl;
    return CUSPMM_OK;
}

extern "C" int cuspmm_mgpu_create_sell(cuspmmMgpuPlan *out, int ngpus, const int *devices, const uint32_t *slicePtrs,
                                       const uint32_t *colIdxs, const float *vals, uint32_t M, uint32_t K, uint32_t sliceH,
                                       uint32_t numSlots, uint32_t maxN) {
    CUSPMM_REQUIRE(out && ngpus >= 1 && slicePtrs && maxN >= 1, "bad arguments");
    CUSPMM_REQUIRE(sliceH == 32, "sliced ELL kernels are built for slices of 32 rows (got %u)", sliceH);
    auto *pl = new cuspmmMgpuPlan_s();
    PlanGuard guard{pl};
    pl->fmt = MG_SELL; pl->M = M; pl->K = K; pl->total = numSlots; pl->maxN = maxN;
    int rc = mgpu_begin(pl, ngpus, devices);
    if (rc) return rc;
    const uint32_t slices = (M + 31) / 32;
    const std::vector<uint32_t> ss = prefix_splits(slicePtrs, slices, (uint32_t)ngpus, /*taper=*/false);   // balanced by slots
    pl->splits.assign(ngpus + 1, 0);
    for (int g = 0; g <= ngpus; ++g) pl->splits[g] = (uint32_t)std::min<uint64_t>(M, (uint64_t)ss[g] * 32);
    for (int g = 0; g < ngpus; ++g) {
        MgpuPanel &q = pl->p[g];
        q.r0 = pl->splits[g]; q.r1 = pl->splits[g + 1];
        q.u0 = ss[g]; q.u1 = ss[g + 1];
        const uint32_t i0 = slicePtrs[q.u0], i1 = slicePtrs[q.u1], ns = q.u1 - q.u0;
        q.cnt = i1 - i0;
        rc = mgpu_alloc_panel(pl, q, ns + 1, 1);
        if (rc) return rc;
        CUSPMM_CUDA(cudaMemcpyAsync(q.ptrs, slicePtrs + q.u0, (size_t)(ns + 1) * 4, cudaMemcpyHostToDevice, q.st));
        if (q.cnt) {
            CUSPMM_CUDA(cudaMemcpyAsync(q.cols, colIdxs + i0, (size_t)q.cnt * 4, cudaMemcpyHostToDevice, q.st));
            CUSPMM_CUDA(cudaMemcpyAsync(q.vals, vals + i0, (size_t)q.cnt * 4, cudaMemcpyHostToDevice, q.st));
        }
        rc = rebase(q.ptrs, ns + 1, i0, q.st);
        if (rc) return rc;
    }
    rc = mgpu_finish(pl);
    if (rc) return rc;
    guard.armed = false;
    *out = pl;
    return CUSPMM_OK;
}

extern "C" int cuspmm_mgpu_create_bsr(cuspmmMgpuPlan *out, int ngpus, const int *devices, const uint32_t *blockRowPtrs,
                                      const uint32_t *blockColIdxs, const float *blocks, uint32_t numBlockRows, uint32_t br,
                                      uint32_t bc, uint32_t K, uint32_t maxN) {
    CUSPMM_REQUIRE(out && ngpus >= 1 && blockRowPtrs && maxN >= 1 && br > 0 && bc > 0, "bad arguments");
    const uint64_t M64 = (uint64_t)numBlockRows * br;
    CUSPMM_REQUIRE(M64 <= 0xFFFFFFFFull, "numBlockRows * br overflows uint32");
    auto *pl = new cuspmmMgpuPlan_s();
    PlanGuard guard{pl};
    pl->fmt = MG_BSR; pl->M = (uint32_t)M64; pl->K = K; pl->total = blockRowPtrs[numBlockRows]; pl->maxN = maxN;
    pl->br = br; pl->bc = bc;
    int rc = mgpu_begin(pl, ngpus, devices);
    if (rc) return rc;
    const std::vector<uint32_t> bs = prefix_splits(blockRowPtrs, numBlockRows, (uint32_t)ngpus, /*taper=*/false);  // balanced by blocks
    pl->splits.assign(ngpus + 1, 0);
    for (int g = 0; g <= ngpus; ++g) pl->splits[g] = bs[g] * br;
    const size_t be = (size_t)br * bc;
    for (int g = 0; g < ngpus; ++g) {
        MgpuPanel &q = pl->p[g];
        q.r0 = pl->splits[g]; q.r1 = pl->splits[g + 1];
        q.u0 = bs[g]; q.u1 = bs[g + 1];
        const uint32_t i0 = blockRowPtrs[q.u0], i1 = blockRowPtrs[q.u1], nbr = q.u1 - q.u0;
        q.cnt = i1 - i0;
        rc = mgpu_alloc_panel(pl, q, nbr + 1, be);
        if (rc) return rc;
        CUSPMM_CUDA(cudaMemcpyAsync(q.ptrs, blockRowPtrs + q.u0, (size_t)(nbr + 1) * 4, cudaMemcpyHostToDevice, q.st));
        if (q.cnt) {
            CUSPMM_CUDA(cudaMemcpyAsync(q.cols, blockColIdxs + i0, (size_t)q.cnt * 4, cudaMemcpyHostToDevice, q.st));
            CUSPMM_CUDA(cudaMemcpyAsync(q.vals, blocks + (size_t)i0 * be, (size_t)q.cnt * be * 4, cudaMemcpyHostToDevice, q.st));
        }
        rc = rebase(q.ptrs, nbr + 1, i0, q.st);
        if (rc) return rc;
    }
    rc = mgpu_finish(pl);
    if (rc) return rc;
    guard.armed = false;
    *out = pl;
    return CUSPMM_OK;
}

extern "C" int cuspmm_mgpu_set_B(cuspmmMgpuPlan pl, const float *B, uint32_t N) {
    CUSPMM_REQUIRE(pl && B && N >= 1 && N <= pl->maxN, "bad arguments (N=%u, maxN=%u)", N, pl ? pl->maxN : 0);
    pl->N = N;
    const size_t bytes = (size_t)pl->K * N * 4;
    MgpuPanel &root = pl->p[0];
    CUSPMM_CUDA(cudaSetDevice(root.dev));
    CUSPMM_CUDA(cudaMemcpyAsync(root.B, B, bytes, cudaMemcpyHostToDevice, root.st));
    CUSPMM_CUDA(cudaStreamSynchronize(root.st));
    for (size_t g = 1; g < pl->p.size(); ++g) {
        MgpuPanel &q = pl->p[g];
        CUSPMM_CUDA(cudaSetDevice(q.dev));
        if (pl->peer) CUSPMM_CUDA(cudaMemcpyPeerAsync(q.B, q.dev, root.B, root.dev, bytes, q.st));   // NVLink
        else CUSPMM_CUDA(cudaMemcpyAsync(q.B, B, bytes, cudaMemcpyHostToDevice, q.st));
    }
    for (auto &q : pl->p) { CUSPMM_CUDA(cudaSetDevice(q.dev)); CUSPMM_CUDA(cudaStreamSynchronize(q.st)); q.tcB = false; }
    CUSPMM_CUDA(cudaSetDevice(root.dev));
    return CUSPMM_OK;
}

namespace cuspmm_b200 { int csr_select_variant(uint32_t M, uint32_t K, uint64_t nnz, uint32_t N, bool vec_ok, bool sell); }   // spmm_csr.cu

// Will this multiply run the tensor-core CSR kernel?  That kernel ACCUMULATES into C (memset + red.add of partial tiles every 64
// chunks): aimed at another GPU's memory that is dozens of round trips of atomics over NVLink per element (measured: 2.5 ms on one
// GPU, 4.9 ms on eight with the peer-store gather).  Such panels are computed into the panel's own C and sent with one peer copy.
static bool mgpu_uses_tensor_kernel(cuspmmMgpuPlan pl, const MgpuPanel &q, int variant) {
    const bool vok = pl->N % 4 == 0;
    switch (pl->fmt) {
    case MG_CSR:
        return variant == 8 || (variant == 0 && cuspmm_b200::csr_select_variant(q.r1 - q.r0, pl->K, q.cnt, pl->N, vok, false) == 8);
    case MG_COO:
        return (variant == 0 || variant == 2) && cuspmm_b200::csr_select_variant(q.r1 - q.r0, pl->K, q.cnt, pl->N, vok, false) == 8;
    case MG_SELL:
        return variant == 6 || (variant == 0 && cuspmm_b200::csr_select_variant(q.r1 - q.r0, pl->K, q.cnt, pl->N, vok, true) == 8);
    default:
        return false;
    }
}

// one multiply of panel q into Cdst on its stream (the current device is q.dev)
static int mgpu_launch(cuspmmMgpuPlan pl, MgpuPanel &q, int variant, float *Cdst) {
    const uint32_t N = pl->N, rows = q.r1 - q.r0;
    if (Cdst != q.C && q.dev != pl->p[0].dev && mgpu_uses_tensor_kernel(pl, q, variant)) {
        int rc = mgpu_launch(pl, q, variant, q.C);
        if (rc) return rc;
        CUSPMM_CUDA(cudaMemcpyPeerAsync(Cdst, pl->p[0].dev, q.C, q.dev, (size_t)rows * N * sizeof(float), q.st));
        return CUSPMM_OK;
    }
    switch (pl->fmt) {
    case MG_CSR:
        return spmm_csr_dispatch(q.ptrs, q.cols, q.vals, rows, pl->K, q.cnt, q.B, N, N, Cdst, N, variant, q.st);
    case MG_COO:
        return cuspmm_spmm_coo(q.ptrs, q.cols, q.vals, rows, pl->K, q.cnt, q.B, N, N, Cdst, N, variant, q.ws, q.wsBytes, q.st);
    case MG_SELL:
        return cuspmm_spmm_sell(q.ptrs, q.cols, q.vals, rows, pl->K, 32, q.cnt, q.B, N, N, Cdst, N, variant, q.st);
    case MG_BSR: {
        if (variant <= 1)
            return cuspmm_spmm_bsr_f32(q.ptrs, q.cols, q.vals, q.u1 - q.u0, pl->br, pl->bc, pl->K, q.B, N, N, Cdst, N, q.st);
        const int type = variant == 2 ? (int)CUSPMM_BLK_BF16 : (int)CUSPMM_BLK_FP16;
        if (q.tc && q.tcType != type) { cuspmm_bsr_tc_plan_destroy(q.tc); q.tc = nullptr; }
        if (!q.tc) {
            int rc = cuspmm_bsr_tc_plan_create(&q.tc, q.ptrs, q.cols, q.vals, q.u1 - q.u0, q.cnt, pl->br, pl->K, pl->maxN,
                                               (cuspmmBlockType)type, q.st);
            if (rc) return rc;
            q.tcType = type;
            q.tcB = false;
        }
        if (!q.tcB) {
            int rc = cuspmm_bsr_tc_prepare_B(q.tc, q.B, N, N, q.st);
            if (rc) return rc;
            q.tcB = true;
        }
        return cuspmm_bsr_tc_run(q.tc, Cdst, N, q.st);
    }
    }
    return set_error(CUSPMM_ERR_INVALID, "unknown plan format");
}

extern "C" int cuspmm_mgpu_run(cuspmmMgpuPlan pl, int variant, int gather, int iters, float *max_device_ms) {
    CUSPMM_REQUIRE(pl && pl->N >= 1 && iters >= 1, "set_B must be called first");
    if (gather && !pl->peer && pl->p.size() > 1)
        return set_error(CUSPMM_ERR_UNSUPPORTED, "gather needs peer access between all GPUs of the plan");
    if (pl->fmt == MG_BSR && variant >= 2 && !(pl->br == pl->bc && (pl->br == 16 || pl->br == 32)))
        return set_error(CUSPMM_ERR_UNSUPPORTED, "tensor-core BSR supports 16x16 and 32x32 blocks (got %ux%u)", pl->br, pl->bc);
    const uint32_t N = pl->N;
    for (auto &q : pl->p) {
        CUSPMM_CUDA(cudaSetDevice(q.dev));
        const uint32_t rows = q.r1 - q.r0;
        float *Cdst = gather ? pl->C0 + (size_t)q.r0 * N : q.C;
        if (rows && pl->fmt == MG_BSR && variant >= 2) {       // plan construction and the cast of B stay outside the timed launches
            int rc = mgpu_launch(pl, q, variant, Cdst);
            if (rc) return rc;
        }
        CUSPMM_CUDA(cudaEventRecord(q.e0, q.st));
        for (int it = 0; it < iters; ++it) {
            if (rows == 0) break;
            int rc = mgpu_launch(pl, q, variant, Cdst);
            if (rc) return rc;
        }
        CUSPMM_CUDA(cudaEventRecord(q.e1, q.st));
    }
    float worst = 0.f;
    for (auto &q : pl->p) {
        CUSPMM_CUDA(cudaSetDevice(q.dev));
        CUSPMM_CUDA(cudaEventSynchronize(q.e1));
        float ms = 0.f;
        CUSPMM_CUDA(cudaEventElapsedTime(&ms, q.e0, q.e1));
        worst = std::max(worst, ms / iters);
    }
    pl->lastGather = gather != 0;
    CUSPMM_CUDA(cudaSetDevice(pl->p[0].dev));
    if (max_device_ms) *max_device_ms = worst;
    return CUSPMM_OK;
}

extern "C" int cuspmm_mgpu_get_splits(cuspmmMgpuPlan pl, uint32_t *splits) {
    CUSPMM_REQUIRE(pl && splits, "null pointer");
    std::copy(pl->splits.begin(), pl->splits.end(), splits);
    return CUSPMM_OK;
}

extern "C" int cuspmm_mgpu_get_counts(cuspmmMgpuPlan pl, uint32_t *counts) {
    CUSPMM_REQUIRE(pl && counts, "null pointer");
    for (size_t g = 0; g < pl->p.size(); ++g) counts[g] = pl->p[g].cnt;
    return CUSPMM_OK;
}

extern "C" int cuspmm_mgpu_get_C(cuspmmMgpuPlan pl, float *C) {
    CUSPMM_REQUIRE(pl && C && pl->N, "bad arguments");
    const uint32_t N = pl->N;
    if (pl->lastGather) {
        CUSPMM_CUDA(cudaSetDevice(pl->p[0].dev));
        CUSPMM_CUDA(cudaMemcpy(C, pl->C0, (size_t)pl->M * N * 4, cudaMemcpyDeviceToHost));
        return CUSPMM_OK;
    }
    for (auto &q : pl->p) {
        CUSPMM_CUDA(cudaSetDevice(q.dev));
        if (q.r1 > q.r0)
            CUSPMM_CUDA(cudaMemcpy(C + (size_t)q.r0 * N, q.C, (size_t)(q.r1 - q.r0) * N * 4, cudaMemcpyDeviceToHost));
    }
    CUSPMM_CUDA(cudaSetDevice(pl->p[0].dev));
    return CUSPMM_OK;
}

extern "C" int cuspmm_mgpu_destroy(cuspmmMgpuPlan pl) {
    if (!pl) return CUSPMM_OK;
    for (auto &q : pl->p) {
        cudaSetDevice(q.dev);
        if (q.st) cudaStreamSynchronize(q.st);
        if (q.tc) cuspmm_bsr_tc_plan_destroy(q.tc);
        cudaFree(q.ptrs); cudaFree(q.cols); cudaFree(q.vals); cudaFree(q.B); cudaFree(q.C); cudaFree(q.ws);
        if (q.st) cudaStreamDestroy(q.st);
        if (q.e0) cudaEventDestroy(q.e0);
        if (q.e1) cudaEventDestroy(q.e1);
    }
    if (!pl->p.empty()) cudaSetDevice(pl->p[0].dev);
    cudaFree(pl->C0);
    delete pl;
    return CUSPMM_OK;
}
