// host_api.cu -- host-buffer entry point and the multi-GPU row-panel engine.
//
//  * cuspmm_spmm_csr_host: what runEngine + spmmCSRWrapper<k> do end to end in the reference
//    (src/engine/engine.cpp:20-44: H2D of A, B, C; kernel; D2H of C), but pipelined: A is cut
//    into nnz-balanced row panels and panel p+1's H2D overlaps panel p's kernel and panel
//    p-1's D2H on three streams.  Device buffers are cached per device (grow-only), so the
//    "prolog" cudaMalloc+memset of the reference (spmm_csr_k3.cu:69-73) is gone.
//  * cuspmm_mgpu_*: north_star (c) -- A split into nnz-balanced row panels over the GPUs of
//    one node (split points from the DEVICE partitioner), B replicated over NVLink with peer
//    copies, each GPU multiplies its panel on its own stream; with gather the kernels store
//    their C rows directly into GPU 0's C through peer memory (compute and "collective" are
//    one kernel: the stores travel over NVLink while the panel is still being computed).
#include "common.cuh"

#include <algorithm>
#include <mutex>
#include <vector>

namespace cuspmm_b200 {

int spmm_csr_dispatch(const uint32_t *, const uint32_t *, const float *, uint32_t, uint32_t, uint32_t,
                      const float *, uint32_t, size_t, float *, size_t, int, cudaStream_t);

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
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
};

struct HostPipe {
    GrowBuf rowPtrs, colIdxs, vals, B, C;
    cudaStream_t up = nullptr, run = nullptr, down = nullptr;
    std::vector<cudaEvent_t> upDone, runDone;
    cudaEvent_t t0 = nullptr, t1 = nullptr, bDone = nullptr;
    bool init = false;
};
static HostPipe g_pipe[16];
static std::mutex g_pipe_mutex[16];    // the per-device pipeline (streams, cached buffers) is used by one call at a time

// Split points by nnz.  taper = false: equal shares (the multi-GPU rule).  taper = true: shares shrink
// linearly towards the end (the host pipeline: the copies are the bottleneck, so what matters is how much
// kernel + D2H is left once the last H2D lands -- a small last panel keeps that tail short).
static std::vector<uint32_t> host_splits(const uint32_t *rowPtrs, uint32_t M, uint32_t nnz, uint32_t parts,
                                         bool taper = false) {
    std::vector<uint32_t> s(parts + 1, 0);
    const uint64_t wsum = (uint64_t)parts * (parts + 1) / 2;      // weights parts, parts-1, ..., 1
    uint64_t wacc = 0;
    for (uint32_t g = 1; g < parts; ++g) {
        wacc += parts - (g - 1);
        const uint64_t t = taper ? ((uint64_t)nnz * wacc) / wsum : ((uint64_t)g * nnz) / parts;
        s[g] = (uint32_t)(std::lower_bound(rowPtrs, rowPtrs + M + 1, (uint32_t)t) - rowPtrs);
        if (s[g] > M) s[g] = M;
        if (s[g] < s[g - 1]) s[g] = s[g - 1];
    }
    s[parts] = M;
    return s;
}

} // namespace cuspmm_b200

using namespace cuspmm_b200;

extern "C" int cuspmm_spmm_csr_host(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals,
                                    uint32_t M, uint32_t K, uint32_t nnz, const float *B, uint32_t N, float *C,
                                    int variant, float *device_ms) {
    CUSPMM_REQUIRE(rowPtrs && B && C && (nnz == 0 || (colIdxs && vals)), "null operand pointer");
    if (M == 0 || N == 0) return CUSPMM_OK;
    int dev = 0;
    CUSPMM_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_pipe_mutex[dev & 15]);
    HostPipe &hp = g_pipe[dev & 15];
    constexpr int kMaxPanels = 16;
    if (!hp.init) {
        CUSPMM_CUDA(cudaStreamCreateWithFlags(&hp.up, cudaStreamNonBlocking));
        CUSPMM_CUDA(cudaStreamCreateWithFlags(&hp.run, cudaStreamNonBlocking));
        CUSPMM_CUDA(cudaStreamCreateWithFlags(&hp.down, cudaStreamNonBlocking));
        hp.upDone.resize(kMaxPanels);
        hp.runDone.resize(kMaxPanels);
        for (int i = 0; i < kMaxPanels; ++i) {
            CUSPMM_CUDA(cudaEventCreateWithFlags(&hp.upDone[i], cudaEventDisableTiming));
            CUSPMM_CUDA(cudaEventCreateWithFlags(&hp.runDone[i], cudaEventDisableTiming));
        }
        CUSPMM_CUDA(cudaEventCreate(&hp.t0));
        CUSPMM_CUDA(cudaEventCreate(&hp.t1));
        CUSPMM_CUDA(cudaEventCreateWithFlags(&hp.bDone, cudaEventDisableTiming));
        hp.init = true;
    }
    CUSPMM_CUDA(hp.rowPtrs.reserve((size_t)(M + 1) * 4));
    CUSPMM_CUDA(hp.colIdxs.reserve((size_t)std::max(nnz, 1u) * 4));
    CUSPMM_CUDA(hp.vals.reserve((size_t)std::max(nnz, 1u) * 4));
    CUSPMM_CUDA(hp.B.reserve((size_t)K * N * 4));
    CUSPMM_CUDA(hp.C.reserve((size_t)M * N * 4));
    uint32_t *dRow = static_cast<uint32_t *>(hp.rowPtrs.p), *dCol = static_cast<uint32_t *>(hp.colIdxs.p);
    float *dVal = static_cast<float *>(hp.vals.p), *dB = static_cast<float *>(hp.B.p), *dC = static_cast<float *>(hp.C.p);

    // panels: the staged kernel streams all of B once per 60-row CTA, so a panel must still give
    // every SM a full CTA (>= ~6400 rows: 16 thin panels made the kernels 3x slower than the copies);
    // within that, more panels = finer overlap of H2D(p+1) / kernel(p) / D2H(p-1)
    const size_t aBytes = (size_t)nnz * 8;
    uint32_t parts = (uint32_t)std::min<size_t>(kMaxPanels, std::max<size_t>(1, aBytes / (32u << 20)));
    parts = std::min(parts, std::max(1u, (M + 3200u) / 6400u));
    if (parts >= 3) ++parts;     // tapering makes the early panels bigger, so one more of them is free
    const std::vector<uint32_t> sp = host_splits(rowPtrs, M, nnz, parts, /*taper=*/true);

    CUSPMM_CUDA(cudaEventRecord(hp.t0, hp.up));
    CUSPMM_CUDA(cudaMemcpyAsync(dRow, rowPtrs, (size_t)(M + 1) * 4, cudaMemcpyHostToDevice, hp.up));
    CUSPMM_CUDA(cudaMemcpyAsync(dB, B, (size_t)K * N * 4, cudaMemcpyHostToDevice, hp.up));
    for (uint32_t p = 0; p < parts; ++p) {
        const uint32_t r0 = sp[p], r1 = sp[p + 1];
        const uint32_t i0 = rowPtrs[r0], i1 = rowPtrs[r1];
        if (i1 > i0) {
            CUSPMM_CUDA(cudaMemcpyAsync(dCol + i0, colIdxs + i0, (size_t)(i1 - i0) * 4, cudaMemcpyHostToDevice, hp.up));
            CUSPMM_CUDA(cudaMemcpyAsync(dVal + i0, vals + i0, (size_t)(i1 - i0) * 4, cudaMemcpyHostToDevice, hp.up));
        }
        CUSPMM_CUDA(cudaEventRecord(hp.upDone[p], hp.up));
        if (r1 == r0) continue;
        CUSPMM_CUDA(cudaStreamWaitEvent(hp.run, hp.upDone[p], 0));
        int rc = spmm_csr_dispatch(dRow + r0, dCol, dVal, r1 - r0, K, i1 - i0, dB, N, N, dC + (size_t)r0 * N, N, variant, hp.run);
        if (rc) return rc;
        CUSPMM_CUDA(cudaEventRecord(hp.runDone[p], hp.run));
        CUSPMM_CUDA(cudaStreamWaitEvent(hp.down, hp.runDone[p], 0));
        CUSPMM_CUDA(cudaMemcpyAsync(C + (size_t)r0 * N, dC + (size_t)r0 * N, (size_t)(r1 - r0) * N * 4,
                                    cudaMemcpyDeviceToHost, hp.down));
    }
    CUSPMM_CUDA(cudaEventRecord(hp.bDone, hp.up));
    CUSPMM_CUDA(cudaStreamWaitEvent(hp.down, hp.bDone, 0));
    CUSPMM_CUDA(cudaEventRecord(hp.t1, hp.down));
    CUSPMM_CUDA(cudaStreamSynchronize(hp.down));
    CUSPMM_CUDA(cudaStreamSynchronize(hp.run));
    if (device_ms) CUSPMM_CUDA(cudaEventElapsedTime(device_ms, hp.t0, hp.t1));
    return CUSPMM_OK;
}

// =============================================================================== multi-GPU
struct MgpuPanel {
    int dev = 0;
    uint32_t r0 = 0, r1 = 0, nnz = 0;
    uint32_t *rowPtrs = nullptr, *colIdxs = nullptr;
    float *vals = nullptr, *B = nullptr, *C = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};

struct cuspmmMgpuPlan_s {
    std::vector<MgpuPanel> p;
    std::vector<uint32_t> splits;
    uint32_t M = 0, K = 0, nnz = 0, maxN = 0, N = 0;
    float *C0 = nullptr;       // full C on device p[0].dev (gather target)
    bool peer = false;
    bool lastGather = false;
};

extern "C" int cuspmm_mgpu_create_csr(cuspmmMgpuPlan *out, int ngpus, const int *devices, const uint32_t *rowPtrs,
                                      const uint32_t *colIdxs, const float *vals, uint32_t M, uint32_t K, uint32_t nnz,
                                      uint32_t maxN) {
    CUSPMM_REQUIRE(out && ngpus >= 1 && rowPtrs && maxN >= 1, "bad arguments");
    int count = 0;
    CUSPMM_CUDA(cudaGetDeviceCount(&count));
    CUSPMM_REQUIRE(ngpus <= count, "asked for %d GPUs, %d visible", ngpus, count);
    auto *pl = new cuspmmMgpuPlan_s();
    struct Guard {            // an error return below must not leak the half-built plan
        cuspmmMgpuPlan_s *&p;
        bool armed = true;
        ~Guard() { if (armed && p) { cuspmm_mgpu_destroy(p); p = nullptr; } }
    } guard{pl};
    pl->M = M; pl->K = K; pl->nnz = nnz; pl->maxN = maxN;
    pl->p.resize(ngpus);
    for (int g = 0; g < ngpus; ++g) pl->p[g].dev = devices ? devices[g] : g;

    // split points from the device partitioner (on the first GPU)
    pl->splits.assign(ngpus + 1, 0);
    {
        CUSPMM_CUDA(cudaSetDevice(pl->p[0].dev));
        uint32_t *dRow = nullptr;
        CUSPMM_CUDA(cudaMalloc(&dRow, (size_t)(M + 1) * 4));
        CUSPMM_CUDA(cudaMemcpy(dRow, rowPtrs, (size_t)(M + 1) * 4, cudaMemcpyHostToDevice));
        int rc = cuspmm_partition_rows_by_nnz(dRow, M, nnz, (uint32_t)ngpus, pl->splits.data(), nullptr);
        cudaFree(dRow);
        if (rc) return rc;
    }
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
    for (int g = 0; g < ngpus; ++g) {
        MgpuPanel &q = pl->p[g];
        q.r0 = pl->splits[g]; q.r1 = pl->splits[g + 1];
        const uint32_t i0 = rowPtrs[q.r0], i1 = rowPtrs[q.r1];
        q.nnz = i1 - i0;
        const uint32_t rows = q.r1 - q.r0;
        CUSPMM_CUDA(cudaSetDevice(q.dev));
        CUSPMM_CUDA(cudaStreamCreateWithFlags(&q.st, cudaStreamNonBlocking));
        CUSPMM_CUDA(cudaEventCreate(&q.e0));
        CUSPMM_CUDA(cudaEventCreate(&q.e1));
        CUSPMM_CUDA(cudaMalloc(&q.rowPtrs, (size_t)(rows + 1) * 4));
        CUSPMM_CUDA(cudaMalloc(&q.colIdxs, (size_t)std::max(q.nnz, 1u) * 4));
        CUSPMM_CUDA(cudaMalloc(&q.vals, (size_t)std::max(q.nnz, 1u) * 4));
        CUSPMM_CUDA(cudaMalloc(&q.B, (size_t)K * maxN * 4));
        CUSPMM_CUDA(cudaMalloc(&q.C, (size_t)std::max(rows, 1u) * maxN * 4));
        CUSPMM_CUDA(cudaMemcpyAsync(q.rowPtrs, rowPtrs + q.r0, (size_t)(rows + 1) * 4, cudaMemcpyHostToDevice, q.st));
        if (q.nnz) {
            CUSPMM_CUDA(cudaMemcpyAsync(q.colIdxs, colIdxs + i0, (size_t)q.nnz * 4, cudaMemcpyHostToDevice, q.st));
            CUSPMM_CUDA(cudaMemcpyAsync(q.vals, vals + i0, (size_t)q.nnz * 4, cudaMemcpyHostToDevice, q.st));
        }
        rebase_kernel<<<(rows + 1 + 255) / 256, 256, 0, q.st>>>(q.rowPtrs, rows + 1, i0);
        CUSPMM_LAUNCH_CHECK("rebase_kernel");
    }
    CUSPMM_CUDA(cudaSetDevice(pl->p[0].dev));
    CUSPMM_CUDA(cudaMalloc(&pl->C0, (size_t)std::max(M, 1u) * maxN * 4));
    for (auto &q : pl->p) { CUSPMM_CUDA(cudaSetDevice(q.dev)); CUSPMM_CUDA(cudaStreamSynchronize(q.st)); }
    CUSPMM_CUDA(cudaSetDevice(pl->p[0].dev));
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
    for (auto &q : pl->p) { CUSPMM_CUDA(cudaSetDevice(q.dev)); CUSPMM_CUDA(cudaStreamSynchronize(q.st)); }
    CUSPMM_CUDA(cudaSetDevice(root.dev));
    return CUSPMM_OK;
}

extern "C" int cuspmm_mgpu_run(cuspmmMgpuPlan pl, int variant, int gather, int iters, float *max_device_ms) {
    CUSPMM_REQUIRE(pl && pl->N >= 1 && iters >= 1, "set_B must be called first");
    if (gather && !pl->peer && pl->p.size() > 1)
        return set_error(CUSPMM_ERR_UNSUPPORTED, "gather needs peer access between all GPUs of the plan");
    const uint32_t N = pl->N;
    for (auto &q : pl->p) {
        CUSPMM_CUDA(cudaSetDevice(q.dev));
        const uint32_t rows = q.r1 - q.r0;
        float *Cdst = gather ? pl->C0 + (size_t)q.r0 * N : q.C;
        CUSPMM_CUDA(cudaEventRecord(q.e0, q.st));
        for (int it = 0; it < iters; ++it) {
            if (rows == 0) break;
            int rc = spmm_csr_dispatch(q.rowPtrs, q.colIdxs, q.vals, rows, pl->K, q.nnz, q.B, N, N, Cdst, N, variant, q.st);
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
        cudaFree(q.rowPtrs); cudaFree(q.colIdxs); cudaFree(q.vals); cudaFree(q.B); cudaFree(q.C);
        if (q.st) cudaStreamDestroy(q.st);
        if (q.e0) cudaEventDestroy(q.e0);
        if (q.e1) cudaEventDestroy(q.e1);
    }
    if (!pl->p.empty()) cudaSetDevice(pl->p[0].dev);
    cudaFree(pl->C0);
    delete pl;
    return CUSPMM_OK;
}
