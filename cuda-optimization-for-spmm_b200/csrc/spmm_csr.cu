// spmm_csr.cu -- CSR SpMM kernels for sm_100a.
//
// Replaces the reference's spmmCSRK1..K4 (src/spmm/csr/spmm_csr_k{1,2,3,4}.cu).  None of
// those designs is kept: every kernel here accumulates in fp32 with FMA in CSR order
// (fixed order, no atomics => bit-reproducible run to run), reads B row-major with
// 128-bit loads and never transposes B.
//
//   variant 1  csr_rowsplit_vec   warp per row, lanes over columns (float4), rows dealt to
//                                 warps in nnz-balanced contiguous ranges found by an
//                                 in-kernel 32-ary search on rowPtrs (no preprocessing pass)
//   variant 2  csr_subwarp_vec    "vector per row": G = 4/8/16 lanes per row, 64-column tiles
//   variant 3  csr_staged         row panel x K-chunks; each K-chunk of B is staged in shared
//                                 memory by TMA bulk copies (cp.async.bulk + mbarrier ring,
//                                 one producer warp) and re-used by all rows of the panel;
//                                 31 consumer warps x 2 rows (or 15 x 2 / 15 x 4, see launch_by_N)
//   variant 4  csr_rowsplit_scalar any N / ldb / alignment (N = 21 in data/small_210)
//   variant 5  csr_dual           variant 3's pipeline with tensor memory as a second operand port: the first rows of
//                                 every B chunk are copied into TMEM (tcgen05.cp) and gathered with tcgen05.ld.x16
//                                 instead of LDS (spmm_csr_tmem.cu)
//   variant 6  csr_nnzsplit       equal nnz ranges per warp, rows cut at range boundaries, ordered carry fix-up
//                                 (spmm_csr_split.cu; needs workspace: cuspmm_spmm_csr_ws)
// All row kernels are templated on the row accessor (RowRef<SELL>), so the same code runs on the
// sliced-ELL layout (spmm_ell.cu calls spmm_sell_rows_dispatch).
#include "common.cuh"

#include <atomic>

#include <stdlib.h>

namespace cuspmm_b200 {

// =============================================================== row access (CSR or sliced ELL)
// The row kernels below are written once against this accessor.  SELL = false: CSR, entry k of row r
// at rowPtrs[r] + k.  SELL = true: sliced ELL (ptrs = slicePtrs), entry k of row r at
// slicePtrs[r/32] + 32k + r%32, k < slice width; a row ends at its first padding entry (col kPad).
template <bool SELL>
struct RowRef {
    uint32_t first, len;
    static constexpr uint32_t stride = SELL ? 32u : 1u;
    __device__ __forceinline__ RowRef(const uint32_t *__restrict__ ptrs, uint32_t r) {
        if constexpr (SELL) {
            const uint32_t sb = __ldg(ptrs + (r >> 5));
            first = sb + (r & 31u);
            len = (__ldg(ptrs + (r >> 5) + 1) - sb) >> 5;
        } else {
            first = __ldg(ptrs + r);
            len = __ldg(ptrs + r + 1) - first;
        }
    }
    __device__ __forceinline__ size_t at(uint32_t k) const { return (size_t)first + (size_t)k * stride; }
};

// =============================================================== variant 1 / 4
// items = one per row + one per non-zero (slot for ELL); warp w owns the rows whose first item
// falls into [w*ipw, (w+1)*ipw).  Whole rows only, so no carries between warps.
template <bool SELL>
__device__ __forceinline__ void warp_row_range(const uint32_t *__restrict__ rowPtrs, uint32_t M,
                                               uint64_t ipw, uint32_t w, uint32_t &r0, uint32_t &r1) {
    // rowPtrs may be a row-panel VIEW of a larger matrix (rowPtrs[0] != 0, absolute offsets
    // into colIdxs/vals), so keys are taken relative to the panel's first entry.
    const uint32_t first = __ldg(rowPtrs);
    auto key = [&](uint32_t p) -> uint64_t {
        if constexpr (SELL) {     // slots before row p: whole slices + p%32 rows of its own slice
            const uint32_t sb = __ldg(rowPtrs + (p >> 5));
            const uint32_t w32 = (p & 31u) ? ((__ldg(rowPtrs + (p >> 5) + 1) - sb) >> 5) * (p & 31u) : 0u;
            return (uint64_t)(sb - first) + w32 + p;
        } else {
            return (uint64_t)(__ldg(rowPtrs + p) - first) + p;
        }
    };
    warp_lower_bound2(M, (uint64_t)w * ipw, (uint64_t)(w + 1) * ipw, key, r0, r1);
}

template <int U, int J, bool SELL>
__global__ void __launch_bounds__(256)
csr_rowsplit_vec_kernel(const uint32_t *__restrict__ rowPtrs, const uint32_t *__restrict__ colIdxs,
                        const float *__restrict__ vals, uint32_t M, uint64_t ipw,
                        const float *__restrict__ B, uint32_t N, size_t ldb,
                        float *__restrict__ C, size_t ldc) {
    const uint32_t lane = lane_id();
    const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t col0 = blockIdx.y * (128u * U) + lane * 4u;   // first column of this lane
    uint32_t r0, r1;
    warp_row_range<SELL>(rowPtrs, M, ipw, w, r0, r1);

    bool valid[U];
#pragma unroll
    for (int u = 0; u < U; ++u) valid[u] = (col0 + u * 128u) < N;

    for (uint32_t r = r0; r < r1; ++r) {
        const RowRef<SELL> row(rowPtrs, r);
        float4 acc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);

        for (uint32_t base = 0; base < row.len; base += 32) {
            const uint32_t k = base + lane;
            uint32_t mc = kPad;
            float mv = 0.f;
            if (k < row.len) { mc = ld_stream(colIdxs + row.at(k)); mv = ld_stream(vals + row.at(k)); }
            // valid entries form a prefix of the batch (ELL padding sits at the end of a row)
            const int cnt = SELL ? __popc(__ballot_sync(0xFFFFFFFFu, mc != kPad)) : (int)min(32u, row.len - base);
            int j = 0;
            for (; j + J <= cnt; j += J) {
                float4 b[J][U];
                float v[J];
#pragma unroll
                for (int kk = 0; kk < J; ++kk) {
                    const uint32_t c = __shfl_sync(0xFFFFFFFFu, mc, j + kk);
                    v[kk] = __shfl_sync(0xFFFFFFFFu, mv, j + kk);
                    const float *brow = B + (size_t)c * ldb + col0;
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (valid[u]) b[kk][u] = __ldg(reinterpret_cast<const float4 *>(brow + u * 128));
                }
#pragma unroll
                for (int kk = 0; kk < J; ++kk)
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (valid[u]) fma4(acc[u], v[kk], b[kk][u]);
            }
            for (; j < cnt; ++j) {
                const uint32_t c = __shfl_sync(0xFFFFFFFFu, mc, j);
                const float v = __shfl_sync(0xFFFFFFFFu, mv, j);
                const float *brow = B + (size_t)c * ldb + col0;
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (valid[u]) fma4(acc[u], v, __ldg(reinterpret_cast<const float4 *>(brow + u * 128)));
            }
            if (SELL && cnt < 32) break;      // reached the padding: the row is finished
        }
        float *crow = C + (size_t)r * ldc + col0;
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (valid[u]) __stcs(reinterpret_cast<float4 *>(crow + u * 128), acc[u]);
    }
}

template <int U, bool SELL>
__global__ void __launch_bounds__(256)
csr_rowsplit_scalar_kernel(const uint32_t *__restrict__ rowPtrs, const uint32_t *__restrict__ colIdxs,
                           const float *__restrict__ vals, uint32_t M, uint64_t ipw,
                           const float *__restrict__ B, uint32_t N, size_t ldb,
                           float *__restrict__ C, size_t ldc) {
    const uint32_t lane = lane_id();
    const uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t col0 = blockIdx.y * (32u * U) + lane;
    uint32_t r0, r1;
    warp_row_range<SELL>(rowPtrs, M, ipw, w, r0, r1);
    for (uint32_t r = r0; r < r1; ++r) {
        const RowRef<SELL> row(rowPtrs, r);
        float acc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) acc[u] = 0.f;
        for (uint32_t base = 0; base < row.len; base += 32) {
            const uint32_t k = base + lane;
            uint32_t mc = kPad;
            float mv = 0.f;
            if (k < row.len) { mc = ld_stream(colIdxs + row.at(k)); mv = ld_stream(vals + row.at(k)); }
            const int cnt = SELL ? __popc(__ballot_sync(0xFFFFFFFFu, mc != kPad)) : (int)min(32u, row.len - base);
            for (int j = 0; j < cnt; ++j) {
                const uint32_t c = __shfl_sync(0xFFFFFFFFu, mc, j);
                const float v = __shfl_sync(0xFFFFFFFFu, mv, j);
                const float *brow = B + (size_t)c * ldb;
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (col0 + u * 32u < N) acc[u] = fmaf(v, __ldg(brow + col0 + u * 32u), acc[u]);
            }
            if (SELL && cnt < 32) break;
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (col0 + u * 32u < N) C[(size_t)r * ldc + col0 + u * 32u] = acc[u];
    }
}

// =============================================================== variant 2
// G lanes per row (G*4 >= column tile), 32/G rows per warp; for narrow N a full warp
// per row would leave most lanes idle.  Rows are dealt round-robin-free: row = global
// group index (short-row regime, no balancing needed).
template <int G, bool SELL>
__global__ void __launch_bounds__(256)
csr_subwarp_vec_kernel(const uint32_t *__restrict__ rowPtrs, const uint32_t *__restrict__ colIdxs,
                       const float *__restrict__ vals, uint32_t M,
                       const float *__restrict__ B, uint32_t N, size_t ldb,
                       float *__restrict__ C, size_t ldc) {
    const uint32_t lane = lane_id();
    const uint32_t gl = lane % G;                    // lane inside the group
    const uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) / G;
    const uint32_t col = blockIdx.y * (4u * G) + gl * 4u;
    const bool valid = col < N;
    const uint32_t r = gid;
    const RowRef<SELL> row(rowPtrs, r < M ? r : 0u);
    const uint32_t len = r < M ? row.len : 0u;
    const uint32_t maxlen = __reduce_max_sync(0xFFFFFFFFu, len);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t base = 0; base < maxlen; base += G) {
        uint32_t mc = kPad;
        float mv = 0.f;
        if (base + gl < len) { mc = ld_stream(colIdxs + row.at(base + gl)); mv = ld_stream(vals + row.at(base + gl)); }
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const uint32_t c = __shfl_sync(0xFFFFFFFFu, mc, j, G);
            const float v = __shfl_sync(0xFFFFFFFFu, mv, j, G);
            if (valid && c != kPad)                  // kPad: past the end of the row (or ELL padding)
                fma4(acc, v, __ldg(reinterpret_cast<const float4 *>(B + (size_t)c * ldb + col)));
        }
    }
    if (valid && r < M) *reinterpret_cast<float4 *>(C + (size_t)r * ldc + col) = acc;
}

// =============================================================== variant 3 (staged)
// One CTA = NW compute warps + 1 producer warp.  The CTA owns R = NW*RW consecutive rows
// of A and a column tile of NT columns of B/C.  K is walked in chunks of KC rows of B;
// the producer streams chunk after chunk into a STAGES-deep shared-memory ring with TMA
// bulk copies (cp.async.bulk.shared.global, completion on an mbarrier); each compute warp
// consumes, for each of its RW rows, the non-zeros whose column falls inside the chunk
// (CSR columns are sorted, so that is a contiguous run found with a per-row cursor) and
// releases the stage.  Every B element fetched from L2 is re-used by ~R*density rows.
// Accumulation per C element is still strictly in CSR order (chunks are visited in
// ascending K), fp32 FMA.
namespace staged {

using pipe::smem_u32;
using pipe::mbar_init;
using pipe::mbar_expect_tx;
using pipe::mbar_arrive;
using pipe::bulk_g2s;
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) { pipe::mbar_wait<0>(bar, parity); }

template <int NT, int NW, int RW, int KC, int STAGES>
struct Cfg {
    static constexpr int kNT = NT, kNW = NW, kRW = RW, kKC = KC, kStages = STAGES;
    static constexpr int kRows = NW * RW;                     // row slots per CTA (upper bound)
    static constexpr int kThreads = (NW + 1) * 32;            // 15+1 warps = 512 threads (128 regs) or 31+1 = 1024 (64 regs)
    static constexpr int kU = NT / 128;                       // float4 per lane per row
    static constexpr size_t kStageBytes = (size_t)KC * NT * sizeof(float);
    static constexpr size_t kSmemBytes = kStageBytes * STAGES + 2 * STAGES * sizeof(uint64_t) + 128;
};

// SELL = false: CSR (rowPtrs = row pointers, entry idx of a row at colIdxs[idx], idx absolute).
// SELL = true : sliced ELL (rowPtrs = slicePtrs; entry j of row r at slicePtrs[r/32] + j*32 + r%32,
//               j in [0, W_slice); padding entries carry colIdx kPad and are never consumed).
// ---- thread-block-cluster helpers (CS > 1: the CTAs of a cluster share every B chunk through multicast TMA)
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t rank) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
                 "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
                 ::"r"(pipe::smem_u32(bar)), "r"(rank) : "memory");
}
// 1-D TMA bulk copy global -> the same shared-memory offset of every CTA in ctaMask; each destination's mbarrier gets the bytes
__device__ __forceinline__ void bulk_g2s_multicast(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint16_t ctaMask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(pipe::smem_u32(dst)), "l"(src), "r"(bytes), "r"(pipe::smem_u32(bar)), "h"(ctaMask) : "memory");
}

// CS = CTAs per cluster (1: no cluster).  CS > 1 (column tile == all of B's columns only): the CS CTAs of a cluster own
// consecutive row panels and walk the same chunks of B; every CTA fetches 1/CS of a chunk and multicasts it into all of them,
// so a chunk crosses the L2 -> SM fabric once per cluster instead of once per CTA (short row panels -- the 1/8 panel of a
// strong-scaled matrix -- are bound by exactly that traffic: 148 CTAs x 52 MB in 0.63 ms = 12 TB/s).  A ring stage is
// refilled when the consumers of ALL CTAs of the cluster have released it (arrives on every CTA's `empty` barrier).
template <class CFG, bool SELL, int CS = 1>
__global__ void __launch_bounds__(CFG::kThreads, 1)
csr_staged_kernel(const uint32_t *__restrict__ rowPtrs, const uint32_t *__restrict__ colIdxs,
                  const float *__restrict__ vals, uint32_t M, uint32_t K, uint32_t rpc,
                  const float *__restrict__ B, uint32_t N, size_t ldb,
                  float *__restrict__ C, size_t ldc) {
    // rpc = rows actually given to each CTA (<= kRows), chosen on the host so that the grid is a
    // whole number of waves of the SM count
    constexpr int NT = CFG::kNT, NW = CFG::kNW, RW = CFG::kRW, KC = CFG::kKC, STAGES = CFG::kStages;
    constexpr int U = CFG::kU;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *tiles = reinterpret_cast<float *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + CFG::kStageBytes * STAGES);
    uint64_t *empty = full + STAGES;

    // warp index through a shuffle: tells the compiler it is warp-uniform, so everything derived
    // from it (row numbers, cursors, trip counts) lives in uniform registers and branches on it
    // need no divergence handling around the shuffles below
    const uint32_t warp = __shfl_sync(0xFFFFFFFFu, threadIdx.x >> 5, 0), lane = lane_id();
    const uint32_t col0 = blockIdx.y * NT;                 // column tile (N % NT == 0 checked on host)
    const uint32_t row0 = blockIdx.x * rpc;
    const uint32_t rowEnd = min(M, row0 + rpc);
    const uint32_t nchunks = (K + KC - 1) / KC;

    // consumer warps that got no rows (the last panel, or rpc < the CTA's row slots) leave at once: spinning on the
    // barriers they would steal issue slots from the working warps (4096^2: 14 of 31 warps have rows)
    const uint32_t activeWarps = rowEnd > row0 ? min((uint32_t)NW, (rowEnd - row0 + RW - 1) / RW) : 0u;
    uint32_t crank = 0, clusterWarps = activeWarps;          // consumer warps of the whole cluster (arrivals per `empty` phase)
    if constexpr (CS > 1) {
        crank = cluster_ctarank();
        clusterWarps = 0;
        const uint32_t firstPanel = blockIdx.x - crank;
#pragma unroll
        for (int r = 0; r < CS; ++r) {
            const uint32_t a0 = (firstPanel + r) * rpc, a1 = min(M, a0 + rpc);
            clusterWarps += a1 > a0 ? min((uint32_t)NW, (a1 - a0 + RW - 1) / RW) : 0u;
        }
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, max(clusterWarps, 1u)); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if constexpr (CS > 1) {
        cluster_sync_all();                                  // every CTA's barriers exist before anybody multicasts / arrives remotely
    } else {
        if (activeWarps == 0 || (warp < NW && warp >= activeWarps)) return;
    }
    const bool clusterIdle = CS > 1 && clusterWarps == 0;    // no rows in the whole cluster: nothing to stream

    if (warp == NW) {
        // ------------------------------------------------------------ producer
        // lane 0 arms the barrier; when the tile rows are not contiguous in B (NT < ldb) all 32
        // lanes issue row copies in parallel (one thread issuing KC copies would be the bottleneck)
        if constexpr (CS > 1) {
            // cluster: this CTA fetches rows [crank * KC / CS, (crank + 1) * KC / CS) of every chunk and multicasts them
            constexpr uint32_t kSlice = KC / CS;
            for (uint32_t ch = 0; ch < nchunks && !clusterIdle; ++ch) {
                const uint32_t s = ch % STAGES, it = ch / STAGES;
                if (it > 0) mbar_wait(empty + s, (it - 1) & 1);
                const uint32_t k0 = ch * KC;
                const uint32_t rows = min((uint32_t)KC, K - k0);
                if (lane == 0) {
                    mbar_expect_tx(full + s, rows * NT * (uint32_t)sizeof(float));          // the whole chunk lands here
                    const uint32_t lo = crank * kSlice, hi = min(rows, lo + kSlice);
                    if (hi > lo)
                        bulk_g2s_multicast(tiles + (size_t)s * KC * NT + (size_t)lo * NT, B + (size_t)(k0 + lo) * ldb + col0,
                                           (hi - lo) * NT * (uint32_t)sizeof(float), full + s, (uint16_t)((1u << CS) - 1u));
                }
            }
        } else
        for (uint32_t ch = 0; ch < nchunks; ++ch) {
            const uint32_t s = ch % STAGES, it = ch / STAGES;
            if (it > 0) mbar_wait(empty + s, (it - 1) & 1);
            const uint32_t k0 = ch * KC;
            const uint32_t rows = min((uint32_t)KC, K - k0);
            float *dst = tiles + (size_t)s * KC * NT;
            if (lane == 0) mbar_expect_tx(full + s, rows * NT * (uint32_t)sizeof(float));
            __syncwarp();
            if ((size_t)NT == ldb) {       // tile rows are contiguous in B: one bulk copy
                if (lane == 0)
                    bulk_g2s(dst, B + (size_t)k0 * ldb + col0, rows * NT * (uint32_t)sizeof(float), full + s);
            } else {
                for (uint32_t i = lane; i < rows; i += 32)
                    bulk_g2s(dst + (size_t)i * NT, B + (size_t)(k0 + i) * ldb + col0,
                             NT * (uint32_t)sizeof(float), full + s);
            }
        }
        if constexpr (CS > 1) cluster_sync_all();            // nobody leaves while a peer may still write here
        return;
    }
    if constexpr (CS > 1) {
        if (warp >= activeWarps) {                           // consumer warp without rows: only the final cluster barrier
            cluster_sync_all();
            return;
        }
    }

    // ---------------------------------------------------------------- consumers
    // Per row: a 32-entry register window of (col, val) (lane l holds entry wbase + l; lanes past
    // the row end hold kPad), wj = entries of the window already consumed.
    uint32_t wbase[RW], wj[RW], end[RW], bcol[RW], off[RW];
    float bval[RW];
    float4 acc[RW][U];
    auto refill = [&](int i, uint32_t from) {
        wbase[i] = from;
        wj[i] = 0;
        bcol[i] = kPad;
        bval[i] = 0.f;
        if (from + lane < end[i]) {
            const size_t at = SELL ? (size_t)off[i] + (size_t)(from + lane) * 32u : (size_t)(from + lane);
            bcol[i] = ld_stream(colIdxs + at);
            bval[i] = ld_stream(vals + at);
        }
    };
#pragma unroll
    for (int i = 0; i < RW; ++i) {
        const uint32_t r = row0 + warp * RW + i;
        uint32_t p0 = 0;
        end[i] = 0;
        off[i] = 0;
        if (r < rowEnd) {
            if constexpr (SELL) {
                const uint32_t sb = __ldg(rowPtrs + (r >> 5));
                off[i] = sb + (r & 31u);
                end[i] = (__ldg(rowPtrs + (r >> 5) + 1) - sb) >> 5;      // slice width in slots
            } else {
                p0 = __ldg(rowPtrs + r);
                end[i] = __ldg(rowPtrs + r + 1);
            }
        }
        p0 = __shfl_sync(0xFFFFFFFFu, p0, 0);
        end[i] = __shfl_sync(0xFFFFFFFFu, end[i], 0);
        off[i] = __shfl_sync(0xFFFFFFFFu, off[i], 0);
        refill(i, p0);
#pragma unroll
        for (int u = 0; u < U; ++u) acc[i][u] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    for (uint32_t ch = 0; ch < nchunks; ++ch) {
        const uint32_t s = ch % STAGES, it = ch / STAGES;
        const uint32_t k0 = ch * KC, k1 = k0 + KC;
        // how many window entries of each row fall into this chunk: columns are ascending, so
        // the lanes with col < k1 form a prefix; drop the wj already-consumed ones.
        uint32_t nn[RW], maxn = 0;
#pragma unroll
        for (int i = 0; i < RW; ++i) {
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, bcol[i] < k1);
            nn[i] = wj[i] < 32 ? __popc(m >> wj[i]) : 0;
            maxn = max(maxn, nn[i]);
        }
        mbar_wait(full + s, it & 1);
        const float4 *tile = reinterpret_cast<const float4 *>(tiles + (size_t)s * KC * NT) + lane;
        // t-th entry of every row in flight together: RW independent shuffle -> LDS -> FMA chains
        for (uint32_t t = 0; t < maxn; ++t) {
#pragma unroll
            for (int i = 0; i < RW; ++i) {
                if (t < nn[i]) {
                    const uint32_t c = __shfl_sync(0xFFFFFFFFu, bcol[i], wj[i] + t);
                    const float v = __shfl_sync(0xFFFFFFFFu, bval[i], wj[i] + t);
                    const float4 *brow = tile + (size_t)(c - k0) * (NT / 4);
#pragma unroll
                    for (int u = 0; u < U; ++u) fma4(acc[i][u], v, brow[u * 32]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < RW; ++i) {
            wj[i] += nn[i];
            // window used up while the row may still have entries inside this chunk (rare: once
            // per 32 non-zeros of a row): refill and finish the chunk entry by entry
            if (wj[i] == 32 && wbase[i] + 32 < end[i]) {
                refill(i, wbase[i] + 32);
                while (true) {
                    if (wj[i] == 32) {
                        if (wbase[i] + 32 >= end[i]) break;
                        refill(i, wbase[i] + 32);
                    }
                    const uint32_t c = __shfl_sync(0xFFFFFFFFu, bcol[i], wj[i]);
                    if (c >= k1) break;                  // also ends at the row end (kPad)
                    const float v = __shfl_sync(0xFFFFFFFFu, bval[i], wj[i]);
                    const float4 *brow = tile + (size_t)(c - k0) * (NT / 4);
#pragma unroll
                    for (int u = 0; u < U; ++u) fma4(acc[i][u], v, brow[u * 32]);
                    ++wj[i];
                }
            }
        }
        __syncwarp();
        if constexpr (CS > 1) {
            if (lane < (uint32_t)CS) mbar_arrive_remote(empty + s, lane);      // lane r releases the stage in CTA r of the cluster
        } else {
            if (lane == 0) mbar_arrive(empty + s);
        }
    }

#pragma unroll
    for (int i = 0; i < RW; ++i) {
        const uint32_t r = row0 + warp * RW + i;
        if (r < rowEnd) {
            float4 *crow = reinterpret_cast<float4 *>(C + (size_t)r * ldc + col0) + lane;
#pragma unroll
            for (int u = 0; u < U; ++u) __stcs(crow + u * 32, acc[i][u]);
        }
    }
    if constexpr (CS > 1) cluster_sync_all();
}

// whole waves: (row panels x column tiles) is made a multiple of the SM count, rows per CTA =
// ceil(M / panels) <= rowSlots
struct GridPlan { uint32_t panels, rpc, waves; };
static GridPlan plan_grid(uint32_t M, uint32_t ytiles, uint32_t rowSlots) {
    const uint32_t sms = (uint32_t)sm_count();
    const uint32_t minPanels = (M + rowSlots - 1) / rowSlots;
    GridPlan g;
    g.waves = (minPanels * ytiles + sms - 1) / sms;
    g.panels = (g.waves * sms) / ytiles;
    if (g.panels < minPanels) g.panels = minPanels;
    g.rpc = (M + g.panels - 1) / g.panels;
    if (g.rpc > rowSlots) g.rpc = rowSlots;
    if (g.rpc == 0) g.rpc = 1;
    g.panels = (M + g.rpc - 1) / g.rpc;
    return g;
}

template <class CFG, bool SELL>
int launch(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals, uint32_t M, uint32_t K,
           const float *B, uint32_t N, size_t ldb, float *C, size_t ldc, cudaStream_t st) {
    auto kern = csr_staged_kernel<CFG, SELL>;
    CUSPMM_CUDA(set_smem_once(kern, CFG::kSmemBytes));
    const uint32_t ytiles = N / CFG::kNT;
    const GridPlan g = plan_grid(M, ytiles, CFG::kRows);
    dim3 grid(g.panels, ytiles);
    kern<<<grid, CFG::kThreads, CFG::kSmemBytes, st>>>(rowPtrs, colIdxs, vals, M, K, g.rpc, B, N, ldb, C, ldc);
    CUSPMM_LAUNCH_CHECK("csr_staged_kernel");
    return CUSPMM_OK;
}

// the staged kernel in clusters of CS CTAs (column tile == all N columns of B, ldb == NT)
template <class CFG, bool SELL, int CS>
int launch_cluster(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals, uint32_t M, uint32_t K,
                   const float *B, uint32_t N, size_t ldb, float *C, size_t ldc, cudaStream_t st) {
    auto kern = csr_staged_kernel<CFG, SELL, CS>;
    CUSPMM_CUDA(set_smem_once(kern, CFG::kSmemBytes));
    const GridPlan g = plan_grid(M, 1, CFG::kRows);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((g.panels + CS - 1) / CS * CS, 1, 1);
    cfg.blockDim = dim3(CFG::kThreads, 1, 1);
    cfg.dynamicSmemBytes = CFG::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CUSPMM_CUDA(cudaLaunchKernelEx(&cfg, kern, rowPtrs, colIdxs, vals, M, K, g.rpc, B, N, ldb, C, ldc));
    CUSPMM_LAUNCH_CHECK("csr_staged_kernel (cluster)");
    return CUSPMM_OK;
}

template <bool SELL>
int launch_by_N(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals, uint32_t M, uint32_t K, uint64_t nnz,
                const float *B, uint32_t N, size_t ldb, float *C, size_t ldc, cudaStream_t st) {
    static const int forceNT = getenv("CUSPMM_STAGED_NT") ? atoi(getenv("CUSPMM_STAGED_NT")) : 0;   // tuning hook
    // (31 warps x 2 rows on 256- or 128-column tiles, to give the CTAs of a 4096^2 matrix full row slots, measured slower than
    //  148 half-filled 512-column CTAs: 0.159 / 0.290 ms against 0.136 ms)
    if (forceNT == 128) return launch<Cfg<128, 31, 1, 64, 4>, SELL>(rowPtrs, colIdxs, vals, M, K, B, N, ldb, C, ldc, st);
    if (N % 512 == 0) {
        // Three shapes of the same kernel (measured in profiles/r01_staged_rw_tuning.txt):
        //   31 warps x 2 rows (1024 threads, 64 regs): 62-row panels, 8 warps per scheduler hide the
        //       shuffle -> LDS -> FMA latency; shared-memory pipe bound.  Best whenever a CTA gets >= 30 rows.
        //   31 warps x 1 row: 31-row panels for short matrices (M < 30 * SMs): twice the warps of 15 x 2 for the same rows
        //       (4096^2: 0.139 -> 0.130 ms);  15 warps x 2 rows: the same on sliced ELL.
        //   15 warps x 4 rows: the first version (60-row panels); kept behind the tuning hook only.
        static const int forceRW = getenv("CUSPMM_STAGED_RW") ? atoi(getenv("CUSPMM_STAGED_RW")) : 0;   // tuning hook
        const uint32_t ytiles = N / 512;
        int shape = forceRW;
        // (sliced ELL keeps 15 x 4: its window refills are strided 128-byte-apart loads, and 16 warps sharing a
        //  slice through a 30 KB L1 measured 6 % slower than 8 warps: 5.24 vs 4.92 ms on large_25605)
        if (shape == 0) shape = plan_grid(M, ytiles, 62).rpc >= 30 ? (SELL ? 4 : 31) : (SELL ? 2 : 1);
        // clusters with multicast TMA: tuning hook CUSPMM_STAGED_CLUSTER = 2 / 4, never selected.  Measured (bit-identical;
        // profiles/r02_cluster_multicast_probe.jsonl): the 1/8 row panel of large_25605 0.62 ms without, 0.92 / 1.84 / 1.95 ms in
        // clusters of 2 / 4 / 8, the full matrix 4.34 -> 4.91 / 6.62 ms: the release of a ring stage now needs a round trip
        // through every CTA of the cluster and a 3-deep ring cannot hide it, and nothing is gained on the other side because
        // short panels are bound by the WRITES of the B chunks into shared memory (410 K of 1.3 M wavefronts per CTA), which
        // multicast does not reduce -- not by L2 -> SM bandwidth.
        static const int forceCS = getenv("CUSPMM_STAGED_CLUSTER") ? atoi(getenv("CUSPMM_STAGED_CLUSTER")) : 0;
        if (N == 512 && ldb == 512 && forceCS > 1) {
            if (shape == 1) {
                if (forceCS == 2) return launch_cluster<Cfg<512, 31, 1, 32, 3>, SELL, 2>(rowPtrs, colIdxs, vals, M, K, B, N, ldb, C, ldc, st);
                if (forceCS == 4) return launch_cluster<Cfg<512, 31, 1, 32, 3>, SELL, 4>(rowPtrs, colIdxs, vals, M, K, B, N, ldb, C, ldc, st);
            } else if (shape == 31) {
                if (forceCS == 2) return launch_cluster<Cfg<512, 31, 2, 32, 3>, SELL, 2>(rowPtrs, colIdxs, vals, M, K, B, N, ldb, C, ldc, st);
                if (forceCS == 4) return launch_cluster<Cfg<512, 31, 2, 32, 3>, SELL, 4>(rowPtrs, colIdxs, vals, M, K, B, N, ldb, C, ldc, st);
            }
        }
        if (shape == 1) return launch<Cfg<512, 31, 1, 32, 3>, SELL>(rowPtrs, colIdxs, vals, M, K, B, N, ldb, C, ldc, st);
        if (shape == 31) return launch<Cfg<512, 31, 2, 32, 3>, SELL>(rowPtrs, colIdxs, vals, M, K, B, N, ldb, C, ldc, st);
        if (shape == 2) return launch<Cfg<512, 15, 2, 32, 3>, SELL>(rowPtrs, colIdxs, vals, M, K, B, N, ldb, C, ldc, st);
        return launch<Cfg<512, 15, 4, 32, 3>, SELL>(rowPtrs, colIdxs, vals, M, K, B, N, ldb, C, ldc, st);
    }
    // N = 128, 256, 384: 128-column tiles, one row per warp (31 rows per CTA keep >= 1 wave of CTAs for M >= 4 500), 4 ring stages;
    // chunks of 64 rows of B for sparse rows (fewer barrier rounds), 32 for dense ones (the 32-entry window lasts 2 chunks).
    // 4000^2, N = 128 (ms): d = 0.5: 0.202 (sub-warp kernel 0.271, cuSPARSE 0.371); d = 0.1: 0.050 (0.072, 0.098).
    // (Tried and dropped: rows on 8-lane groups, four rows per warp at a time -- 0.85 ms, too few warps per SM to hide the
    //  shuffle -> LDS -> FMA chain; 15 warps x 8 rows -- 1.26 ms, 33 CTAs.)
    const double density = (double)nnz / ((double)M * (double)K);
    if (density >= 0.3) return launch<Cfg<128, 31, 1, 32, 4>, SELL>(rowPtrs, colIdxs, vals, M, K, B, N, ldb, C, ldc, st);
    return launch<Cfg<128, 31, 1, 64, 4>, SELL>(rowPtrs, colIdxs, vals, M, K, B, N, ldb, C, ldc, st);
}

} // namespace staged

// variant 5: the staged design with the B chunks in tensor memory (spmm_csr_tmem.cu)
template <bool SELL>
int spmm_rows_tmem(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals, uint32_t M, uint32_t K, uint64_t nnz,
                   const float *B, uint32_t N, size_t ldb, float *C, size_t ldc, cudaStream_t st);
void tmem_planned_grid(uint32_t M, uint32_t N, uint64_t *ctas, uint32_t *rows_per_cta);
// variant 7: every B read from tensor memory, columns split over the TMEM lane quarters (spmm_csr_quad.cu)
template <bool SELL>
int spmm_rows_quad(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals, uint32_t M, uint32_t K, uint64_t nnz,
                   const float *B, uint32_t N, size_t ldb, float *C, size_t ldc, cudaStream_t st);
void quad_planned_grid(uint32_t M, uint32_t N, uint64_t *ctas, uint32_t *rows_per_cta);
// variant 8: A tiles made dense in shared memory, tcgen05.mma with a three-product tf32 / bf16 split (spmm_csr_tc.cu); CSR only
int spmm_csr_tc(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals, uint32_t M, uint32_t K, uint32_t nnz,
                const float *B, uint32_t N, size_t ldb, float *C, size_t ldc, bool sell, cudaStream_t st);
bool tc_kernel_wins(uint32_t M, uint32_t K, uint64_t nnz, uint32_t N, bool sell);   // is variant 8 expected to beat the fp32 kernels?
// variant 6: nnz split that cuts rows, ordered carry fix-up (spmm_csr_split.cu); needs workspace
size_t spmm_csr_split_workspace(uint32_t nnz, uint32_t N);
int spmm_csr_split(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals, uint32_t M, uint32_t K, uint32_t nnz,
                   const float *B, uint32_t N, size_t ldb, float *C, size_t ldc, void *ws, size_t wsBytes, cudaStream_t st);

// =============================================================== host dispatch
static bool vec_ok(const float *B, size_t ldb, const float *C, size_t ldc, uint32_t N) {
    return (N % 4 == 0) && (ldb % 4 == 0) && (ldc % 4 == 0) &&
           ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
}

static uint32_t pick_warps(uint32_t M) {
    // enough warps for ~4 resident CTAs of 8 warps on every SM, never more than one per row
    uint64_t want = (uint64_t)sm_count() * 8 * 8;
    return (uint32_t)(want < M ? want : (M ? M : 1));
}

// Kernel selection (DESIGN.md "kernel selection"; measured in profiles/r01_sweep.jsonl and
// profiles/r01_density_sweep_large_25605.txt):
//   - no 128-bit loads possible                       -> scalar row-split (4)
//   - N % 512 == 0, M >= 1024 and each staged B row is re-used often enough by a 60-row panel
//     (density * 60 >= 1.6, i.e. >= ~2.7 % dense)     -> staged (3): B tiles through shared memory
//     ... and >= 5.5 non-zeros per staged B row and CTA  -> dual-path staged (5): part of every B chunk in tensor memory
//   - N <= 512, or short rows (< 96 nnz/row)          -> sub-warp per row (2): 64-column tiles, many rows in flight
//   - otherwise (wide N, long rows)                   -> warp per row, nnz-balanced (1): A is re-read N/512 times only
// Tensor-core mode (cuspmm_set_csr_tensor_mode): 1 = the selector may choose variant 8, 0 = fp32 FMA kernels only.
// Unset: the environment variable CUSPMM_TENSOR (0 disables) decides, default on.
static std::atomic<int> g_tensorMode{-1};
static bool tensor_mode_on() {
    int m = g_tensorMode.load(std::memory_order_relaxed);
    if (m < 0) {
        const char *e = getenv("CUSPMM_TENSOR");
        m = (e && atoi(e) == 0) ? 0 : 1;
        g_tensorMode.store(m, std::memory_order_relaxed);
    }
    return m != 0;
}

int csr_select_variant(uint32_t M, uint32_t K, uint64_t nnz, uint32_t N, bool vec_ok, bool sell) {
    if (!vec_ok) return 4;
    // (on the sliced layout nnz counts slots, padding included: the kernel builds the padding too, as zeros it never stores)
    if (M && K && tensor_mode_on() && tc_kernel_wins(M, K, nnz, N, sell)) return 8;
    const double density = (double)nnz / ((double)M * (double)K);
    const double per_row = (double)nnz / (double)M;
    // N = 128: one 128-column tile, 31 rows per CTA (N = 256 / 384 would need row-wise TMA copies of 512 bytes: 25605^2,
    // N = 256: 10.2 ms against 3.4 ms for the sub-warp kernel)
    if (N == 128 && M >= 1024 && density * 31.0 >= 1.6) return 3;
    if (N % 512 == 0 && M >= 1024 && density * 60.0 >= 1.6) {
        // staged.  The dual-path kernel (5) is faster when a CTA's panel re-uses every staged B row often enough for the
        // chunk to outlast the TMEM copy round trip: rows per CTA x density >= 5.5 non-zeros per B row, with at least one full
        // wave of CTAs.  Measured (spmm_csr_tmem.cu, profiles/r01_sweep.jsonl): 25605^2 d=0.10 (5.8) +6 %, 20000^2 d=0.10 (4.6) -3 %,
        // d=0.12 (5.5) +4 %, d=0.3..0.5 +28..32 %; N=4096 (2-D tensor-map TMA): 11008x4096 d=0.10 (5.5) +6 %, d=0.5 +30 %
        uint64_t ctas = 0;
        uint32_t rpc = 0;
        tmem_planned_grid(M, N, &ctas, &rpc);
        const double reuse = (double)rpc * density;
        // (on the sliced-ELL layout the staged kernel is slower to begin with, so the switch comes earlier: 4.5; with less than a
        //  wave of CTAs the dual path still wins from ~15 non-zeros per B row (8 on sliced ELL): 4000^2 N=512 d=0.3 +5 %,
        //  d=0.5 +5..10 %)
        const double need = sell ? 4.5 : 5.4;      // (5.5 measured; a shade lower so that "10 % x 55 rows" does not fall on the wrong side of rounding)
        return ((ctas >= (uint64_t)sm_count() && reuse >= need) || reuse >= (sell ? 8.0 : 15.0)) ? 5 : 3;
    }
    // very short rows: the nnz-balanced warp-per-row kernel wins once a row spans several 64-column tiles of the sub-warp
    // kernel (20000^2, 14 nnz/row, N=512: 0.047 vs 0.055 ms; 4000^2, 40 nnz/row, N=2048: 0.082 vs 0.087 ms)
    if ((N >= 1024 && per_row < 96.0) || (N >= 512 && per_row < 24.0)) return 1;
    if (N <= 512 || per_row < 96.0) return 2;
    return 1;
}

// One dispatcher for both row layouts.  nnz = non-zeros (CSR) or slots (sliced ELL): only used
// for load balancing and kernel selection.
template <bool SELL>
static int rows_dispatch(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals,
                         uint32_t M, uint32_t K, uint32_t nnz, const float *B, uint32_t N, size_t ldb,
                         float *C, size_t ldc, int variant, cudaStream_t st) {
    CUSPMM_REQUIRE(variant >= 0 && variant <= CUSPMM_CSR_NUM_VARIANTS, "%s variant %d does not exist", SELL ? "ELL row" : "CSR", variant);
    if (variant == 6)
        return set_error(CUSPMM_ERR_WORKSPACE, "CSR variant 6 needs workspace: call cuspmm_spmm_csr_ws (cuspmm_spmm_csr_workspace bytes)");
    CUSPMM_REQUIRE(ldb >= N && ldc >= N, "ldb/ldc (%zu/%zu) must be >= N (%u)", ldb, ldc, N);
    if (M == 0 || N == 0) return CUSPMM_OK;
    CUSPMM_REQUIRE(rowPtrs && B && C && (nnz == 0 || (colIdxs && vals)), "null operand pointer");
    const bool vok = vec_ok(B, ldb, C, ldc, N);

    if (variant == 0) variant = csr_select_variant(M, K, nnz, N, vok, SELL);

    // debug guard for the precondition of the staged kernels (ascending columns inside a row): with CUSPMM_CHECK_SORTED set
    // every call that is about to run variant 3 / 5 / 7 / 8 on CSR first verifies it on the device (one pass over colIdxs, a
    // stream synchronisation) and fails with CUSPMM_ERR_INVALID instead of computing garbage
    static const bool checkSorted = getenv("CUSPMM_CHECK_SORTED") != nullptr;
    if (!SELL && checkSorted && (variant == 3 || variant == 5 || variant == 7 || variant == 8) && nnz) {
        uint32_t bad = 0;
        const int rc = cuspmm_csr_check_sorted(rowPtrs, colIdxs, M, K, &bad, st);
        if (rc) return rc;
        if (bad)
            return set_error(CUSPMM_ERR_INVALID, "%u rows have column indices that are not strictly ascending (or >= K): the staged "
                             "kernels (variants 3, 5, 7, 8, hence variant 0 on this shape) need sorted rows; use variant 1, 2, 4 or 6", bad);
    }

    // variants 1 and 2 keep their work decomposition but fall back to 32-bit loads when N, ldb/ldc or
    // the base pointers rule out 128-bit ones (N = 21 in data/small_210); variant 4 forces that path
    if ((variant == 1 || variant == 2) && !vok) variant = 4;

    const uint32_t warps = pick_warps(M);
    const uint64_t ipw = ((uint64_t)nnz + M + warps - 1) / warps;
    const uint32_t blocks = (warps + 7) / 8;

    switch (variant) {
    case 1: {
        // column tile per warp: 512 columns (4 float4 per lane) re-read A least; few rows (or a forced tile) -> narrower tiles,
        // so that a long row's entries are spread over several warps and more B rows are in flight per row (GL7d25: 0.082 /
        // 0.061 / 0.053 ms at 512 / 256 / 128 columns; 16 B rows in flight per lane instead of 8: slower, short rows fall into the tail loop)
        static const int forceU = getenv("CUSPMM_ROWSPLIT_U") ? atoi(getenv("CUSPMM_ROWSPLIT_U")) : 0;   // tuning hook
        int U = N > 256 ? 4 : (N > 128 ? 2 : 1);
        if (forceU == 1 || forceU == 2 || forceU == 4) U = forceU;
        const dim3 grid(blocks, (N + 128u * U - 1) / (128u * U));
        if (U == 4) csr_rowsplit_vec_kernel<4, 2, SELL><<<grid, 256, 0, st>>>(rowPtrs, colIdxs, vals, M, ipw, B, N, ldb, C, ldc);
        else if (U == 2) csr_rowsplit_vec_kernel<2, 4, SELL><<<grid, 256, 0, st>>>(rowPtrs, colIdxs, vals, M, ipw, B, N, ldb, C, ldc);
        else csr_rowsplit_vec_kernel<1, 8, SELL><<<grid, 256, 0, st>>>(rowPtrs, colIdxs, vals, M, ipw, B, N, ldb, C, ldc);
        CUSPMM_LAUNCH_CHECK("csr_rowsplit_vec_kernel");
        return CUSPMM_OK;
    }
    case 2: {
        const uint32_t G = N <= 16 ? 4 : (N <= 32 ? 8 : 16);
        const uint32_t rows_per_block = 256 / G;
        dim3 grid((M + rows_per_block - 1) / rows_per_block, (N + 4 * G - 1) / (4 * G));
        if (G == 4) csr_subwarp_vec_kernel<4, SELL><<<grid, 256, 0, st>>>(rowPtrs, colIdxs, vals, M, B, N, ldb, C, ldc);
        else if (G == 8) csr_subwarp_vec_kernel<8, SELL><<<grid, 256, 0, st>>>(rowPtrs, colIdxs, vals, M, B, N, ldb, C, ldc);
        else csr_subwarp_vec_kernel<16, SELL><<<grid, 256, 0, st>>>(rowPtrs, colIdxs, vals, M, B, N, ldb, C, ldc);
        CUSPMM_LAUNCH_CHECK("csr_subwarp_vec_kernel");
        return CUSPMM_OK;
    }
    case 3: {
        if (!(vok && N % 128 == 0))
            return set_error(CUSPMM_ERR_UNSUPPORTED, "staged kernel needs N %% 128 == 0 and aligned B/C (N=%u)", N);
        return staged::launch_by_N<SELL>(rowPtrs, colIdxs, vals, M, K, nnz, B, N, ldb, C, ldc, st);
    }
    case 5: {
        if (!(vok && N % 512 == 0))
            return set_error(CUSPMM_ERR_UNSUPPORTED, "TMEM-staged kernel needs N %% 512 == 0 and aligned B/C (N=%u)", N);
        return spmm_rows_tmem<SELL>(rowPtrs, colIdxs, vals, M, K, nnz, B, N, ldb, C, ldc, st);
    }
    case 7: {
        if (!(vok && N % 512 == 0))
            return set_error(CUSPMM_ERR_UNSUPPORTED, "all-TMEM kernel needs N %% 512 == 0 and aligned B/C (N=%u)", N);
        return spmm_rows_quad<SELL>(rowPtrs, colIdxs, vals, M, K, nnz, B, N, ldb, C, ldc, st);
    }
    case 8:
        return spmm_csr_tc(rowPtrs, colIdxs, vals, M, K, nnz, B, N, ldb, C, ldc, SELL, st);
    case 4: {
        dim3 grid(blocks, (N + 127) / 128);
        csr_rowsplit_scalar_kernel<4, SELL><<<grid, 256, 0, st>>>(rowPtrs, colIdxs, vals, M, ipw, B, N, ldb, C, ldc);
        CUSPMM_LAUNCH_CHECK("csr_rowsplit_scalar_kernel");
        return CUSPMM_OK;
    }
    }
    return set_error(CUSPMM_ERR_INVALID, "unreachable");
}

int spmm_csr_dispatch(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals,
                      uint32_t M, uint32_t K, uint32_t nnz, const float *B, uint32_t N, size_t ldb,
                      float *C, size_t ldc, int variant, cudaStream_t st) {
    return rows_dispatch<false>(rowPtrs, colIdxs, vals, M, K, nnz, B, N, ldb, C, ldc, variant, st);
}

// Few rows with enough non-zeros each: whole rows per warp leave the machine idle and the longest row is the critical path
// (GL7d25, 2798 rows of 2..422 non-zeros, N = 512: 0.072 ms; cut into equal nnz ranges: see profiles/r01_real_matrices.jsonl).
static bool prefer_split(uint32_t M, uint32_t K, uint64_t nnz, uint32_t N, bool vok) {
    if (!vok || M == 0) return false;
    const int sel = csr_select_variant(M, K, nnz, N, vok, false);
    if (sel == 3 || sel == 5 || sel == 8) return false;
    const uint64_t rowWarps = (uint64_t)M * ((N + 511) / 512);
    return rowWarps < (uint64_t)sm_count() * 32 && (double)nnz / M >= 16.0;
}

size_t spmm_csr_workspace_bytes(uint32_t M, uint32_t K, uint32_t nnz, uint32_t N, int variant) {
    if (variant == 6) return spmm_csr_split_workspace(nnz, N);
    if (variant == 0 && prefer_split(M, K, nnz, N, N % 4 == 0)) return spmm_csr_split_workspace(nnz, N);
    return 0;
}

int spmm_csr_dispatch_ws(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals,
                         uint32_t M, uint32_t K, uint32_t nnz, const float *B, uint32_t N, size_t ldb,
                         float *C, size_t ldc, int variant, void *ws, size_t wsBytes, cudaStream_t st) {
    CUSPMM_REQUIRE(ldb >= N && ldc >= N, "ldb/ldc (%zu/%zu) must be >= N (%u)", ldb, ldc, N);
    if (M == 0 || N == 0) return CUSPMM_OK;
    if (variant == 0 && ws && B && C && prefer_split(M, K, nnz, N, vec_ok(B, ldb, C, ldc, N)) &&
        wsBytes >= spmm_csr_split_workspace(nnz, N) && (reinterpret_cast<uintptr_t>(ws) & 15) == 0)
        variant = 6;
    if (variant == 6) {
        CUSPMM_REQUIRE(rowPtrs && B && C && (nnz == 0 || (colIdxs && vals)), "null operand pointer");
        return spmm_csr_split(rowPtrs, colIdxs, vals, M, K, nnz, B, N, ldb, C, ldc, ws, wsBytes, st);
    }
    return rows_dispatch<false>(rowPtrs, colIdxs, vals, M, K, nnz, B, N, ldb, C, ldc, variant, st);
}

// the same kernels on a sliced-ELL matrix (called from spmm_ell.cu); variant numbering as CSR
int spmm_sell_rows_dispatch(const uint32_t *slicePtrs, const uint32_t *colIdxs, const float *vals,
                            uint32_t M, uint32_t K, uint32_t slots, const float *B, uint32_t N, size_t ldb,
                            float *C, size_t ldc, int variant, cudaStream_t st) {
    return rows_dispatch<true>(slicePtrs, colIdxs, vals, M, K, slots, B, N, ldb, C, ldc, variant, st);
}

} // namespace cuspmm_b200

extern "C" int cuspmm_set_csr_tensor_mode(int mode) {
    const bool was = cuspmm_b200::tensor_mode_on();
    cuspmm_b200::g_tensorMode.store(mode != 0 ? 1 : 0, std::memory_order_relaxed);
    return was ? 1 : 0;
}

// what variant 0 resolves to for this shape on the current device (16-byte aligned operands assumed)
extern "C" int cuspmm_csr_selected_variant(uint32_t M, uint32_t K, uint32_t nnz, uint32_t N, int sliced_ell) {
    return cuspmm_b200::csr_select_variant(M, K, nnz, N, N % 4 == 0, sliced_ell != 0);
}

extern "C" size_t cuspmm_spmm_csr_workspace(uint32_t M, uint32_t K, uint32_t nnz, uint32_t N, int variant) {
    return cuspmm_b200::spmm_csr_workspace_bytes(M, K, nnz, N, variant);
}

extern "C" int cuspmm_spmm_csr_ws(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals,
                                  uint32_t M, uint32_t K, uint32_t nnz, const float *B, uint32_t N, size_t ldb,
                                  float *C, size_t ldc, int variant, void *workspace, size_t workspace_bytes, void *stream) {
    return cuspmm_b200::spmm_csr_dispatch_ws(rowPtrs, colIdxs, vals, M, K, nnz, B, N, ldb, C, ldc, variant, workspace,
                                             workspace_bytes, cuspmm_b200::as_stream(stream));
}

extern "C" int cuspmm_spmm_csr(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals,
                               uint32_t M, uint32_t K, uint32_t nnz, const float *B, uint32_t N, size_t ldb,
                               float *C, size_t ldc, int variant, void *stream) {
    return cuspmm_b200::spmm_csr_dispatch(rowPtrs, colIdxs, vals, M, K, nnz, B, N, ldb, C, ldc, variant,
                                          cuspmm_b200::as_stream(stream));
}
