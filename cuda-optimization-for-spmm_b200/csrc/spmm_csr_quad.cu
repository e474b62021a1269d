// spmm_csr_quad.cu -- CSR / sliced-ELL SpMM with EVERY B read served from tensor memory (CSR variant 7, ELL variant 5).
//
// The staged kernels (variants 3 / 5) give a warp whole 512-column rows of C, so a B row must be visible to all four TMEM
// lane quarters: replicated x4, TMEM holds 32 rows of B and most non-zeros still fetch B with 4 x LDS.128 (16 wavefronts
// of the 128 B/clk shared-memory pipe per non-zero -- the bound of those kernels).  This kernel splits the COLUMNS over
// the lane quarters instead: a warp owns RW = 8 rows x the 128 columns of its own quarter (warp % 4), four warps (a
// "group") cover the 512 columns of the same 8 rows.  Nothing is replicated:
//
//   TMEM[lane 32 q + l][column 4 (k mod 128) + j] = B[k][128 q + 4 l + j]          (a ring of 128 rows of B = 4 chunks)
//
// is exactly what tcgen05.cp.128x256b writes from a row-major 2 KB-per-row chunk in shared memory (two rows per copy,
// LBO = 2048, SBO = 128; layout verified in scripts/ubench/tmem_probe.cu), and a non-zero costs each of the four warps ONE
// tcgen05.ld.32x32b.x4 + two packed FFMA2 instead of LDS.128 + 4 FFMA.  Micro-benchmark (profiles/r01_tmem_microbench.txt,
// "TMEM.x4 only, 128 cols/warp"): 9.3-10 SM clocks per non-zero against 17.6 through shared memory.  Shared memory only
// carries the TMA writes and the copy engine's reads of each chunk (64 KB each way per 32 rows of B) plus the A entries.
//
// Pipeline per CTA (one per SM, 512 columns of TMEM):
//   issuer warp(s):  TMA bulk copy of chunk c (32 rows of B, 64 KB) -> smem ring stage c % 3            [tma_full]
//                    16 x tcgen05.cp.128x256b stage -> TMEM stage c % 4, then tcgen05.commit           [t_full, smem_free]
//   consumer warps:  wait t_full[c % 4]; for each of their 8 rows: the entries with k0 <= col < k1 in CSR order,
//                    acc[row] += val * TMEM row(col);  arrive t_empty[c % 4]
// Per C element the terms are added in ascending-column (= storage) order with one IEEE fma each: bit-identical to
// variants 1-5.  Column indices must ascend inside a row (cuspmm_b200.h).
//
// A side: no 32-entry register windows (8 rows x 2 registers would not fit): every visit of a chunk a warp loads, per
// row, the next L = 8 entries (lane = 8 * row + entry: two 32-lane loads cover the 8 rows), counts by ballot how many fall
// into the chunk (a prefix: ascending columns), parks (TMEM address, value) pairs in a 512-byte shared-memory scratch of
// its own and reads them back as warp-uniform LDS.128 broadcasts (two entries each).  The loads for the NEXT visit are
// issued before the current entries are processed.  A row with 8 or more entries inside one chunk makes the warp visit the
// chunk again (dense rows: every visit is full, the overhead is amortised).
//
// Measured (profiles/r02_ncu_quad_v1_v2_v3_summary.txt): bit-identical to variants 1-5 on every shape tried, shared-memory
// wavefronts 0.44 G instead of 0.90 G (dual) / 1.19 G (staged) on large_25605 -- but 4.75 ms against 4.13 ms for the dual-path
// kernel: four warps share every non-zero, each paying the per-row control flow for 2 FFMA2 of payload, so the kernel
// executes 3.6 G warp instructions (dual: 2.8 G) and is bound by issue slots (68 %) before the tcgen05.ld rate.  Two
// restructurings of the consumer loop (fixed batches of four entries on a sliding window; one visit per acquired stage)
// were slower still (5.2-5.9 ms; git history of this file).  The selector therefore never picks this variant; it stays
// callable (CSR variant 7 / ELL variant 5) and tested as the reference point for that design.
#include "tmem_common.cuh"

namespace cuspmm_b200 {
namespace quadk {

using namespace tmemk;

__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t desc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}

// NG groups of 4 consumer warps (8 rows each), NI issuer warps (issuer 0 also drives the TMA ring), HB = TMEM loads in
// flight per wait (2 or 4)
template <int NG, int NI, int HB = 4, int STAGES = 3>
struct QuadCfg {
    static constexpr int kNG = NG, kNI = NI, kHB = HB, kStages = STAGES;
    static constexpr int kRW = 8, kL = 8, kKC = 32, kTS = 4;
    static constexpr int kNC = NG * 4;                      // consumer warps
    static constexpr int kRows = NG * kRW;
    static constexpr int kThreads = (kNC + NI) * 32;
    static constexpr int kSlots = kRW * kL / 32;            // 32-lane loads per visit
    static constexpr uint32_t kRowBytes = kNT * sizeof(float);                  // 2 KB
    static constexpr uint32_t kStageBytes = kKC * kRowBytes;                    // 64 KB
    static constexpr uint32_t kScratchBytes = kRW * kL * 8;                     // per consumer warp
    static constexpr size_t kSmemBytes = (size_t)kStageBytes * STAGES + (size_t)kScratchBytes * kNC +
                                         (2 * STAGES + 2 * kTS) * sizeof(uint64_t) + 16 + 128;
    static_assert(kThreads <= 1024 && kSmemBytes <= 232448, "CTA limits");
    static_assert(HB == 2 || HB == 4, "HB");
};

// CNT (1..4) consecutive entries of one row: (TMEM address, value) pairs at shared address sp (warp-uniform LDS
// broadcasts), CNT tcgen05.ld.x4 in flight, one wait, then the FMAs in entry order.  Loads and wait are ONE asm statement per
// CNT (straight-line code, no predicated tcgen05.ld: ptxas turns those into branches around stack round trips), so no use
// of the loaded registers can move above the wait.
__device__ __forceinline__ void fma_entry(float2 (&acc)[2], float v, uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3) {
    const float2 v2 = make_float2(v, v);
    acc[0] = __ffma2_rn(v2, make_float2(__uint_as_float(b0), __uint_as_float(b1)), acc[0]);
    acc[1] = __ffma2_rn(v2, make_float2(__uint_as_float(b2), __uint_as_float(b3)), acc[1]);
}
template <int CNT>
__device__ __forceinline__ void row_batch(float2 (&acc)[2], uint32_t sp);
template <>
__device__ __forceinline__ void row_batch<1>(float2 (&acc)[2], uint32_t sp) {
    uint32_t b[4];
    float v0;
    asm volatile(
        "{\n\t.reg .b32 a0;\n\t"
        "ld.shared.v2.b32 {a0, %4}, [%5];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [a0];\n\t"
        "tcgen05.wait::ld.sync.aligned;\n\t}"
        : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=f"(v0)
        : "r"(sp) : "memory");
    fma_entry(acc, v0, b[0], b[1], b[2], b[3]);
}
template <>
__device__ __forceinline__ void row_batch<2>(float2 (&acc)[2], uint32_t sp) {
    uint32_t b[8];
    float v0, v1;
    asm volatile(
        "{\n\t.reg .b32 a0, a1;\n\t"
        "ld.shared.v4.b32 {a0, %8, a1, %9}, [%10];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [a0];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%4, %5, %6, %7}, [a1];\n\t"
        "tcgen05.wait::ld.sync.aligned;\n\t}"
        : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]), "=f"(v0), "=f"(v1)
        : "r"(sp) : "memory");
    fma_entry(acc, v0, b[0], b[1], b[2], b[3]);
    fma_entry(acc, v1, b[4], b[5], b[6], b[7]);
}
template <>
__device__ __forceinline__ void row_batch<3>(float2 (&acc)[2], uint32_t sp) {
    uint32_t b[12];
    float v0, v1, v2;
    asm volatile(
        "{\n\t.reg .b32 a0, a1, a2;\n\t"
        "ld.shared.v4.b32 {a0, %12, a1, %13}, [%15];\n\t"
        "ld.shared.v2.b32 {a2, %14}, [%15+16];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [a0];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%4, %5, %6, %7}, [a1];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%8, %9, %10, %11}, [a2];\n\t"
        "tcgen05.wait::ld.sync.aligned;\n\t}"
        : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]),
          "=r"(b[8]), "=r"(b[9]), "=r"(b[10]), "=r"(b[11]), "=f"(v0), "=f"(v1), "=f"(v2)
        : "r"(sp) : "memory");
    fma_entry(acc, v0, b[0], b[1], b[2], b[3]);
    fma_entry(acc, v1, b[4], b[5], b[6], b[7]);
    fma_entry(acc, v2, b[8], b[9], b[10], b[11]);
}
template <>
__device__ __forceinline__ void row_batch<4>(float2 (&acc)[2], uint32_t sp) {
    uint32_t b[16];
    float v0, v1, v2, v3;
    asm volatile(
        "{\n\t.reg .b32 a0, a1, a2, a3;\n\t"
        "ld.shared.v4.b32 {a0, %16, a1, %17}, [%20];\n\t"
        "ld.shared.v4.b32 {a2, %18, a3, %19}, [%20+16];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [a0];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%4, %5, %6, %7}, [a1];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%8, %9, %10, %11}, [a2];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%12, %13, %14, %15}, [a3];\n\t"
        "tcgen05.wait::ld.sync.aligned;\n\t}"
        : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]),
          "=r"(b[8]), "=r"(b[9]), "=r"(b[10]), "=r"(b[11]), "=r"(b[12]), "=r"(b[13]), "=r"(b[14]), "=r"(b[15]),
          "=f"(v0), "=f"(v1), "=f"(v2), "=f"(v3)
        : "r"(sp) : "memory");
    fma_entry(acc, v0, b[0], b[1], b[2], b[3]);
    fma_entry(acc, v1, b[4], b[5], b[6], b[7]);
    fma_entry(acc, v2, b[8], b[9], b[10], b[11]);
    fma_entry(acc, v3, b[12], b[13], b[14], b[15]);
}
// n (1..HB, warp-uniform) entries at sp
template <int HB>
__device__ __forceinline__ void row_entries(float2 (&acc)[2], uint32_t sp, uint32_t n) {
    if constexpr (HB == 4) {
        if (n >= 4) row_batch<4>(acc, sp);
        else if (n == 3) row_batch<3>(acc, sp);
        else if (n == 2) row_batch<2>(acc, sp);
        else row_batch<1>(acc, sp);
    } else {
        if (n >= 2) row_batch<2>(acc, sp);
        else row_batch<1>(acc, sp);
    }
}

template <class CFG, bool SELL>
__global__ void __launch_bounds__(CFG::kThreads, 1)
csr_quad_kernel(const uint32_t *__restrict__ rowPtrs, const uint32_t *__restrict__ colIdxs,
                const float *__restrict__ vals, uint32_t M, uint32_t K, uint32_t rpc,
                const float *__restrict__ B, size_t ldb, float *__restrict__ C, size_t ldc,
                const __grid_constant__ CUtensorMap tmapB, int useTmap) {
    constexpr int NC = CFG::kNC, NI = CFG::kNI, RW = CFG::kRW, L = CFG::kL, KC = CFG::kKC, TS = CFG::kTS;
    constexpr int STAGES = CFG::kStages, NS = CFG::kSlots, HB = CFG::kHB;
    constexpr uint32_t kStageBytes = CFG::kStageBytes, kRowBytes = CFG::kRowBytes;
    constexpr uint32_t STEP = SELL ? 32u : 1u;              // distance between consecutive entries of a row
    static_assert(RW == 8 && L == 8 && NS == 2, "the lane <-> (row, entry) mapping below is written for 8 x 8");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *ring = smem_raw;
    unsigned char *scratch = smem_raw + (size_t)kStageBytes * STAGES;
    uint64_t *tma_full = reinterpret_cast<uint64_t *>(scratch + (size_t)CFG::kScratchBytes * NC);   // TMA -> issuers
    uint64_t *smem_free = tma_full + STAGES;   // copies out of the ring stage complete -> TMA producer
    uint64_t *t_full = smem_free + STAGES;     // copies into the TMEM stage complete -> consumers
    uint64_t *t_empty = t_full + TS;           // consumers -> issuers
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(t_empty + TS);

    const uint32_t warp = __shfl_sync(0xFFFFFFFFu, threadIdx.x >> 5, 0), lane = lane_id();
    const uint32_t col0 = blockIdx.y * kNT;
    const uint32_t row0 = blockIdx.x * rpc;
    const uint32_t rowEnd = min(M, row0 + rpc);
    // groups without rows skip the pipeline (they would only spin on the barriers); with no rows at all nothing runs
    const uint32_t activeGroups = rowEnd > row0 ? min((uint32_t)CFG::kNG, (rowEnd - row0 + RW - 1) / RW) : 0u;
    const uint32_t activeWarps = activeGroups * 4u;
    const uint32_t nchunks = activeWarps ? (K + KC - 1) / KC : 0u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(tma_full + s, 1); mbar_init(smem_free + s, NI); }
        for (int t = 0; t < TS; ++t) { mbar_init(t_full + t, NI); mbar_init(t_empty + t, max(activeWarps, 1u)); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == NC) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= NC) {
        // ------------------------------------------------------------ issuers (issuer 0 = TMA producer too)
        const uint32_t j = warp - NC;
        // Two task streams driven by non-blocking barrier tests so that neither holds up the other:
        //   TMA load of chunk c (issuer 0):  B rows -> ring stage c % STAGES, once the copies of chunk c - STAGES are complete
        //   copies of chunk c (all issuers): ring stage -> TMEM stage c % TS, once the TMA data has landed and chunk
        //                                    c - TS has been consumed by every consumer warp
        auto tma_ready = [&](uint32_t c) -> bool {
            return c < (uint32_t)STAGES || mbar_test(smem_free + c % STAGES, ((c - STAGES) / STAGES) & 1);
        };
        auto tma_issue = [&](uint32_t c) {            // whole warp
            const uint32_t s = c % STAGES;
            const uint32_t k0 = c * KC;
            const uint32_t rows = min((uint32_t)KC, K - k0);
            unsigned char *dst = ring + (size_t)s * kStageBytes;
            if (useTmap) {                            // column tile narrower than B: ONE 2-D tensor copy (32 rows x 2 KB box)
                if (lane == 0) {
                    mbar_expect_tx(tma_full + s, kStageBytes);
                    tma_box_2d(dst, &tmapB, col0 / 2, k0, tma_full + s);      // the map counts 8-byte elements
                }
                return;
            }
            if (lane == 0) mbar_expect_tx(tma_full + s, rows * kRowBytes);
            __syncwarp();
            if ((size_t)kNT == ldb) {
                if (lane == 0) bulk_g2s(dst, B + (size_t)k0 * ldb + col0, rows * kRowBytes, tma_full + s);
            } else if (lane < rows) {
                bulk_g2s(dst + (size_t)lane * kRowBytes, B + (size_t)(k0 + lane) * ldb + col0, kRowBytes, tma_full + s);
            }
        };
        auto cp_ready = [&](uint32_t c) -> bool {
            if (!mbar_test(tma_full + c % STAGES, (c / STAGES) & 1)) return false;
            return c < (uint32_t)TS || mbar_test(t_empty + c % TS, ((c - TS) / TS) & 1);
        };
        auto cp_issue = [&](uint32_t c) {             // lane 0
            tc_fence_after();
            // rows 2p, 2p + 1 of the chunk -> TMEM columns 8p .. 8p + 7 of the stage, all 128 lanes (a row = 128 lanes x 16 B)
            const uint64_t desc0 = make_desc(smem_u32(ring + (size_t)(c % STAGES) * kStageBytes), kRowBytes, 128u);
            const uint32_t t0 = tmem_base + (c % TS) * (KC * 4);
            for (uint32_t p = j; p < KC / 2; p += NI)
                tmem_cp_128x256b(t0 + p * 8, desc0 + (uint64_t)(p * (2 * kRowBytes / 16)));
            tc_commit(t_full + c % TS);
            tc_commit(smem_free + c % STAGES);
        };
        uint32_t cc = 0, tc = (j == 0) ? 0u : nchunks, polls = 0;
        uint64_t idle0 = 0;
        while (cc < nchunks || tc < nchunks) {
            bool did = false;
            if (tc < nchunks) {                       // warp-uniform: the test result is broadcast from lane 0
                const bool r = __shfl_sync(0xFFFFFFFFu, (lane == 0 && tma_ready(tc)) ? 1 : 0, 0) != 0;
                if (r) { tma_issue(tc); ++tc; did = true; }
            }
            if (cc < nchunks) {
                const bool r = __shfl_sync(0xFFFFFFFFu, (lane == 0 && cp_ready(cc)) ? 1 : 0, 0) != 0;
                if (r) { if (lane == 0) cp_issue(cc); ++cc; did = true; }
            }
            if (did) { polls = 0; idle0 = 0; }
            else {
                __nanosleep(32);
                if ((++polls & 4095u) == 0) {         // no progress for 20 s of wall clock: a lost arrive must fail loudly
                    const uint64_t now = pipe::global_ns();
                    if (idle0 == 0) idle0 = now;
                    else if (now - idle0 > pipe::kWaitLimitNs) __trap();
                }
            }
        }
    } else if (warp < activeWarps) {
        // ------------------------------------------------------------ consumers
        const uint32_t q = warp & 3u, g = warp >> 2;
        const uint32_t rbase = row0 + g * RW;
        const uint32_t tq = tmem_base + ((q * 32u) << 16);                // this warp's lane quarter
        uint2 *sc = reinterpret_cast<uint2 *>(scratch + (size_t)warp * CFG::kScratchBytes);
        uint32_t sc_sa = smem_u32(sc);
        asm volatile("" : "+r"(sc_sa));
        // slot s, lane l: entry (l & 7) of the look-ahead of row 4 s + (l >> 3)
        uint32_t cur[NS], endp[NS], ecol[NS];
        float eval[NS];
        float2 acc[RW][2];                     // acc[i] = columns 128 q + 4 lane .. + 3 of row i
#pragma unroll
        for (int i = 0; i < RW; ++i) acc[i][0] = acc[i][1] = make_float2(0.f, 0.f);
        auto fetch = [&](int s) {
            ecol[s] = kPad;
            eval[s] = 0.f;
            if (cur[s] < endp[s]) {
                ecol[s] = __ldg(colIdxs + cur[s]);
                eval[s] = __ldg(vals + cur[s]);
            }
        };
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const uint32_t r = rbase + s * 4 + (lane >> 3);
            cur[s] = endp[s] = 0;
            if (r < rowEnd) {
                if constexpr (SELL) {
                    const uint32_t sb = __ldg(rowPtrs + (r >> 5));
                    const uint32_t w = (__ldg(rowPtrs + (r >> 5) + 1) - sb) >> 5;      // slice width in slots
                    cur[s] = sb + (r & 31u) + (lane & 7u) * 32u;
                    endp[s] = sb + (r & 31u) + w * 32u;
                } else {
                    cur[s] = __ldg(rowPtrs + r) + (lane & 7u);
                    endp[s] = __ldg(rowPtrs + r + 1);
                }
            }
            fetch(s);
        }

        uint32_t ch = 0;
        bool fresh = true;
        while (ch < nchunks) {
            const uint32_t k1 = (ch + 1) * KC;
            // entries of the look-ahead inside this chunk: a prefix of every row's 8 lanes (ascending columns)
            uint32_t m[NS];
            bool more = false;
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                m[s] = __ballot_sync(0xFFFFFFFFu, ecol[s] < k1);
                sc[s * 32 + lane] = make_uint2(tq + ((ecol[s] & 127u) << 2), __float_as_uint(eval[s]));
#pragma unroll
                for (int b = 0; b < 4; ++b) more |= ((m[s] >> (8 * b)) & 0xFFu) == 0xFFu;     // all 8 inside: maybe more behind them
            }
            __syncwarp();
            // advance and issue the loads of the next visit now: their latency hides behind this visit's entries
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                cur[s] += (uint32_t)__popc((m[s] >> (lane & 24u)) & 0xFFu) * STEP;
                fetch(s);
            }
            if (fresh) {
                mbar_wait(t_full + ch % TS, (ch / TS) & 1);
                tc_fence_after();
            }
#pragma unroll
            for (int i = 0; i < RW; ++i) {
                const uint32_t n = (uint32_t)__popc((m[i >> 2] >> (8 * (i & 3))) & 0xFFu);     // warp-uniform
                const uint32_t sp = sc_sa + i * (L * 8);
                if (n > 0) {
                    row_entries<HB>(acc[i], sp, n);
                    if (n > HB) {
                        row_entries<HB>(acc[i], sp + HB * 8, n - HB);
                        if (HB == 2 && n > 4) {
                            row_entries<HB>(acc[i], sp + 32, n - 4);
                            if (n > 6) row_entries<HB>(acc[i], sp + 48, n - 6);
                        }
                    }
                }
            }
            __syncwarp();                              // every lane is done with the scratch before the next visit overwrites it
            if (!more) {
                tc_fence_before();
                if (lane == 0) mbar_arrive(t_empty + ch % TS);
                ++ch;
                fresh = true;
            } else {
                fresh = false;
            }
        }

#pragma unroll
        for (int i = 0; i < RW; ++i) {
            const uint32_t r = rbase + i;
            if (r < rowEnd) {
                float4 o;
                o.x = acc[i][0].x; o.y = acc[i][0].y; o.z = acc[i][1].x; o.w = acc[i][1].y;
                __stcs(reinterpret_cast<float4 *>(C + (size_t)r * ldc + col0 + q * 128u) + lane, o);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == NC) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

template <class CFG, bool SELL>
static int launch_quad(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals, uint32_t M, uint32_t K,
                       const float *B, uint32_t N, size_t ldb, float *C, size_t ldc, cudaStream_t st) {
    auto kern = csr_quad_kernel<CFG, SELL>;
    CUSPMM_CUDA(set_smem_once(kern, CFG::kSmemBytes));
    const uint32_t ytiles = N / kNT;
    const GridPlan g = plan_grid(M, ytiles, CFG::kRows);
    dim3 grid(g.panels, ytiles);
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    static const bool noTmap = getenv("CUSPMM_NO_TMAP") != nullptr;        // tuning hook: row-wise bulk copies instead
    const int useTmap = ((size_t)kNT != ldb && !noTmap && make_tmap_B(&map, B, K, N, ldb, CFG::kKC)) ? 1 : 0;
    kern<<<grid, CFG::kThreads, CFG::kSmemBytes, st>>>(rowPtrs, colIdxs, vals, M, K, g.rpc, B, ldb, C, ldc, map, useTmap);
    CUSPMM_LAUNCH_CHECK("csr_quad_kernel");
    return CUSPMM_OK;
}

} // namespace quadk

// the grid the quad kernel would use: CTAs and rows per CTA, for the selector
void quad_planned_grid(uint32_t M, uint32_t N, uint64_t *ctas, uint32_t *rows_per_cta) {
    const uint32_t ytiles = N / tmemk::kNT;
    const tmemk::GridPlan g = tmemk::plan_grid(M, ytiles ? ytiles : 1, quadk::QuadCfg<7, 1>::kRows);
    *ctas = (uint64_t)g.panels * ytiles;
    *rows_per_cta = g.rpc;
}

// variant 7 of the row kernels (CSR and sliced ELL): N % 512 == 0, 16-byte aligned B/C
template <bool SELL>
int spmm_rows_quad(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals, uint32_t M, uint32_t K, uint64_t nnz,
                   const float *B, uint32_t N, size_t ldb, float *C, size_t ldc, cudaStream_t st) {
    if (N % tmemk::kNT != 0)
        return set_error(CUSPMM_ERR_UNSUPPORTED, "the all-TMEM kernel needs N %% 512 == 0 (N=%u)", N);
    (void)nnz;
    // tuning hook: CUSPMM_QUAD_SHAPE = <groups><issuers><loads per wait>, e.g. 714 (default), 724, 712, 624
    static const int shape = getenv("CUSPMM_QUAD_SHAPE") ? atoi(getenv("CUSPMM_QUAD_SHAPE")) : 714;
    switch (shape) {
    case 724: return quadk::launch_quad<quadk::QuadCfg<7, 2, 4>, SELL>(rowPtrs, colIdxs, vals, M, K, B, N, ldb, C, ldc, st);
    case 712: return quadk::launch_quad<quadk::QuadCfg<7, 1, 2>, SELL>(rowPtrs, colIdxs, vals, M, K, B, N, ldb, C, ldc, st);
    case 722: return quadk::launch_quad<quadk::QuadCfg<7, 2, 2>, SELL>(rowPtrs, colIdxs, vals, M, K, B, N, ldb, C, ldc, st);
    case 624: return quadk::launch_quad<quadk::QuadCfg<6, 2, 4>, SELL>(rowPtrs, colIdxs, vals, M, K, B, N, ldb, C, ldc, st);
    case 614: return quadk::launch_quad<quadk::QuadCfg<6, 1, 4>, SELL>(rowPtrs, colIdxs, vals, M, K, B, N, ldb, C, ldc, st);
    default: return quadk::launch_quad<quadk::QuadCfg<7, 1, 4>, SELL>(rowPtrs, colIdxs, vals, M, K, B, N, ldb, C, ldc, st);
    }
}
template int spmm_rows_quad<false>(const uint32_t *, const uint32_t *, const float *, uint32_t, uint32_t, uint64_t,
                                   const float *, uint32_t, size_t, float *, size_t, cudaStream_t);
template int spmm_rows_quad<true>(const uint32_t *, const uint32_t *, const float *, uint32_t, uint32_t, uint64_t,
                                  const float *, uint32_t, size_t, float *, size_t, cudaStream_t);

} // namespace cuspmm_b200
