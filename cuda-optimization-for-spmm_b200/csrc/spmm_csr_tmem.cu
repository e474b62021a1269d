// spmm_csr_tmem.cu -- CSR / sliced-ELL SpMM with TENSOR MEMORY as a second operand port (CSR variant 5, ELL variant 4).
//
// The staged kernel (variant 3, spmm_csr.cu) is bound by the shared-memory data pipe: every non-zero needs its own row
// of B, 4 x LDS.128 per warp for 512 columns = 16 wavefronts at 128 B/clk/SM (measured 17.6 clk per non-zero,
// profiles/r01_tmem_microbench.txt; ncu: pipe 94 % busy).  Blackwell has a second on-chip operand store with its own
// read port: tensor memory (256 KB/SM, 128 lanes x 512 columns).  One tcgen05.ld.32x32b.x16 hands every lane 16
// consecutive columns of its TMEM lane -- 2 KB per warp instruction, measured 3.2 clk/SM (641 B/clk/SM, 5x the
// shared-memory pipe) -- and its column address is a run-time value, so it can be the gathered B row of a non-zero.
// No tensor-core MMA is involved: the arithmetic stays fp32 FMA in CSR order (bit-identical to variants 1..4); TMEM
// is used purely as a faster operand cache ("CSR/COO/ELL are not treated as dense contractions").
//
// TMEM layout: a B row of 512 columns occupies 16 TMEM columns of EVERY lane quarter (a warp can only read the
// quarter (warp % 4)): lane l, column 16 kk + 4 u + j  =  B[k0 + kk][128 u + 4 l + j], so lane l ends up with the same
// four float4 it would have fetched with LDS.128.  It is filled from the shared-memory stage by
// tcgen05.cp.32x128b.warpx4, which replicates a 512-byte piece of a B row (32 lanes x 16 B) into all four quarters;
// the replication caps TMEM at 32 rows of B.
//
// History (git 0c1fac0, profiles/r01_tmem_microbench.txt): a kernel that served EVERY non-zero from TMEM (2 x 16-row
// TMEM ring behind a shared-memory staging ring, 22 consumer + 8 copy-issuing warps) was bit-exact but slower than
// variant 3 (5.9 vs 4.33 ms on large_25605): 32 rows are too few to decouple the consumer warps from the copy round
// trip.  The kernel below keeps variant 3's pipeline and uses TMEM for part of every chunk instead.
#include "tmem_common.cuh"

namespace cuspmm_b200 {
namespace tmemk {

// no-swizzle descriptor with SBO = 128: the 32 rows of a .32x128b copy are 512 contiguous bytes (LBO unused: one 16-byte column)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) { return make_desc(smem_addr, 128u, 128u); }
__device__ __forceinline__ void tmem_cp_32x128b_x4(uint32_t taddr, uint64_t desc) {
    asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}
// 16 columns of this warp's lane quarter -> 16 registers per lane; load and wait in ONE asm statement so that
// neither the compiler nor ptxas can move a use of the registers above the wait
__device__ __forceinline__ void tmem_ld16_wait(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}

// =====================================================================================================
// Dual operand path: variant 3's pipeline (3-stage shared-memory ring of 32-row chunks of B, TMA bulk
// copies, NW consumer warps x 2 rows) with tensor memory as a SECOND operand port: the first TR rows of
// every chunk are also copied into TMEM (replicated into the four lane quarters, see above).  A non-zero
// whose column falls into those rows fetches its B row with one tcgen05.ld.x16 and costs no shared-memory
// wavefronts; the others use 4 x LDS.128 as in variant 3.  The shared-memory data pipe -- the bound of
// variant 3 -- carries
//   (1 - TR/32) * 16 wavefronts per non-zero + the TMA writes + the copy engine's reads of TR rows
// instead of 16 per non-zero + the TMA writes.  TMEM: TS stages x TR rows x 16 columns <= 512 columns.
// Warps: 0..NW-1 consumers, NW..NW+NI-1 copy issuers; issuer 0 also drives the TMA ring.
// (Tried and dropped: loading the TMEM rows of a chunk with their own bulk copy and barrier so that the copies into TMEM start before
//  the rest of the chunk has landed -- the second barrier wait per chunk costs more than it saves: 4.19 vs 4.09 ms;
//  keeping the TMEM rows out of the ring -- a separate double buffer for them and a 4..5-deep
//  ring of the other rows -- measured 4.22 ms against 4.11 ms for this version on large_25605.)
template <int NW, int NI, int TR, int TS, int KC = 32, int STAGES = 3>
struct DualCfg {
    static constexpr int kNW = NW, kNI = NI, kTR = TR, kTS = TS, kRW = 2, kKC = KC, kStages = STAGES;
    static_assert(TR * TS * 16 <= 512 && TR >= 1 && TR <= KC && KC <= 32 && TS <= kStages, "TMEM has 512 columns; a TMEM stage lives as long as its ring stage");
    static constexpr int kRows = NW * kRW;
    static constexpr int kThreads = (NW + NI) * 32;
    static constexpr uint32_t kRowBytes = kNT * sizeof(float);                  // 2 KB
    static constexpr uint32_t kStageBytes = kKC * kRowBytes;                    // 64 KB
    static constexpr size_t kSmemBytes = (size_t)kStageBytes * kStages + 3 * kStages * sizeof(uint64_t) + 16 + 128;
    static_assert(kSmemBytes <= 232448, "more than 227 KB of shared memory");
};

template <class CFG, bool SELL>
__global__ void __launch_bounds__(CFG::kThreads, 1)
csr_dual_kernel(const uint32_t *__restrict__ rowPtrs, const uint32_t *__restrict__ colIdxs,
                const float *__restrict__ vals, uint32_t M, uint32_t K, uint32_t rpc,
                const float *__restrict__ B, size_t ldb, float *__restrict__ C, size_t ldc,
                const __grid_constant__ CUtensorMap tmapB, int useTmap) {
    constexpr int NW = CFG::kNW, NI = CFG::kNI, TR = CFG::kTR, TS = CFG::kTS, RW = CFG::kRW, KC = CFG::kKC;
    constexpr int STAGES = CFG::kStages, NT = kNT;
    constexpr uint32_t kStageBytes = CFG::kStageBytes, kRowBytes = CFG::kRowBytes;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *ring = smem_raw;
    uint64_t *tma_full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)kStageBytes * STAGES);   // TMA -> issuers
    uint64_t *full = tma_full + STAGES;       // copies of the chunk complete (implies the TMA data landed) -> consumers
    uint64_t *empty = full + STAGES;          // consumers -> TMA producer (ring stage) and issuers (TMEM stage)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(empty + STAGES);

    const uint32_t warp = __shfl_sync(0xFFFFFFFFu, threadIdx.x >> 5, 0), lane = lane_id();
    const uint32_t col0 = blockIdx.y * NT;
    const uint32_t row0 = blockIdx.x * rpc;
    const uint32_t rowEnd = min(M, row0 + rpc);
    // consumer warps without rows skip the pipeline (they would only spin on the barriers); with no rows at all nothing runs
    const uint32_t activeWarps = rowEnd > row0 ? min((uint32_t)NW, (rowEnd - row0 + RW - 1) / RW) : 0u;
    const uint32_t nchunks = activeWarps ? (K + KC - 1) / KC : 0u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(tma_full + s, 1); mbar_init(full + s, NI); mbar_init(empty + s, max(activeWarps, 1u)); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == NW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= NW) {
        // ------------------------------------------------------------ issuers (issuer 0 = TMA producer too)
        const uint32_t j = warp - NW;
        const bool rowwise = (size_t)NT != ldb && NI > 1 && !useTmap;
        // Two task streams, both driven by non-blocking barrier tests so that neither holds up the other:
        //   TMA load of chunk c (issuer 0):  B rows -> ring stage c % 3, once chunk c - 3 is consumed
        //   copies of chunk c (all issuers): first TR rows of the stage -> TMEM stage c % TS, once the TMA data
        //                                    has landed and chunk c - TS is consumed
        auto tma_ready = [&](uint32_t c) -> bool {
            return c < (uint32_t)STAGES || mbar_test(empty + c % STAGES, ((c - STAGES) / STAGES) & 1);
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
            if ((size_t)NT == ldb) {
                if (lane == 0) bulk_g2s(dst, B + (size_t)k0 * ldb + col0, rows * kRowBytes, tma_full + s);
            } else if (lane < rows) {
                bulk_g2s(dst + (size_t)lane * kRowBytes, B + (size_t)(k0 + lane) * ldb + col0, kRowBytes, tma_full + s);
            }
        };
        auto cp_ready = [&](uint32_t c) -> bool {
            if (!mbar_test(tma_full + c % STAGES, (c / STAGES) & 1)) return false;
            if (c < (uint32_t)TS) return true;
            const uint32_t cp = c - TS;
            return mbar_test(empty + cp % STAGES, (cp / STAGES) & 1);
        };
        auto cp_issue = [&](uint32_t c) {             // lane 0
            tc_fence_after();
            const uint64_t desc0 = make_desc(smem_u32(ring + (size_t)(c % STAGES) * kStageBytes));
            const uint32_t t0 = tmem_base + (c % TS) * (TR * 16);
            // piece p = 512 bytes: row p/4 of the chunk, columns 128 (p%4) ..; dealt round-robin to the copying issuers.
            // When the column tile is narrower than B the TMA load of a chunk is 32 row copies issued by issuer 0's lanes:
            // that warp then copies nothing (it still commits, so the barrier count stays NI).
            const uint32_t nCopy = rowwise ? NI - 1 : NI, jj = rowwise ? j - 1 : j;
            if (!(rowwise && j == 0))
                for (uint32_t p = jj; p < TR * 4; p += nCopy)
                    tmem_cp_32x128b_x4(t0 + p * 4, desc0 + (uint64_t)(p * 32));
            tc_commit(full + c % STAGES);
        };
        uint32_t cc = 0, tc = (j == 0) ? 0u : nchunks, idle = 0;
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
            if (did) idle = 0;
            else { __nanosleep(64); if (++idle > (1u << 28)) __trap(); }   // ~20 s without progress: a lost arrive must fail loudly, not hang the GPU
        }
    } else if (warp < activeWarps) {
        // ------------------------------------------------------------ consumers
        uint32_t wbase[RW], wj[RW], end[RW], bcol[RW], off[RW];
        float bval[RW];
        float2 acc[RW][8];                     // acc[i][2u + h] = columns 128u + 4 lane + 2h, +1 (packed fp32x2 FMAs)
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
                    end[i] = (__ldg(rowPtrs + (r >> 5) + 1) - sb) >> 5;
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
            for (int e = 0; e < 8; ++e) acc[i][e] = make_float2(0.f, 0.f);
        }
        const uint32_t tq = tmem_base + (((warp & 3u) * 32u) << 16);     // this warp's lane quarter

        // __ffma2_rn: packed fp32x2 FMA (FFMA2), per component the same IEEE fma as fmaf(), half the issue slots
        auto fma_row = [&](int i, float v, const uint32_t (&b)[16]) {
            const float2 v2 = make_float2(v, v);
#pragma unroll
            for (int e = 0; e < 8; ++e)
                acc[i][e] = __ffma2_rn(v2, make_float2(__uint_as_float(b[2 * e]), __uint_as_float(b[2 * e + 1])), acc[i][e]);
        };
        // cbase_v: TMEM address of column k0's row (stage base - 16 k0), kept in ONE vector register per chunk; the address
        // of a B row is formed with one LEA and moved to the uniform file (ptxas otherwise re-derives the base for every entry)
        auto from_tmem = [&](int i, uint32_t cbase_v, uint32_t c, float v) {
            uint32_t b[16];
            tmem_ld16_wait(cbase_v + (c << 4), b);
            fma_row(i, v, b);
        };
        // shared-space address of this lane's 16 bytes of row 0 / stage 0, made opaque so that it stays in ONE register
        // (ptxas otherwise re-derives it from SR_TID.X for every non-zero: 8 extra instructions per entry)
        uint32_t lane_sa = smem_u32(ring) + lane * 16u;
        asm volatile("" : "+r"(lane_sa));
        // chunk_sa = this lane's address in the stage, minus k0 rows: the B row of column c is at chunk_sa + c * 2048
        auto from_smem = [&](int i, uint32_t chunk_sa, uint32_t c, float v) {
            uint32_t b[16];
            const uint32_t a = chunk_sa + c * (uint32_t)(NT * sizeof(float));
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]) : "r"(a));
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+512];" : "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]) : "r"(a));
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+1024];" : "=r"(b[8]), "=r"(b[9]), "=r"(b[10]), "=r"(b[11]) : "r"(a));
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+1536];" : "=r"(b[12]), "=r"(b[13]), "=r"(b[14]), "=r"(b[15]) : "r"(a));
            fma_row(i, v, b);
        };

        for (uint32_t ch = 0; ch < nchunks; ++ch) {
            const uint32_t s = ch % STAGES, ts = ch % TS;
            const uint32_t k0 = ch * KC, k1 = k0 + KC, kT = k0 + TR;
            // window entries of each row inside this chunk (nn) and inside its TMEM-resident rows (nT): the
            // columns are ascending and the consumed entries are a prefix below k0
            uint32_t nn[RW], nT[RW], maxn = 0;
#pragma unroll
            for (int i = 0; i < RW; ++i) {
                nn[i] = __popc(__ballot_sync(0xFFFFFFFFu, bcol[i] < k1)) - wj[i];
                nT[i] = __popc(__ballot_sync(0xFFFFFFFFu, bcol[i] < kT)) - wj[i];
                maxn = max(maxn, nn[i]);
            }
            mbar_wait(full + s, (ch / STAGES) & 1);
            tc_fence_after();
            const uint32_t tile = lane_sa + s * kStageBytes - k0 * kRowBytes;       // stage row 0 = column k0 (mod 2^32)
            uint32_t cbase = tq + ts * (TR * 16) - (k0 << 4);
            asm volatile("" : "+r"(cbase));
            // t-th entry of both rows together while both have one (their shuffles are issued back to back, so the
            // second row's broadcast latency hides behind the first row's loads), then the longer row's tail
            static_assert(RW == 2, "the fused loop below is written for two rows per warp");
            const uint32_t minn = min(nn[0], nn[1]);
            uint32_t t = 0;
            for (; t < minn; ++t) {
                const uint32_t c0 = __shfl_sync(0xFFFFFFFFu, bcol[0], wj[0] + t);
                const float v0 = __shfl_sync(0xFFFFFFFFu, bval[0], wj[0] + t);
                const uint32_t c1 = __shfl_sync(0xFFFFFFFFu, bcol[1], wj[1] + t);
                const float v1 = __shfl_sync(0xFFFFFFFFu, bval[1], wj[1] + t);
                if (t < nT[0]) from_tmem(0, cbase, c0, v0);
                else from_smem(0, tile, c0, v0);
                if (t < nT[1]) from_tmem(1, cbase, c1, v1);
                else from_smem(1, tile, c1, v1);
            }
            for (; t < maxn; ++t) {
#pragma unroll
                for (int i = 0; i < RW; ++i) {
                    if (t < nn[i]) {
                        const uint32_t c = __shfl_sync(0xFFFFFFFFu, bcol[i], wj[i] + t);
                        const float v = __shfl_sync(0xFFFFFFFFu, bval[i], wj[i] + t);
                        if (t < nT[i]) from_tmem(i, cbase, c, v);
                        else from_smem(i, tile, c, v);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < RW; ++i) {
                wj[i] += nn[i];
                // window used up inside the chunk: the rest of the chunk entry by entry for this row.  (Going over the chunk
                // again with the two-row loop instead measured 1-3 % slower at every density: 4.20 / 9.26 / 14.56 ms.)
                if (wj[i] == 32 && wbase[i] + 32 < end[i]) {
                    refill(i, wbase[i] + 32);
                    while (true) {
                        if (wj[i] == 32) {
                            if (wbase[i] + 32 >= end[i]) break;
                            refill(i, wbase[i] + 32);
                        }
                        const uint32_t c = __shfl_sync(0xFFFFFFFFu, bcol[i], wj[i]);
                        if (c >= k1) break;
                        const float v = __shfl_sync(0xFFFFFFFFu, bval[i], wj[i]);
                        if (c < kT) from_tmem(i, cbase, c, v);
                        else from_smem(i, tile, c, v);
                        ++wj[i];
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        }

#pragma unroll
        for (int i = 0; i < RW; ++i) {
            const uint32_t r = row0 + warp * RW + i;
            if (r < rowEnd) {
                float4 *crow = reinterpret_cast<float4 *>(C + (size_t)r * ldc + col0) + lane;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float4 o;
                    o.x = acc[i][2 * u].x;
                    o.y = acc[i][2 * u].y;
                    o.z = acc[i][2 * u + 1].x;
                    o.w = acc[i][2 * u + 1].y;
                    __stcs(crow + u * 32, o);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == NW) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

template <class CFG, bool SELL>
static int launch_dual(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals, uint32_t M, uint32_t K,
                       const float *B, uint32_t N, size_t ldb, float *C, size_t ldc, cudaStream_t st) {
    auto kern = csr_dual_kernel<CFG, SELL>;
    CUSPMM_CUDA(set_smem_once(kern, CFG::kSmemBytes));
    const uint32_t ytiles = N / kNT;
    const GridPlan g = plan_grid(M, ytiles, CFG::kRows);
    dim3 grid(g.panels, ytiles);
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    static const bool noTmap = getenv("CUSPMM_NO_TMAP") != nullptr;        // tuning hook: row-wise bulk copies instead
    const int useTmap = ((size_t)kNT != ldb && !noTmap && make_tmap_B(&map, B, K, N, ldb, CFG::kKC)) ? 1 : 0;
    kern<<<grid, CFG::kThreads, CFG::kSmemBytes, st>>>(rowPtrs, colIdxs, vals, M, K, g.rpc, B, ldb, C, ldc, map, useTmap);
    CUSPMM_LAUNCH_CHECK("csr_dual_kernel");
    return CUSPMM_OK;
}

} // namespace tmemk

// variant 5 of the row kernels (CSR and sliced ELL): N % 512 == 0, 16-byte aligned B/C
// the grid the dual-path kernel (29 consumer warps x 2 rows) would use: CTAs and rows per CTA, for the selector
void tmem_planned_grid(uint32_t M, uint32_t N, uint64_t *ctas, uint32_t *rows_per_cta) {
    const uint32_t ytiles = N / tmemk::kNT;
    const tmemk::GridPlan g = tmemk::plan_grid(M, ytiles ? ytiles : 1, tmemk::DualCfg<29, 3, 10, 3>::kRows);
    *ctas = (uint64_t)g.panels * ytiles;
    *rows_per_cta = g.rpc;
}

template <bool SELL>
int spmm_rows_tmem(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals, uint32_t M, uint32_t K, uint64_t nnz,
                   const float *B, uint32_t N, size_t ldb, float *C, size_t ldc, cudaStream_t st) {
    if (N % tmemk::kNT != 0)
        return set_error(CUSPMM_ERR_UNSUPPORTED, "TMEM-staged kernel needs N %% 512 == 0 (N=%u)", N);
    static const int shape = getenv("CUSPMM_TMEM_SHAPE") ? atoi(getenv("CUSPMM_TMEM_SHAPE")) : 0;   // tuning hook
    // dual operand path: consumers, issuers, TMEM rows per chunk, TMEM stages  (large_25605: 10 rows x 3 stages 4.11 ms,
    // 8 x 3: 4.16, 6 x 3: 4.25, 4 x 3: 4.36, 1 x 3: 4.52, 16 x 2: 4.51; variant 3: 4.33)
    // (also measured on large_25605, d = 0.10: 24-row chunks x 4 ring stages 4.20 ms, 20-row chunks x 4 stages 4.71 ms, 30 consumer +
    //  2 issuer warps 4.39 ms -- the per-chunk bookkeeping outweighs the deeper ring / the larger TMEM share)
    // TMEM rows per chunk by density (large_25605, ms):   d     0.10   0.15   0.2    0.3    0.5
    //   the longer a chunk keeps the consumers busy,       v3     4.33   6.40   8.48   12.65  21.0
    //   the better a 2-deep TMEM ring of 16 rows hides     10x3   4.16   5.60   7.15   10.28  16.67
    //   its refill round trip                              16x2   4.51   5.89   7.10   9.68   15.11
    const double density = (double)nnz / ((double)M * (double)K);
    const bool sparse = shape == 1 || (shape != 2 && density < 0.2);
    static const int wide4 = getenv("CUSPMM_TMEM_WIDE4") ? atoi(getenv("CUSPMM_TMEM_WIDE4")) : 1;   // tuning hook
    // column tile narrower than B (2-D tensor-map TMA, or 32 row copies without it): 28 consumers + a fourth issuer warp
    // (11008x4096 d=0.1: 2.32 ms against 2.35 with 29 + 3; without the tensor map 2.80)
    if ((size_t)tmemk::kNT != ldb && wide4) {
        if (sparse) return tmemk::launch_dual<tmemk::DualCfg<28, 4, 10, 3>, SELL>(rowPtrs, colIdxs, vals, M, K, B, N, ldb, C, ldc, st);
        return tmemk::launch_dual<tmemk::DualCfg<28, 4, 16, 2>, SELL>(rowPtrs, colIdxs, vals, M, K, B, N, ldb, C, ldc, st);
    }
    if (sparse) return tmemk::launch_dual<tmemk::DualCfg<29, 3, 10, 3>, SELL>(rowPtrs, colIdxs, vals, M, K, B, N, ldb, C, ldc, st);
    return tmemk::launch_dual<tmemk::DualCfg<29, 3, 16, 2>, SELL>(rowPtrs, colIdxs, vals, M, K, B, N, ldb, C, ldc, st);
}
template int spmm_rows_tmem<false>(const uint32_t *, const uint32_t *, const float *, uint32_t, uint32_t, uint64_t,
                                   const float *, uint32_t, size_t, float *, size_t, cudaStream_t);
template int spmm_rows_tmem<true>(const uint32_t *, const uint32_t *, const float *, uint32_t, uint32_t, uint64_t,
                                  const float *, uint32_t, size_t, float *, size_t, cudaStream_t);

} // namespace cuspmm_b200
