// spmm_bsr_tc.cu -- tensor-core BSR SpMM for sm_100a: tcgen05.mma, accumulators in TMEM,
// operands fed by TMA bulk copies, fp32 accumulate (north_star: "BSR: the only path treated
// as a dense contraction").  Replaces spmmBSRK1 (src/spmm/bsr/spmm_bsr_k1.cu) for 16x16 and
// 32x32 blocks stored as bf16 / fp16.
//
// Mapping: UMMA-M is 128, far larger than a block, so the kernel
// computes C^T tiles:   D[128 n  x  bs rows] += B^T[128 n  x  16 k] * A_blk^T[16 k  x  bs rows]
//   MMA "A" operand = 128 columns of B for 16 consecutive k            (M = 128, K = 16)
//   MMA "B" operand = 16 k-columns of the sparse block, all bs rows    (N = bs,  K = 16)
//   D (TMEM)        = lane n, column i  ==  C[blockRow*bs + i][n]      (fp32)
// One CTA owns one block row and up to 512 columns (4 D tiles = 4*bs TMEM columns) and walks
// the block row's blocks through a 4-stage smem ring.
//
// Operand layout.  Both operands are K-major with NO swizzle, i.e. UMMA "core matrices" of
// 8 rows x 16 bytes stored contiguously (128 B): element (row, k) lives at
//     [k / 8][row][k % 8]          ("K-blocked")
// so that  SBO (next 8 rows) = 128 B  and  LBO (next 8 k) = rows * 16 B.
// The plan pre-tiles the blocks that way once; prepare_B() casts B to bf16/fp16 and re-tiles
// it to Bq[Kpad/8][Npad][8] once per B.  With that layout every stage is filled by plain 1-D
// TMA bulk copies (cp.async.bulk.shared.global, UBLKCP): one for the block (bs*bs*2 B) and one
// for the bs x Ntile slab of Bq (contiguous when the CTA covers whole rows of Bq).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one
// elected lane), warps 2..5 = epilogue (tcgen05.ld 32x32b -> coalesced 128-byte row stores).
#include "common.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

namespace cuspmm_b200 {
namespace bsrtc {

constexpr int kThreads = 192;
constexpr int kMaxTileN = 512;     // columns of C per CTA (4 UMMA M-tiles)

using pipe::smem_u32;
using pipe::mbar_init;
using pipe::mbar_expect_tx;
using pipe::bulk_g2s;
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) { pipe::mbar_wait<0>(bar, parity); }
__device__ __forceinline__ void umma_commit(uint64_t *bar) {   // arrives on bar when all prior MMAs are done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, no swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version = 1, [61,64) layout = 0.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 @ [4,6),
// a/b format (F16 = 0, BF16 = 1) @ [7,10) / [10,13), both K-major, N>>3 @ [17,23), M>>4 @ [24,29).
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt, uint32_t M, uint32_t N) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// ring depth (measured, cfg4: 16x16 4 stages 0.275 ms, 6 stages 0.38 ms; 32x32 2 stages 0.192 ms, 3 stages 0.180 ms: the kernel
// is bound by the bytes in flight from L2, ncu: 168 registers -> 2 CTAs per SM).  First sizing:
// so that 3 CTAs fit on an SM for both block sizes (16.5 KB x 4 = 66 KB, 34 KB x 2 = 68 KB):
// the prologue (TMEM alloc, barrier init) and the epilogue of one CTA then overlap the main loops of
// the other two.  With 4 stages of 34 KB only one CTA fits and the 32x32 kernel was 2x slower.
#ifndef CUSPMM_BSR16_STAGES
#define CUSPMM_BSR16_STAGES 4
#endif
#ifndef CUSPMM_BSR32_STAGES
#define CUSPMM_BSR32_STAGES 3
#endif
template <int BS>
struct Smem {
    static constexpr int kStages = BS == 16 ? CUSPMM_BSR16_STAGES : CUSPMM_BSR32_STAGES;
    static constexpr uint32_t kBlockBytes = BS * BS * 2;
    static constexpr uint32_t kSlabBytes = BS * kMaxTileN * 2;          // bs k-rows x 512 n x 16 bit
    static constexpr uint32_t kStageBytes = kSlabBytes + kBlockBytes;
    static constexpr uint32_t kTotal = kStages * kStageBytes + (2 * kStages + 1) * 8 + 16 + 128;
};

// grid = (numBlockRows, ceil(Npad / 512)); FMT: 1 = bf16, 0 = fp16
template <int BS, int FMT>
__global__ void __launch_bounds__(kThreads)
bsr_tc_kernel(const uint32_t *__restrict__ blockRowPtrs, const uint32_t *__restrict__ blockColIdxs,
              const uint16_t *__restrict__ blocksQ,   // [numBlocks][BS/8][BS][8]
              const uint16_t *__restrict__ Bq,        // [Kpad/8][Npad][8]
              uint32_t Npad, uint32_t N, float *__restrict__ C, size_t ldc) {
    extern __shared__ __align__(128) unsigned char smem[];
    using S = Smem<BS>;
    constexpr int kStages = S::kStages;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + kStages * S::kStageBytes);
    uint64_t *empty = full + kStages;
    uint64_t *accum_full = empty + kStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_full + 1);

    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t R = blockIdx.x;
    const uint32_t n0 = blockIdx.y * kMaxTileN;
    const uint32_t ntile = min((uint32_t)kMaxTileN, Npad - n0);      // multiple of 128
    const uint32_t tiles = ntile / 128;
    const uint32_t bstart = __ldg(blockRowPtrs + R), bend = __ldg(blockRowPtrs + R + 1);
    const uint32_t nblk = bend - bstart;
    constexpr uint32_t kTmemCols = 4 * BS;                            // 64 or 128: power of two >= 32

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(accum_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            for (uint32_t i = 0; i < nblk; ++i) {
                const uint32_t s = i % kStages, it = i / kStages;
                if (it > 0) mbar_wait(empty + s, (it - 1) & 1);
                unsigned char *slab = smem + s * S::kStageBytes;
                unsigned char *blk = slab + S::kSlabBytes;
                const uint32_t b = bstart + i;
                const uint32_t bcol = __ldg(blockColIdxs + b);
                const uint32_t slabBytes = (BS / 8) * ntile * 16;
                mbar_expect_tx(full + s, slabBytes + S::kBlockBytes);
                bulk_g2s(blk, blocksQ + (size_t)b * BS * BS, S::kBlockBytes, full + s);
                const uint16_t *src = Bq + ((size_t)bcol * (BS / 8) * Npad + n0) * 8;
                if (ntile == Npad) {
                    bulk_g2s(slab, src, slabBytes, full + s);             // BS/8 k-groups are contiguous
                } else {
#pragma unroll
                    for (int kb = 0; kb < BS / 8; ++kb)
                        bulk_g2s(slab + (size_t)kb * ntile * 16, src + (size_t)kb * Npad * 8, ntile * 16, full + s);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(FMT, 128, BS);
            for (uint32_t i = 0; i < nblk; ++i) {
                const uint32_t s = i % kStages, it = i / kStages;
                mbar_wait(full + s, it & 1);
                tc_fence_after();
                const uint32_t slab = smem_u32(smem + s * S::kStageBytes);
                const uint32_t blk = slab + S::kSlabBytes;
#pragma unroll
                for (int ks = 0; ks < BS / 16; ++ks) {
                    // B operand: block k-groups 2ks, 2ks+1: [kb][BS rows][8]  -> LBO = BS*16, SBO = 128
                    const uint64_t bdesc = make_desc(blk + ks * 2 * BS * 16, BS * 16, 128);
                    for (uint32_t t = 0; t < tiles; ++t) {
                        // A operand: slab k-groups 2ks, 2ks+1, rows n = 128t .. 128t+127 -> LBO = ntile*16, SBO = 128
                        const uint64_t adesc = make_desc(slab + ks * 2 * ntile * 16 + t * 128 * 16, ntile * 16, 128);
                        umma_f16(tmem_base + t * BS, adesc, bdesc, idesc, (i | ks) ? 1u : 0u);
                    }
                }
                umma_commit(empty + s);          // stage s may be refilled once these MMAs have read it
            }
            umma_commit(accum_full);             // all accumulators final
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5)
        const uint32_t q = warp & 3;             // TMEM lane quarter this warp may access
        if (nblk > 0) {
            mbar_wait(accum_full, 0);
            tc_fence_after();
        }
        for (uint32_t t = 0; t < tiles; ++t) {
            uint32_t v[BS];
            if (nblk > 0) {
                const uint32_t taddr = tmem_base + ((q * 32u) << 16) + t * BS;
                if constexpr (BS == 16) {
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                                 : "r"(taddr));
                } else {
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                                 : "r"(taddr));
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int j = 0; j < BS; ++j) v[j] = 0u;       // a block row without blocks is a zero row of C
            }
            const uint32_t n = n0 + t * 128 + q * 32 + lane;
            if (n < N) {
                float *cp = C + (size_t)R * BS * ldc + n;
#pragma unroll
                for (int j = 0; j < BS; ++j) __stcs(cp + (size_t)j * ldc, __uint_as_float(v[j]));   // 32 lanes = one 128-byte line
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ==================================================================================== panel kernel (union walk)
// The block-row kernel above drags a bs x 512 slab of B through L2 -> SM for EVERY stored block (32 flop per byte at 32x32;
// ncu r01: L2 -> SM 12-16 TB/s, tensor pipe 17-20 %).  This kernel gives a CTA a PANEL of P consecutive block rows x 128 columns
// of C (one UMMA M-tile) and walks the UNION of the panel's block columns in ascending order: the slab of B belonging to block
// column j is fetched ONCE and multiplied against every block row of the panel that stores a block in column j.  At 10 %
// block density a panel of 11 block rows touches 69 % of the block columns for 1.1 blocks per column: 0.62 slabs per block
// instead of 1 (P = 22 at 16x16: 0.41).  Accumulators: D_p = TMEM columns [p * bs, (p + 1) * bs), P * bs <= 512.
//   warp 0      walker: P cursors (lane = block row) over the panel's block column indices (staged in shared memory), REDUX.min =
//               next block column of the union, ballot = the rows that hit it; writes a stage descriptor (column, hit mask, block ids)
//   warps 2..5  copiers: copier w fills the stages i = w (mod 4) with cp.async (16 bytes per lane: the slab's 2 KB pieces and the
//               hit blocks), one commit group per stage, and publishes a stage once its group has completed
//               (cp.async.wait_group + fence.proxy.async, the operands are read by the tensor core through the async proxy);
//               afterwards the same warps run the epilogue (tcgen05.ld.x16 per block row -> 128-byte row stores)
//   warp 1      MMA issuer: for every hit row p: D_p += slab^T x block^T (bs/16 x tcgen05.mma M128 N=bs K16); commit frees the stage
// First version (TMA bulk copies issued by the walker warp, profiles/r02_bsr_panel_probe.jsonl): 0.60 ms at 32x32 against 0.18 ms
// for the block-row kernel -- a stage was 5-6 small bulk copies and one warp retires about one bulk copy per ~200 clocks.
// Grid = (ceil(numBlockRows / P), Npad / 128); the host picks P so that the grid is just under a whole number of waves.
constexpr int kHitsPerStage = 4;
constexpr int kCopiers = 4;
template <int BS>
struct PanelSmem {
    static constexpr int kStages = BS == 16 ? 24 : 12;
    static constexpr uint32_t kBlockBytes = BS * BS * 2;
    static constexpr uint32_t kSlabBytes = BS * 128 * 2;                     // bs k-rows x 128 n x 16 bit
    static constexpr uint32_t kStageBytes = kSlabBytes + kHitsPerStage * kBlockBytes;
    static constexpr uint32_t kDescWords = 8;                                // column, hit mask, block ids[4], pad
    // one bitmap of block columns per block row of the panel (built once by the whole CTA): the union walk then needs, per 32
    // block columns, one LDS + one REDUX.or, and per block column of the union one ballot -- instead of LDS -> REDUX.min ->
    // ballot per column on cursors.  32 rows x 64 words: panels of matrices with at most 2048 block columns.
    static constexpr uint32_t kBmWords = 64;
    static constexpr uint32_t kTotal = kStages * kStageBytes + (3 * kStages + 1) * 8 + kStages * kDescWords * 4 + 32 * kBmWords * 4 + 16 + 128;
    static_assert(kTotal <= 232448, "more than 227 KB of shared memory");
};

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}

template <int BS, int FMT>
__global__ void __launch_bounds__(kThreads)
bsr_tc_panel_kernel(const uint32_t *__restrict__ blockRowPtrs, const uint32_t *__restrict__ blockColIdxs,
                    const uint16_t *__restrict__ blocksQ,   // [numBlocks][BS/8][BS][8]
                    const uint16_t *__restrict__ Bq,        // [Kpad/8][Npad][8]
                    uint32_t numBlockRows, uint32_t numBlockCols, uint32_t P, uint32_t tmemCols, uint32_t Npad, uint32_t N,
                    float *__restrict__ C, size_t ldc) {
    extern __shared__ __align__(128) unsigned char smem[];
    using S = PanelSmem<BS>;
    constexpr int kStages = S::kStages;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + kStages * S::kStageBytes);   // stage operands landed   (copier -> issuer)
    uint64_t *empty = full + kStages;                                                   // stage consumed          (issuer -> walker)
    uint64_t *desc_full = empty + kStages;                                              // descriptor written      (walker -> copier)
    uint64_t *accum_full = desc_full + kStages;
    uint32_t *desc = reinterpret_cast<uint32_t *>(accum_full + 1);                      // [kStages][kDescWords]
    uint32_t *bm = desc + kStages * S::kDescWords;                                      // [32 rows][kBmWords]: block columns of each panel row
    uint32_t *tmem_slot = bm + 32 * S::kBmWords;

    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t R0 = blockIdx.x * P;
    const uint32_t n0 = blockIdx.y * 128;
    const uint32_t bmW = (((numBlockCols + 31u) >> 5) < S::kBmWords) ? ((numBlockCols + 31u) >> 5) : S::kBmWords;
    for (uint32_t t = threadIdx.x; t < 32 * S::kBmWords; t += kThreads) bm[t] = 0u;
    __syncthreads();
    for (uint32_t p = 0; p < P && R0 + p < numBlockRows; ++p) {
        const uint32_t b0 = __ldg(blockRowPtrs + R0 + p), b1 = __ldg(blockRowPtrs + R0 + p + 1);
        for (uint32_t t = b0 + threadIdx.x; t < b1; t += kThreads) {
            const uint32_t c = __ldg(blockColIdxs + t);
            atomicOr(bm + p * S::kBmWords + (c >> 5), 1u << (c & 31u));
        }
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); mbar_init(desc_full + s, 1); }
        mbar_init(accum_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ walker: the union of the panel's block columns
        // lane = block row of the panel; cur = its next block (block columns ascend inside a block row, as the converter and
        // scipy write them, so the bitmap order is the storage order)
        uint32_t cur = 0;
        if (lane < P && R0 + lane < numBlockRows) cur = __ldg(blockRowPtrs + R0 + lane);
        uint32_t i = 0;
        auto open_stage = [&]() -> uint32_t * {          // wait until stage i % kStages has been consumed, return its descriptor
            const uint32_t s = i % kStages, it = i / kStages;
            if (it > 0) mbar_wait(empty + s, (it - 1) & 1);
            return desc + s * S::kDescWords;
        };
        for (uint32_t w = 0; w < bmW; ++w) {
            const uint32_t word = bm[lane * S::kBmWords + w];
            uint32_t uni = __reduce_or_sync(0xFFFFFFFFu, word);
            while (uni) {
                const uint32_t bit = (uint32_t)__ffs((int)uni) - 1u;
                uni &= uni - 1u;
                const uint32_t j = w * 32u + bit;
                const bool mine = (word >> bit) & 1u;
                uint32_t hit = __ballot_sync(0xFFFFFFFFu, mine);
                while (hit) {
                    uint32_t take = 0, h = hit;
#pragma unroll
                    for (int k = 0; k < kHitsPerStage; ++k) { take |= h & (0u - h); h &= h - 1u; }
                    uint32_t *d = open_stage();
                    if ((take >> lane) & 1u) d[2 + __popc(take & ((1u << lane) - 1u))] = cur;      // block ids in slot order
                    if (lane == 0) { d[0] = j; d[1] = take; }
                    __syncwarp();
                    if (lane == 0) pipe::mbar_arrive(desc_full + i % kStages);
                    hit &= ~take;
                    ++i;
                }
                if (mine) ++cur;
            }
        }
        // end markers: one per copier (an empty hit mask); the issuer stops at the first
        for (int k = 0; k < kCopiers; ++k) {
            uint32_t *d = open_stage();
            if (lane == 0) {
                d[1] = 0u;
                pipe::mbar_arrive(desc_full + i % kStages);
            }
            ++i;
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(FMT, 128, BS);
            uint32_t touched = 0;
            for (uint32_t i = 0;; ++i) {
                const uint32_t s = i % kStages, it = i / kStages;
                mbar_wait(full + s, it & 1);
                uint32_t m = desc[s * S::kDescWords + 1];
                if (m == 0u) break;
                // (the generic -> async proxy fence is executed by the copier warps after cp.async.wait_group, before they
                //  arrive on `full`; a second one here costs the issuer several hundred clocks per stage)
                tc_fence_after();
                const uint32_t slab = smem_u32(smem + s * S::kStageBytes);
                uint32_t blk = slab + S::kSlabBytes;
                while (m) {
                    const uint32_t p = (uint32_t)__ffs((int)m) - 1u;
                    m &= m - 1u;
#pragma unroll
                    for (int ks = 0; ks < BS / 16; ++ks) {
                        const uint64_t bdesc = make_desc(blk + ks * 2 * BS * 16, BS * 16, 128);       // block k-groups 2ks, 2ks+1
                        const uint64_t adesc = make_desc(slab + ks * 2 * 128 * 16, 128 * 16, 128);    // slab k-groups 2ks, 2ks+1
                        umma_f16(tmem_base + p * BS, adesc, bdesc, idesc, (((touched >> p) & 1u) | (uint32_t)ks) ? 1u : 0u);
                    }
                    touched |= 1u << p;
                    blk += S::kBlockBytes;
                }
                umma_commit(empty + s);          // stage s may be refilled once these MMAs have read it
            }
            umma_commit(accum_full);             // all accumulators final
        }
    } else {
        // ------------------------------------------------------------------ copiers (warps 2..5), then the epilogue
        const uint32_t w = warp - 2;
        // Up to LAG commit groups (= stages) of this warp are in flight.  A stage is published (fence.proxy.async + arrive on
        // `full`) as soon as its group has completed: when LAG groups are pending, or whenever the next descriptor is not there
        // yet.  (Publishing a stage only when a LATER one was issued tied the issuer to the walker: with the ring 4 * LAG stages
        // short the whole pipeline advanced 8 stages per round trip.)
        constexpr int LAG = BS == 16 ? 4 : 2;
        uint32_t pend[LAG];
        int npend = 0;
#pragma unroll
        for (int k = 0; k < LAG; ++k) pend[k] = 0xFFFFFFFFu;      // pend[0] = oldest
        auto retire_oldest = [&]() {              // npend >= 1: wait until only the npend - 1 younger groups are pending
            if (npend == 1) asm volatile("cp.async.wait_group 0;" ::: "memory");
            else if (npend == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else if (npend == 3) asm volatile("cp.async.wait_group 2;" ::: "memory");
            else asm volatile("cp.async.wait_group 3;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) pipe::mbar_arrive(full + pend[0]);
#pragma unroll
            for (int k = 0; k + 1 < LAG; ++k) pend[k] = pend[k + 1];
            pend[LAG - 1] = 0xFFFFFFFFu;
            --npend;
        };
        static_assert(LAG <= 4, "retire_oldest covers up to four pending groups");
        for (uint32_t i = w;;) {
            const uint32_t s = i % kStages;
            if (npend > 0) {                     // something in flight: do not block on the walker, retire instead
                const bool ready = __shfl_sync(0xFFFFFFFFu, pipe::mbar_try_wait(desc_full + s, (i / kStages) & 1) ? 1 : 0, 0) != 0;
                if (!ready) { retire_oldest(); continue; }
            } else {
                mbar_wait(desc_full + s, (i / kStages) & 1);
            }
            const uint32_t *d = desc + s * S::kDescWords;
            const uint32_t j = d[0], take = d[1];
            if (take == 0u) {                    // end of the walk: publish what is still in flight, pass the marker on
                while (npend > 0) retire_oldest();
                if (lane == 0) pipe::mbar_arrive(full + s);
                break;
            }
            if (npend == LAG) retire_oldest();   // make room: at most LAG groups in flight
            const uint32_t slab = smem_u32(smem + s * S::kStageBytes);
            const uint16_t *src = Bq + ((size_t)j * (BS / 8) * Npad + n0) * 8;
#pragma unroll
            for (int kb = 0; kb < BS / 8; ++kb)
#pragma unroll
                for (int r = 0; r < 4; ++r)      // 128 columns x 16 bytes per k-group
                    cp_async16(slab + kb * 2048 + (r * 32 + lane) * 16, src + ((size_t)kb * Npad + r * 32 + lane) * 8);
            const uint32_t hits = (uint32_t)__popc(take);
            for (uint32_t h = 0; h < hits; ++h) {
                const uint16_t *bsrc = blocksQ + (size_t)d[2 + h] * BS * BS;
                const uint32_t bdst = slab + S::kSlabBytes + h * S::kBlockBytes;
#pragma unroll
                for (int r = 0; r < (int)(S::kBlockBytes / 512); ++r)
                    cp_async16(bdst + (r * 32 + lane) * 16, bsrc + (r * 32 + lane) * 8);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
#pragma unroll
            for (int k = 0; k < LAG; ++k)
                if (k == npend) pend[k] = s;
            ++npend;
            i += kCopiers;
        }
        // ------------------------------------------------------------------ epilogue
        const uint32_t q = warp & 3;             // TMEM lane quarter this warp may access
        mbar_wait(accum_full, 0);
        tc_fence_after();
        const uint32_t n = n0 + q * 32 + lane;
        for (uint32_t p = 0; p < P && R0 + p < numBlockRows; ++p) {
            const uint32_t R = R0 + p;
            const bool has = __ldg(blockRowPtrs + R + 1) > __ldg(blockRowPtrs + R);      // a block row without blocks is a zero row of C
#pragma unroll
            for (int half = 0; half < BS / 16; ++half) {
                uint32_t v[16];
                if (has) {
                    const uint32_t taddr = tmem_base + ((q * 32u) << 16) + p * BS + half * 16;
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
                                 "tcgen05.wait::ld.sync.aligned;"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                                 : "r"(taddr) : "memory");
                } else {
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) v[jj] = 0u;
                }
                if (n < N) {
                    float *cp = C + ((size_t)R * BS + half * 16) * ldc + n;
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) __stcs(cp + (size_t)jj * ldc, __uint_as_float(v[jj]));   // 32 lanes = one 128-byte line
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmemCols) : "memory");
    }
}

// ------------------------------------------------------------------ operand preparation
template <typename T> __device__ __forceinline__ uint16_t cvt16(float x);
template <> __device__ __forceinline__ uint16_t cvt16<__nv_bfloat16>(float x) { return __bfloat16_as_ushort(__float2bfloat16_rn(x)); }
template <> __device__ __forceinline__ uint16_t cvt16<__half>(float x) { return __half_as_ushort(__float2half_rn(x)); }

// blocks fp32 [nb][bs][bs] (row-major, the reference's layout) -> [nb][bs/8][bs][8] 16-bit
template <typename T>
__global__ void tile_blocks_kernel(const float *__restrict__ blocks, uint64_t total, uint32_t bs, uint16_t *__restrict__ out) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over output elements
    if (idx >= total) return;
    const uint32_t per = bs * bs;
    const uint64_t b = idx / per;
    const uint32_t o = (uint32_t)(idx % per);
    const uint32_t e = o % 8, i = (o / 8) % bs, kb = o / (8 * bs);
    out[idx] = cvt16<T>(blocks[b * per + (uint64_t)i * bs + kb * 8 + e]);
}

// B fp32 [K][N] (ldb) -> Bq[Kpad/8][Npad][8] 16-bit, zero padded; one thread per (kb, n): 8 coalesced reads, one 16-byte write
template <typename T>
__global__ void tile_B_kernel(const float *__restrict__ B, uint32_t K, uint32_t N, size_t ldb, uint32_t Kpad8, uint32_t Npad,
                              uint4 *__restrict__ out) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)Kpad8 * Npad) return;
    const uint32_t n = (uint32_t)(idx % Npad), kb = (uint32_t)(idx / Npad);
    uint16_t h[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const uint32_t k = kb * 8 + e;
        h[e] = (k < K && n < N) ? cvt16<T>(__ldg(B + (size_t)k * ldb + n)) : (uint16_t)0;
    }
    uint4 v;
    v.x = h[0] | ((uint32_t)h[1] << 16); v.y = h[2] | ((uint32_t)h[3] << 16);
    v.z = h[4] | ((uint32_t)h[5] << 16); v.w = h[6] | ((uint32_t)h[7] << 16);
    out[idx] = v;
}

} // namespace bsrtc
} // namespace cuspmm_b200

using namespace cuspmm_b200;

struct cuspmmBsrTcPlan_s {
    const uint32_t *blockRowPtrs = nullptr, *blockColIdxs = nullptr;   // borrowed (caller keeps them alive)
    uint16_t *blocksQ = nullptr, *Bq = nullptr;
    uint32_t numBlockRows = 0, numBlocks = 0, bs = 0, K = 0, Kpad = 0, maxN = 0, maxNpad = 0, N = 0, Npad = 0;
    int type = 0;
    int dev = 0;
    bool ownsBuffers = true;           // false: blocksQ / Bq were lent by the caller (host pipeline: cached staging buffers)
};

static bool plan_on_current_device(const cuspmmBsrTcPlan_s *p) {
    int dev = -1;
    return cudaGetDevice(&dev) == cudaSuccess && dev == p->dev;
}

namespace cuspmm_b200 {
// blocksQ_ext / Bq_ext: device buffers lent to the plan (numBlocks * bs * bs and Kpad * maxNpad 16-bit elements); null: the plan allocates
int bsr_tc_plan_create_impl(cuspmmBsrTcPlan *out, const uint32_t *blockRowPtrs, const uint32_t *blockColIdxs,
                                         const float *blocks, uint32_t numBlockRows, uint32_t numBlocks, uint32_t bs,
                                         uint32_t K, uint32_t maxN, cuspmmBlockType type, void *stream, void *blocksQ_ext, void *Bq_ext) {
    CUSPMM_REQUIRE(out && blockRowPtrs && maxN >= 1, "bad arguments");
    if (bs != 16 && bs != 32)
        return set_error(CUSPMM_ERR_UNSUPPORTED, "tensor-core BSR supports 16x16 and 32x32 blocks (got %u)", bs);
    CUSPMM_REQUIRE(type == CUSPMM_BLK_BF16 || type == CUSPMM_BLK_FP16, "unknown block type %d", (int)type);
    cudaStream_t st = as_stream(stream);
    auto *p = new cuspmmBsrTcPlan_s();
    p->blockRowPtrs = blockRowPtrs; p->blockColIdxs = blockColIdxs;
    p->numBlockRows = numBlockRows; p->numBlocks = numBlocks; p->bs = bs; p->K = K; p->type = (int)type;
    p->Kpad = (K + bs - 1) / bs * bs;
    p->maxN = maxN; p->maxNpad = (maxN + 127) / 128 * 128;
    cudaGetDevice(&p->dev);
    const uint64_t total = (uint64_t)numBlocks * bs * bs;
    if (blocksQ_ext && Bq_ext) {
        p->blocksQ = static_cast<uint16_t *>(blocksQ_ext);
        p->Bq = static_cast<uint16_t *>(Bq_ext);
        p->ownsBuffers = false;
    } else {
        cudaError_t me = cudaMalloc(&p->blocksQ, (total ? total : 1) * 2);
        if (me == cudaSuccess) me = cudaMalloc(&p->Bq, (size_t)p->Kpad * p->maxNpad * 2);
        if (me != cudaSuccess) {
            cudaFree(p->blocksQ); cudaFree(p->Bq); delete p;
            return set_error(CUSPMM_ERR_CUDA, "cudaMalloc of the tensor-core BSR plan failed: %s", cudaGetErrorString(me));
        }
    }
    if (total) {
        const unsigned grid = (unsigned)((total + 255) / 256);
        if (type == CUSPMM_BLK_BF16) bsrtc::tile_blocks_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(blocks, total, bs, p->blocksQ);
        else bsrtc::tile_blocks_kernel<__half><<<grid, 256, 0, st>>>(blocks, total, bs, p->blocksQ);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            if (p->ownsBuffers) { cudaFree(p->blocksQ); cudaFree(p->Bq); }
            delete p;
            return set_error(CUSPMM_ERR_CUDA, "launch of tile_blocks_kernel failed: %s", cudaGetErrorString(e));
        }
        count_launch();
    }
    *out = p;
    return CUSPMM_OK;
}
size_t bsr_tc_blocks_bytes(uint32_t numBlocks, uint32_t bs) { return (size_t)(numBlocks ? numBlocks : 1) * bs * bs * 2; }
size_t bsr_tc_B_bytes(uint32_t K, uint32_t bs, uint32_t maxN) {
    return (size_t)((K + bs - 1) / bs * bs) * ((maxN + 127) / 128 * 128) * 2;
}
} // namespace cuspmm_b200

extern "C" int cuspmm_bsr_tc_plan_create(cuspmmBsrTcPlan *out, const uint32_t *blockRowPtrs, const uint32_t *blockColIdxs,
                                         const float *blocks, uint32_t numBlockRows, uint32_t numBlocks, uint32_t bs,
                                         uint32_t K, uint32_t maxN, cuspmmBlockType type, void *stream) {
    return cuspmm_b200::bsr_tc_plan_create_impl(out, blockRowPtrs, blockColIdxs, blocks, numBlockRows, numBlocks, bs, K, maxN, type,
                                                stream, nullptr, nullptr);
}

extern "C" int cuspmm_bsr_tc_prepare_B(cuspmmBsrTcPlan p, const float *B, uint32_t N, size_t ldb, void *stream) {
    CUSPMM_REQUIRE(p && B && N >= 1 && N <= p->maxN && ldb >= N, "bad arguments (N=%u, maxN=%u)", N, p ? p->maxN : 0);
    CUSPMM_REQUIRE(plan_on_current_device(p), "the plan was created on device %d; make it current before prepare_B", p->dev);
    p->N = N;
    p->Npad = (N + 127) / 128 * 128;
    const uint64_t total = (uint64_t)(p->Kpad / 8) * p->Npad;
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (p->type == CUSPMM_BLK_BF16)
        bsrtc::tile_B_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(B, p->K, N, ldb, p->Kpad / 8, p->Npad, reinterpret_cast<uint4 *>(p->Bq));
    else
        bsrtc::tile_B_kernel<__half><<<grid, 256, 0, as_stream(stream)>>>(B, p->K, N, ldb, p->Kpad / 8, p->Npad, reinterpret_cast<uint4 *>(p->Bq));
    CUSPMM_LAUNCH_CHECK("tile_B_kernel");
    return CUSPMM_OK;
}

template <int BS, int FMT>
static int launch_tc(cuspmmBsrTcPlan p, float *C, size_t ldc, cudaStream_t st) {
    // The block-row kernel is the default.  The panel kernel (union walk) is correct (tests/test_gpu_bsr_tc.py runs it through
    // the hook) but slower as built, on BASELINE configs[3] (profiles/r02_bsr_panel_probe.jsonl; block-row kernel 0.18 / 0.28 ms):
    //   TMA bulk copies issued by the walking warp           0.60 ms (32x32, P = 11) / 1.41 ms (16x16, P = 22): 5-6 small copies per
    //                                                        stage, one warp retires about one bulk copy per ~200 clk
    //   walker + four cp.async copier warps                  0.40 ms / 0.98 ms -- and it stayed there through four changes of the
    //       pipeline (deeper commit groups, a word-parallel walk on per-row bitmaps, no proxy fence in the issuer, stages published
    //       as soon as they land): ~620 clocks per stage whatever the stage does.  Little's law: a panel CTA has one 12-stage ring
    //       of 16 KB stages that are on average 64 % full = 120 KB in flight, and a stage lives ~7 400 clocks (L2 latency under
    //       load + three mbarrier hand-offs: walker -> copier -> issuer -> walker), i.e. 16 B/clk per SM; the block-row kernel keeps
    //       204 KB in flight for ~4 500 clocks = 45 B/clk.  Needing 0.62x the bytes does not pay for moving them at 0.37x the rate.
    // Tuning hooks: CUSPMM_BSR_PANEL = 1 selects the panel kernel, CUSPMM_BSR_P its panel height.
    static const int forcePanel = getenv("CUSPMM_BSR_PANEL") ? atoi(getenv("CUSPMM_BSR_PANEL")) : -1;
    static const int forceP = getenv("CUSPMM_BSR_P") ? atoi(getenv("CUSPMM_BSR_P")) : 0;
    const uint32_t ytiles = p->Npad / 128;
    const uint32_t sms = (uint32_t)sm_count();
    constexpr uint32_t Pmax = 512 / BS > 32 ? 32 : 512 / BS;
    uint32_t P = 1;
    for (uint32_t w = 1; w <= 64; ++w) {                 // the smallest whole number of waves whose panels fit into TMEM
        const uint32_t want = (w * sms) / ytiles;
        if (want == 0) continue;
        P = (p->numBlockRows + want - 1) / want;
        if (P <= Pmax) break;
        P = Pmax;
    }
    if (forceP > 0) P = (uint32_t)forceP > Pmax ? Pmax : (uint32_t)forceP;
    if (P < 1) P = 1;
    const bool usePanel = forcePanel > 0 && p->Kpad / BS <= 32u * bsrtc::PanelSmem<BS>::kBmWords;     // bitmaps hold 2048 block columns
    if (usePanel) {
        auto kern = bsrtc::bsr_tc_panel_kernel<BS, FMT>;
        CUSPMM_CUDA(set_smem_once(kern, bsrtc::PanelSmem<BS>::kTotal));
        uint32_t cols = 32;
        while (cols < P * BS) cols <<= 1;
        dim3 grid((p->numBlockRows + P - 1) / P, ytiles);
        kern<<<grid, bsrtc::kThreads, bsrtc::PanelSmem<BS>::kTotal, st>>>(p->blockRowPtrs, p->blockColIdxs, p->blocksQ, p->Bq,
                                                                           p->numBlockRows, p->Kpad / BS, P, cols, p->Npad, p->N, C, ldc);
        CUSPMM_LAUNCH_CHECK("bsr_tc_panel_kernel");
        return CUSPMM_OK;
    }
    auto kern = bsrtc::bsr_tc_kernel<BS, FMT>;
    CUSPMM_CUDA(set_smem_once(kern, bsrtc::Smem<BS>::kTotal));
    dim3 grid(p->numBlockRows, (p->Npad + bsrtc::kMaxTileN - 1) / bsrtc::kMaxTileN);
    kern<<<grid, bsrtc::kThreads, bsrtc::Smem<BS>::kTotal, st>>>(p->blockRowPtrs, p->blockColIdxs, p->blocksQ, p->Bq, p->Npad, p->N, C, ldc);
    CUSPMM_LAUNCH_CHECK("bsr_tc_kernel");
    return CUSPMM_OK;
}

extern "C" int cuspmm_bsr_tc_run(cuspmmBsrTcPlan p, float *C, size_t ldc, void *stream) {
    CUSPMM_REQUIRE(p && C && p->N >= 1, "prepare_B must be called before run");
    CUSPMM_REQUIRE(ldc >= p->N, "ldc must be >= N");
    CUSPMM_REQUIRE(plan_on_current_device(p), "the plan was created on device %d; make it current before run", p->dev);
    if (p->numBlockRows == 0) return CUSPMM_OK;
    cudaStream_t st = as_stream(stream);
    if (p->bs == 16) return p->type == CUSPMM_BLK_BF16 ? launch_tc<16, 1>(p, C, ldc, st) : launch_tc<16, 0>(p, C, ldc, st);
    return p->type == CUSPMM_BLK_BF16 ? launch_tc<32, 1>(p, C, ldc, st) : launch_tc<32, 0>(p, C, ldc, st);
}

extern "C" int cuspmm_bsr_tc_plan_destroy(cuspmmBsrTcPlan p) {
    if (!p) return CUSPMM_OK;
    if (p->ownsBuffers) {
        cudaFree(p->blocksQ);
        cudaFree(p->Bq);
    }
    delete p;
    return CUSPMM_OK;
}
