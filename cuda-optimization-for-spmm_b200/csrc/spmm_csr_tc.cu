// spmm_csr_tc.cu -- CSR SpMM on the tensor cores (variant 8, "tensor"): tiles of A are made DENSE in shared memory on the
// fly and multiplied with tcgen05.mma; fp32 accuracy comes from a three-product split of both operands.
//
// Why.  The fp32 CUDA-core kernels (variants 1-7) need one distinct element of B per FMA and are bound by the SM's operand
// bandwidth (4.1 ms on 25605^2 x 512 at 10 %, 22 % of the FMA pipe).  A B200 tensor core does 2048 tf32 /
// 4096 bf16 FMAs per clock and SM against 128 fp32 FMAs on the CUDA cores: from a few percent density upwards it is cheaper to
// multiply the zeros as well.  Work is then independent of nnz:  M x K x N x (1 + 1/2 + 1/2) tensor FMAs.
//
// Arithmetic (what replaces the reference's fp32 multiply-add, src/spmm/csr/spmm_csr_k3.cu:30-45 / spmm_csr.cpp):
//     a = a_t + r_a,  a_t = tf32(a) (round to nearest, 11 significant bits),  |r_a| <= 2^-11 |a|  (r_a exact in fp32)
//     b = b_t + r_b   likewise
//     a*b ~= a_t*b_t  (kind::tf32, exact products)  +  bf16(a)*bf16(r_b)  +  bf16(r_a)*bf16(b)   (kind::f16, bf16 operands)
// accumulated in fp32 in tensor memory.  Dropped / perturbed terms: r_a*r_b (2^-22) and the bf16 rounding (2^-8) of the factor
// that multiplies a remainder (2^-11) in each of the two corrections: |error| <= 2^-17 |a||b| per product in the worst case
// (7.6e-6; observed maximum over 10^6 one-entry rows 4.4e-6), inside the 1e-5 bound of the parity tests for every matrix;
// on sums of many terms the errors average out: ~5e-7 of sum|a||b| at 25605^2, the same order as fp32 accumulation.
//
// Accumulation.  tcgen05.mma does not round its fp32 accumulator to nearest: it truncates.  Every MMA step loses on average
// 2^-25.5 of |accumulator| TOWARDS ZERO -- a bias, not noise: after the 6400 steps of K = 25605 the first version was 1.2e-5
// of sum|a||b| off (7.7e-3 absolute on |C| ~ 20), outside the tolerance.  So the accumulators are drained into C every
// kFlushChunks = 64 chunks (256 steps: bias <= 5e-6 of |partial sum| even when all terms have one sign) by reductions that
// round to nearest in L2, and start over from zero.  Drain path: TMEM -> registers -> a swizzled 4 KB staging tile per epilogue
// warp -> cp.reduce.async.bulk.tensor (.add) -- ~8 000 clocks per drain; the first form, red.global.add.v4.f32 straight from
// registers, cost ~20 000 (the LSU retires about one reduction lane per 1.3 clocks) and is kept for a C that a tensor map
// cannot describe (unaligned, odd pitch) and for the one-CTA-per-tile hook.
//
// Data layout.  UMMA operands are K-major without swizzle: "core matrices" of 8 rows x 16 bytes, contiguous (128 B);
// element (row, k) of a tf32 operand lives at [k / 4][row][k % 4], of a bf16 operand at [k / 8][row][k % 8].
//   The two correction products share ONE bf16 operand of interleaved pairs: per k the A side holds (bf16(a), bf16(r_a)), the
//   B side (bf16(r_b), bf16(b)), so a K = 16 MMA sums both corrections over 8 k -- and the pair of an entry is one 32-bit store.
//   B: a prepare kernel (one pass over B, 12 bytes per element of HBM traffic) writes, per 256-column tile ct and 16-row
//      chunk c, one contiguous 32 KB record  [b_t: 4 k-groups x 256 columns x 4 fp32 | pairs: 4 k-groups x 256 x 4 pairs]
//      so that a stage of B is ONE TMA bulk copy.
//   A: a CTA owns 256 rows (two UMMA M = 128 blocks) x 256 columns of C = all 512 columns of tensor memory.  256 builder threads
//      own one row each: per 16-column chunk they clear their row of the stage (8 x 16 B) and scatter the row's non-zeros that
//      fall into the chunk (the rows are sorted by column: a cursor per thread; col/val arrive 4 at a time through a
//      per-row cp.async ring in shared memory, requested chunks ahead) as [a_t | pair], two 32-bit stores 16 KB apart.
//   D: TMEM columns [0, 256) = rows 0..127 of the tile, [256, 512) = rows 128..255; lane = row, column = n.
// Per chunk and M block: 2 x tcgen05.mma kind::tf32 (M128 N256 K8) + 2 x kind::f16 (M128 N256 K16) = 1024 tensor clocks
// per chunk against 32 KB of B from L2 (32 B/clk/SM) and ~410 non-zeros to place (10 % density).
//
// What bounds it (measured, 25605^2 x 512; clock64 around the issuer's waits, profiles/r02_tc_*): shared-memory bandwidth.
// The MMAs of a chunk read 96 KB of operands (B twice, once per M block: 768 clocks of the 128 B/clk pipe), TMA writes 32 KB,
// the builders clear 32 KB and scatter: ~1500 clocks per chunk with no non-zeros at all, 2050 at 10 % (the scattered 4-byte
// stores, 16 bytes apart per row, cost ~540 of them).  The tensor pipe itself needs ~1160.
//
// Decomposition.  The work of a tile does not depend on its non-zeros, so tiles are spread over a persistent grid of one CTA
// per SM: floor(tiles / grid) whole tiles per CTA, the remaining tiles cut into equal runs of chunks ("stream-K").  All pieces
// (drains of one CTA, partial tiles of several) meet in C through reductions on a zeroed C; the pieces of one CTA arrive in
// program order, those of two CTAs that share a tile in arrival order (run-to-run differences in the last bits of those tiles).
//
// Warp roles (736 threads): warp 0 = TMA producer for B, warp 1 = TMEM allocator + MMA issuer of M block 0, warps 2..17 = builders
// (two lanes per row), warps 18..21 = epilogue (tcgen05.ld 32x32b.x32 -> 16-byte reductions / stores), warp 22 = MMA issuer of
// M block 1.
//
// Semantics that differ from the sparse kernels: a zero of A is multiplied with B, so an Inf/NaN anywhere in B would poison
// rows that never reference it.  The prepare kernel therefore raises a device flag when B holds a non-finite value (or one
// that tf32 rounding would overflow); the tensor kernel then exits at once and the launcher's second kernel -- a plain fp32
// warp-per-row kernel, which runs only if the flag is set -- computes C instead.  No host synchronisation either way.
#include "tmem_common.cuh"

#include <cuda_bf16.h>
#include <cub/device/device_scan.cuh>

namespace cuspmm_b200 {
namespace csrtc {

using namespace tmemk;
// The waits of this kernel are short and sit on the critical path of a 4-deep pipeline of ~1100-clock chunks: plain try_wait
// polling (no suspend-time hint; spinning on test_wait instead made no difference).
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) { pipe::mbar_wait<0>(bar, parity); }

constexpr int kTileN = 256, kKC = 16;
constexpr int kRowsPerCta = 256;                                     // rows of A one CTA builds: two UMMA M blocks of 128
constexpr int kEpilogue = 128;
constexpr int kIssuers = 2;                                           // MMA-issuing threads: one per M block (accumulator)
// LPR = builder threads per row (1, or 2: the entries of a row dealt to two adjacent lanes by parity)
constexpr int threads_for(int lpr) { return 64 + kRowsPerCta * lpr + kEpilogue + 32; }      // the last warp hosts the second issuer
constexpr uint32_t kABytes = 32768, kBBytes = 32768;                 // per stage of A (one CTA) / per chunk record of B (256 columns)
constexpr uint32_t kOffT = 0, kOffP = 16384;                         // tf32 | bf16 pairs inside a stage of A
constexpr uint32_t kPrefetchChunks = 12;                             // L2 prefetch distance of the B producer, in chunks
constexpr uint32_t kFlushChunks = 64;                                // 256 accumulation steps between two drains of the accumulators
// NCTA = 1: one CTA per tile of 256 rows x 256 columns.  NCTA = 2: a CTA PAIR (cluster of two, tcgen05 cta_group::2, UMMA M = 256)
// per tile of 512 rows x 256 columns: each CTA builds its own 256 rows of A but holds only HALF of every B chunk (128 of the
// 256 columns), so the MMAs read B from shared memory once per 256 rows instead of once per 128, TMA writes half as much, L2
// delivers half as much, and the space buys rings of 4 + 4 stages instead of 3 + 3.
// LPR = 2 is the layout for dense rows (>= 12 % density): two builder lanes per row AND a ring of 8 blocks (32 entries) per row instead
// of 4 -- a row that uses 8 entries per chunk empties a 16-entry ring faster than the copies land, and every chunk then waits for
// everything outstanding (50 % dense: 4.4 ms with 4 blocks) -- paid for with one stage of A (3 instead of 4).
template <int NCTA, int LPR = 1>
struct Cfg {
    // the builders need more than one chunk time (clear, scatter, proxy fence, arrive): with 3 stages the issuer found the
    // next chunk's A late by ~150 clocks every chunk.  A pair has the room for 4 (its B stages are half as large).
    static constexpr int kRingBlocks = (NCTA == 2 && LPR == 2) ? 8 : 4;
    static constexpr int kStages = (NCTA == 1 || LPR == 2) ? 3 : 4;   // stages of A = barriers (indexed by chunk % kStages)
    static constexpr int kAStages = kStages;
    static constexpr int kBStages = 3;                                // buffers of B (chunk % 3): TMA needs no builder latency covered
    static constexpr uint32_t kBStage = kBBytes / NCTA;               // bytes of a B chunk one CTA holds: [b_t | pairs]
    static constexpr uint32_t kBOffP = kBStage / 2;
    static constexpr uint32_t kBLbo = (kTileN / NCTA) * 16;           // next k group of the B operand
    static constexpr uint32_t kAOff = 0;
    static constexpr uint32_t kBOff = kAStages * kABytes;
    // pairs: 4 x 4 KB staging tiles (32 rows x 32 columns, 128-byte swizzle) for the accumulator drains through TMA reductions
    static constexpr uint32_t kOutOff = kBOff + kBStages * kBStage;
    static constexpr uint32_t kOutBytes = NCTA == 2 ? 4 * 4096 : 0;
    static constexpr uint32_t kRingOff = kOutOff + kOutBytes;         // [colIdxs | vals][slot 4][row 256][16 B]
    static constexpr uint32_t kRingBytes = 2 * kRingBlocks * kRowsPerCta * 16;
    static constexpr uint32_t kRingVals = kRingBlocks * kRowsPerCta * 16;       // offset of the values inside the ring
    static constexpr uint32_t kRingMask = (kRingBlocks - 1) << 2;               // entry index bits that select the block
    static constexpr uint32_t kBarOff = kRingOff + kRingBytes;
    static constexpr uint32_t kNumBars = 2 * kStages + 2;           // full, empty, accum_full, accum_empty
    static constexpr uint32_t kSmemTotal = kBarOff + kNumBars * 8 + 16 + 128;
    static_assert(kOutOff % 1024 == 0, "the staging tiles are addressed with the 128-byte swizzle pattern (1024-byte period)");
    static constexpr int kTileM = kRowsPerCta * NCTA;
    static_assert(kSmemTotal <= 232448, "more than 227 KB of shared memory");
};

struct Plan {
    uint32_t tilesN, chunks, grid, fullWaves, remTiles;
    uint32_t flushChunks;           // the accumulators are drained into C every so many chunks (see "Accumulation")
    uint64_t unitsPerCta;           // chunks of the remaining tiles per CTA
};

// the segments (tile, chunk range) of CTA c, in order; every warp role walks the same sequence
struct SegIter {
    uint32_t i, W, G, c, chunks;
    uint64_t u0, u1;
    __device__ SegIter(const Plan &pl, uint32_t cta) : i(0), W(pl.fullWaves), G(pl.grid), c(cta), chunks(pl.chunks) {
        const uint64_t total = (uint64_t)pl.remTiles * pl.chunks;
        u0 = (uint64_t)cta * pl.unitsPerCta;
        u1 = u0 + pl.unitsPerCta;
        if (u0 > total) u0 = total;
        if (u1 > total) u1 = total;
    }
    __device__ bool next(uint32_t &tile, uint32_t &kb, uint32_t &ke) {
        if (i < W) { tile = c + i * G; kb = 0; ke = chunks; ++i; return true; }
        if (u0 >= u1) return false;
        const uint32_t t = (uint32_t)(u0 / chunks);
        kb = (uint32_t)(u0 - (uint64_t)t * chunks);
        const uint64_t len = (u1 - u0) < (uint64_t)(chunks - kb) ? (u1 - u0) : (uint64_t)(chunks - kb);
        ke = kb + (uint32_t)len;
        tile = W * G + t;
        u0 += len;
        return true;
    }
};

// ---------------------------------------------------------------------------------------------- operand split
__device__ __forceinline__ float tf32_rn(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
// v = t + r: t on the tf32 grid (never rounded up to infinity), r the exact fp32 remainder (0 for a non-finite v)
__device__ __forceinline__ void split_tf32(float v, float &t, float &r) {
    t = tf32_rn(v);
    if (!(fabsf(t) < __int_as_float(0x7f800000))) t = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    r = (fabsf(v) < __int_as_float(0x7f800000)) ? v - t : 0.0f;
}
__device__ __forceinline__ uint16_t bf16_bits(float v) {
    // round to nearest, but towards zero where that would leave the finite range
    const __nv_bfloat16 h = fabsf(v) < 1.7e38f ? __float2bfloat16_rn(v) : __float2bfloat16_rz(v);
    return __bfloat16_as_ushort(h);
}
// (Tried: two bf16 pieces per operand and all four cross products in one bf16 operand -- one 8-byte store per entry, one kind
//  of MMA.  bf16 keeps 8 significant bits, two pieces 16: the dropped terms are 2^-15 |a||b|, 1.1e-5 observed on short rows.
//  Three pieces need six products.  The tf32 main product is what makes three products enough.)

// ---------------------------------------------------------------------------------------------- B -> tiled operand records
// grid (2 * chunks, tilesN), 256 threads: thread = one column n, 8 consecutive k.  A record is NCTA parts of 256 / NCTA columns,
// each [b_t: 4 k-groups x columns x 16 B | pairs: the same shape]
template <int NCTA>
__global__ void __launch_bounds__(256)
csr_tc_prepare_B(const float *__restrict__ B, uint32_t K, uint32_t N, size_t ldb, unsigned char *__restrict__ Bt, uint32_t chunks,
                 uint32_t *__restrict__ flag) {
    using CF = Cfg<NCTA>;
    constexpr uint32_t kCols = kTileN / NCTA;
    const uint32_t half = blockIdx.x & 1u, chunk = blockIdx.x >> 1, ct = blockIdx.y;
    const uint32_t n = ct * kTileN + threadIdx.x;
    const uint32_t part = threadIdx.x / kCols, nloc = threadIdx.x % kCols;
    const uint32_t k0 = blockIdx.x * 8u;
    float b[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = (n < N && k0 + j < K) ? __ldcs(B + (size_t)(k0 + j) * ldb + n) : 0.0f;
    float t[8], r[8];
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        split_tf32(b[j], t[j], r[j]);
        // non-finite, or so large that the tf32 rounding above had to be replaced by truncation
        bad |= !(fabsf(b[j]) < 1.7e38f);
    }
    if (__any_sync(0xFFFFFFFFu, bad) && lane_id() == 0) atomicOr(flag, 1u);
    unsigned char *rec = Bt + ((size_t)ct * chunks + chunk) * kBBytes + part * CF::kBStage;
    float4 *pt = reinterpret_cast<float4 *>(rec + (half * 2u) * CF::kBLbo + nloc * 16u);
    pt[0] = make_float4(t[0], t[1], t[2], t[3]);
    pt[kCols] = make_float4(t[4], t[5], t[6], t[7]);             // next k group
    // the two correction products share ONE bf16 operand: per k the pair (bf16(r_b), bf16(b)) here against (bf16(a), bf16(r_a)) on
    // the A side, so that a K = 16 MMA sums  bf16(a) bf16(r_b) + bf16(r_a) bf16(b)  over 8 k
    uint4 p0, p1;
    p0.x = bf16_bits(r[0]) | ((uint32_t)bf16_bits(b[0]) << 16);
    p0.y = bf16_bits(r[1]) | ((uint32_t)bf16_bits(b[1]) << 16);
    p0.z = bf16_bits(r[2]) | ((uint32_t)bf16_bits(b[2]) << 16);
    p0.w = bf16_bits(r[3]) | ((uint32_t)bf16_bits(b[3]) << 16);
    p1.x = bf16_bits(r[4]) | ((uint32_t)bf16_bits(b[4]) << 16);
    p1.y = bf16_bits(r[5]) | ((uint32_t)bf16_bits(b[5]) << 16);
    p1.z = bf16_bits(r[6]) | ((uint32_t)bf16_bits(b[6]) << 16);
    p1.w = bf16_bits(r[7]) | ((uint32_t)bf16_bits(b[7]) << 16);
    uint4 *pp = reinterpret_cast<uint4 *>(rec + CF::kBOffP + (half * 2u) * CF::kBLbo + nloc * 16u);
    pp[0] = p0;
    pp[kCols] = p1;
}

// ---------------------------------------------------------------------------------------------- MMA
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 @ [4,6), a/b format (BF16 = 1, TF32 = 2) @ [7,10) / [10,13),
// both operands K-major, N >> 3 @ [17,23), M >> 4 @ [24,29)
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt, uint32_t M, uint32_t N) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
template <int NCTA>
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (NCTA == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
template <int NCTA>
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {   // always accumulates
    if constexpr (NCTA == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
}
// arrives on bar (pair: on the barrier at this offset in BOTH CTAs) once every MMA issued so far by this thread has completed
template <int NCTA>
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    if constexpr (NCTA == 1) tc_commit(bar);
    else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
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
                 ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}
__device__ __forceinline__ bool mbar_test_cluster(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
// The per-chunk "my half is ready" of rank 1.  Relaxed: a release at cluster scope is a full memory barrier for the thread
// (ncu: ~70x the samples of an MMA issue on that one instruction, more than a chunk time), and there is nothing of this thread's
// to release -- what the tensor core of THIS CTA will read was written by the builders (each fenced its stores for the async
// proxy and arrived with release on the local barrier this thread has just acquired) and by TMA (complete_tx observed).
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint64_t *bar, uint32_t rank) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
                 "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}"
                 ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}
// wait for a phase that a thread of the OTHER CTA of the pair completes: acquire at cluster scope
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0, polls = 0;
    uint64_t t0 = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) break;
        if ((++polls & 4095u) == 0) {
            const uint64_t now = pipe::global_ns();
            if (t0 == 0) t0 = now;
            else if (now - t0 > pipe::kWaitLimitNs) __trap();
        }
    }
}

__device__ __forceinline__ uint32_t pick(const uint4 &v, uint32_t i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }
__device__ __forceinline__ float pick(const float4 &v, uint32_t i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

// cp.async with a source size: the bytes past src_bytes are written as zeros (the tail of the arrays)
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst_smem, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_zero16(uint32_t addr) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
}
__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// VEC: colIdxs and vals are 16-byte aligned (blocks of 4 entries are fetched with one cp.async each)
// NCTA = 2: launched in clusters of two CTAs; blockIdx.x / 2 is the pair, %cluster_ctarank the half.  Only rank 0 issues MMAs.
//   barriers (same offsets in both CTAs): see full / empty below; accum_full is arrived in BOTH CTAs by the multicast
//   tcgen05.commit of rank 0; accum_empty lives in rank 0 and collects the epilogue warps of both CTAs.
// SELL: the rows come from a sliced-ELL matrix (rowPtrs = slicePtrs; entry j of row r at slicePtrs[r / 32] + 32 j + r % 32, padding
// entries have column 0xFFFFFFFF and end the row): the builder's cursor then counts entries of the row, not positions in colIdxs.
template <bool VEC, int NCTA, int LPR, bool SELL>
__global__ void __launch_bounds__(threads_for(LPR), 1)
csr_tc_kernel(const uint32_t *__restrict__ rowPtrs, const uint32_t *__restrict__ colIdxs, const float *__restrict__ vals,
              uint32_t M, uint32_t nnzTotal, const unsigned char *__restrict__ Bt, uint32_t N, float *__restrict__ C, size_t ldc,
              Plan pl, const uint32_t *__restrict__ flag, int vecC, const __grid_constant__ CUtensorMap tmapC, int tmaDrain) {
    using CF = Cfg<NCTA, LPR>;
    constexpr int kStages = CF::kStages, kAStages = kStages;
    constexpr int kBuilders = kRowsPerCta * LPR, kThreads = threads_for(LPR);
    extern __shared__ __align__(128) unsigned char smem[];
    if (*flag) return;                                        // B holds a non-finite value: the fp32 kernel that follows computes C
    unsigned char *stA = smem + CF::kAOff;
    unsigned char *stB = smem + CF::kBOff;
    // ONE full / empty barrier per stage (mbarrier operations cost the issuing thread ~80 clocks each, and the MMA issuer is a
    // single thread: with separate barriers for A, B and the peer it spent more time on barriers than the MMAs of a chunk take):
    //   full[s]   <- the 8 builder warps of this CTA + the TMA producer (its arrive.expect_tx; the bytes complete it)
    //                + in rank 0 of a pair one arrive from rank 1, whose warp 1 waits for rank 1's own full[s]
    //   empty[s]  <- one tcgen05.commit per issuer (multicast into both CTAs of a pair): builders AND producer wait on it
    // TWO issuing threads, one per M block (independent accumulators): what a single thread does per chunk -- barrier test or wait,
    // fence, descriptors, 8 MMAs, commit -- took longer than the tensor pipe needs for the MMAs.
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + CF::kBarOff);
    uint64_t *empty = full + kStages;
    uint64_t *accum_full = empty + kStages;
    uint64_t *accum_empty = accum_full + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(accum_empty + 1);

    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t rank = NCTA == 1 ? 0u : cluster_ctarank();
    const bool rank0 = rank == 0;
    const uint32_t unit = NCTA == 1 ? blockIdx.x : blockIdx.x >> 1;      // CTA or pair: what the plan hands tiles to
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full + s, kBuilders / 32 + 1 + ((NCTA == 2 && rank0) ? 1 : 0));
            mbar_init(empty + s, kIssuers);
        }
        mbar_init(accum_full, kIssuers);
        mbar_init(accum_empty, NCTA * kEpilogue / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if constexpr (NCTA == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (NCTA == 2) cluster_sync_all();             // the barriers and the allocation of the other CTA exist
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    SegIter seg(pl, unit);
    uint32_t tile, kb, ke;
    const uint32_t F = pl.flushChunks;

    if (warp == 0) {
        // ---------------------------------------------------------------------------- B producer (this CTA's part of every chunk)
        if (lane == 0) {
            // chunk u (counted over the whole CTA) lands in buffer u % 3 and completes barrier full[u % kStages]; the buffer is
            // free once the MMAs of chunk u - 3 are done = phase of empty[(u - 3) % kStages]
            uint32_t u = 0, sb = 0;
            while (seg.next(tile, kb, ke)) {
                const uint32_t ct = tile % pl.tilesN;
                const unsigned char *src = Bt + ((size_t)ct * pl.chunks + kb) * kBBytes + rank * CF::kBStage;
                for (uint32_t k = kb; k < ke; ++k, ++u, src += kBBytes) {
                    // the records of the next chunks are pulled into L2 well ahead of the copy into shared memory
                    if (k + kPrefetchChunks < pl.chunks)
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + (size_t)kPrefetchChunks * kBBytes), "r"(CF::kBStage) : "memory");
                    if (u >= (uint32_t)CF::kBStages) {
                        const uint32_t w = u - CF::kBStages;
                        mbar_wait(empty + w % kStages, (w / kStages) & 1);
                    }
                    mbar_expect_tx(full + u % kStages, CF::kBStage);
                    bulk_g2s(stB + sb * CF::kBStage, src, CF::kBStage, full + u % kStages);
                    if (++sb == (uint32_t)CF::kBStages) sb = 0;
                }
            }
        }
    } else if (warp == 1 || warp == kThreads / 32 - 1) {
        const uint32_t m = warp == 1 ? 0u : 1u;               // the M block (accumulator) this issuer owns
        if (lane == 0 && rank == 0) {
            // ------------------------------------------------------------------------ MMA issuer of M block m
            constexpr uint32_t idT = make_idesc(2, 128 * NCTA, kTileN), idH = make_idesc(1, 128 * NCTA, kTileN);
            const uint32_t d = tmem_base + m * kTileN;
            uint32_t sa = 0, roundA = 0, sb = 0, pieces = 0;
            bool ready = false;                               // the operands of the chunk about to be issued were seen complete already
#ifdef CUSPMM_TC_DEBUG
            long long dbgW = 0, dbgD = 0, dbgN = 0, dbgR = 0;
            const long long dbgT0 = clock64();
#endif
            while (seg.next(tile, kb, ke)) {
                for (uint32_t k = kb; k < ke; ++k) {
                    const bool first = (k - kb) % F == 0;     // first chunk of a piece: the accumulator starts over
#ifdef CUSPMM_TC_DEBUG
                    const long long t0 = clock64();
#endif
                    if (first && pieces > 0) {
                        if constexpr (NCTA == 1) mbar_wait(accum_empty, (pieces - 1) & 1); else mbar_wait_cluster(accum_empty, (pieces - 1) & 1);
                        tc_fence_after();
                    }
#ifdef CUSPMM_TC_DEBUG
                    const long long t1 = clock64();
#endif
                    if (!ready) {
                        if constexpr (NCTA == 1) mbar_wait(full + sa, roundA & 1); else mbar_wait_cluster(full + sa, roundA & 1);
                    }
#ifdef CUSPMM_TC_DEBUG
                    dbgD += t1 - t0; dbgW += clock64() - t1; ++dbgN; dbgR += ready;
#endif
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(stA + sa * kABytes) + m * 8192, b0 = smem_u32(stB + sb * CF::kBStage);
                    // Operand tiles hold 4 k-groups of 4 k: a tf32 MMA (K = 8) and a pair MMA (K = 16 = 8 k x 2) both take two of
                    // them.  A (M = 128 rows per CTA): next k group 2048 B; B (N = 256 columns, 128 per CTA of a pair): kBLbo; next 8
                    // rows / columns 128 B
                    umma_tf32<NCTA>(d, make_desc(a0 + kOffT, 2048, 128), make_desc(b0, CF::kBLbo, 128), idT, first ? 0u : 1u);
                    umma_tf32<NCTA>(d, make_desc(a0 + kOffT + 4096, 2048, 128), make_desc(b0 + 2 * CF::kBLbo, CF::kBLbo, 128), idT, 1u);
                    // bf16(a) x bf16(r_b) + bf16(r_a) x bf16(b): k 0..7, k 8..15
                    umma_bf16<NCTA>(d, make_desc(a0 + kOffP, 2048, 128), make_desc(b0 + CF::kBOffP, CF::kBLbo, 128), idH);
                    umma_bf16<NCTA>(d, make_desc(a0 + kOffP + 4096, 2048, 128), make_desc(b0 + CF::kBOffP + 2 * CF::kBLbo, CF::kBLbo, 128), idH);
                    umma_commit<NCTA>(empty + sa);            // (with the other issuer's commit) the stage may be refilled
                    if (++sa == kStages) { sa = 0; ++roundA; }
                    if (++sb == (uint32_t)CF::kBStages) sb = 0;
                    if ((k + 1 - kb) % F == 0 || k + 1 == ke) { umma_commit<NCTA>(accum_full); ++pieces; }
                    // The tensor pipe is still busy with what was just queued: look at the barrier of the NEXT chunk now (one
                    // non-blocking test), so that its latency is not paid after the queue has run dry.
                    if constexpr (NCTA == 1) ready = mbar_test(full + sa, roundA & 1); else ready = mbar_test_cluster(full + sa, roundA & 1);
                }
            }
#ifdef CUSPMM_TC_DEBUG
            if (unit == 0 || unit == 37)
                printf("unit %u issuer %u: %lld chunks (%lld ready early), %lld clk per chunk; waited per chunk: operands %lld, drain %lld\n",
                       unit, m, dbgN, dbgR, (clock64() - dbgT0) / dbgN, dbgW / dbgN, dbgD / dbgN);
#endif
        } else if (NCTA == 2 && lane == 0 && m == 0) {
            // ------------------------------------------------------------------------ rank 1: tell rank 0 when this half is ready
            uint32_t sa = 0, roundA = 0;
            while (seg.next(tile, kb, ke)) {
                for (uint32_t k = kb; k < ke; ++k) {
                    mbar_wait(full + sa, roundA & 1);
                    mbar_arrive_remote_relaxed(full + sa, 0);
                    if (++sa == kStages) { sa = 0; ++roundA; }
                }
            }
        }
    } else if (warp < 2 + kBuilders / 32) {
        // ---------------------------------------------------------------------------- builders
        // Two adjacent lanes share a row: lane half h takes the entries of the row whose index has parity h.  The scatter loop of a
        // warp is as long as its longest lane; with one thread per row that was ~5.4 iterations per chunk at 10 % density for 1.6
        // entries on average, and the chunk's operands are complete only when the slowest warp is.
        const uint32_t bt = threadIdx.x - 64;
        const uint32_t t = LPR == 2 ? bt >> 1 : bt, h = LPR == 2 ? bt & 1u : 0u;     // t: row this thread builds (M block t / 128, row t % 128)
        const uint32_t mblk = t >> 7, rowInBlk = t & 127;
        const uint32_t offT = kOffT + mblk * 8192 + rowInBlk * 16;
        const uint32_t rowInTile = NCTA == 1 ? t : mblk * 256 + rank * 128 + rowInBlk;
        const uint32_t ring = smem_u32(smem + CF::kRingOff) + t * 16u;   // this row's slot 0 of the colIdxs ring; vals 16 KB further
        const uint32_t stA32 = smem_u32(stA);
        uint32_t s = 0, round = 0;
#ifdef CUSPMM_TC_DEBUG
        long long dbgRefill = 0, dbgWaitE = 0, dbgZero = 0, dbgLoop = 0, dbgFence = 0, dbgCh = 0;
#endif
        while (seg.next(tile, kb, ke)) {
            const uint32_t r = (tile / pl.tilesN) * CF::kTileM + rowInTile;
            // cursor space: CSR -- positions in colIdxs / vals (entry p at address p); sliced ELL -- entries of the row (entry p at
            // address ebase + 32 p)
            uint32_t p0 = 0, end = 0, ebase = 0;
            if (r < M) {
                if constexpr (SELL) {
                    const uint32_t sb = __ldg(rowPtrs + (r >> 5));
                    ebase = sb + (r & 31u);
                    end = (__ldg(rowPtrs + (r >> 5) + 1) - sb) >> 5;
                } else {
                    p0 = __ldg(rowPtrs + r);
                    end = __ldg(rowPtrs + r + 1);
                }
                if (kb > 0) {                                 // first entry of the row at or after column 16 kb (padding: 0xFFFFFFFF)
                    const uint32_t target = kb * kKC;
                    uint32_t lo = p0, hi = end;
                    while (lo < hi) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (__ldg(colIdxs + (SELL ? ebase + 32u * mid : mid)) < target) lo = mid + 1; else hi = mid;
                    }
                    p0 = lo;
                }
            }
            uint32_t p = LPR == 2 ? p0 + ((h - p0) & 1u) : p0;   // this lane's next entry
            // Entries reach the lanes through a ring in shared memory: 4 slots of 4 entries (16 B of colIdxs + 16 B of vals) per
            // row, filled with cp.async by the even lane.  (A register FIFO filled with ordinary loads does not work here: the
            // scoreboard is per warp and register, so a lane shifting its FIFO waits for the load another lane issued a moment ago
            // -- ncu on the first version: 29 % of all samples on that move, one exposed memory latency per chunk.)  Refills are
            // issued at the start of a chunk, one commit group per chunk, and cp.async.wait_group 2 (+ __syncwarp for the odd
            // lane) makes the groups older than two chunks visible: a block is requested 12..16 entries before it is read.
            // Entry p of the row lives at ring + (p / 4 % 4) * 4096 + (p % 4) * 4.  Both lanes track fb / lb identically.
            uint32_t fb = p0 & ~3u;                           // next block to request
            uint32_t lb = fb;                                 // entries below lb have landed
            auto refill = [&](uint32_t plow) {                // plow: the lower of the two lanes' next entries
                uint32_t lim = (plow & ~3u) + 4u * CF::kRingBlocks;
                if (lim > end) lim = end;
                while (fb < lim) {
                    if (h == 0) {
                        const uint32_t dst = ring + ((fb & CF::kRingMask) << 10);
                        if constexpr (SELL) {
#pragma unroll
                            for (uint32_t e = 0; e < 4; ++e) {
                                const bool in = fb + e < end;
                                const size_t at = (size_t)ebase + 32u * (size_t)(in ? fb + e : fb);
                                cp_async4(dst + 4 * e, colIdxs + at, in ? 4u : 0u);
                                cp_async4(dst + CF::kRingVals + 4 * e, vals + at, in ? 4u : 0u);
                            }
                        } else if (VEC) {
                            const uint32_t valid = (nnzTotal - fb < 4u ? nnzTotal - fb : 4u) * 4u;
                            cp_async16(dst, colIdxs + fb, valid);
                            cp_async16(dst + CF::kRingVals, vals + fb, valid);
                        } else {
#pragma unroll
                            for (uint32_t e = 0; e < 4; ++e) {
                                const bool in = fb + e < nnzTotal;
                                cp_async4(dst + 4 * e, colIdxs + (in ? fb + e : fb), in ? 4u : 0u);
                                cp_async4(dst + CF::kRingVals + 4 * e, vals + (in ? fb + e : fb), in ? 4u : 0u);
                            }
                        }
                    }
                    fb += 4;
                }
            };
            auto lower_of_pair = [&]() {
                if constexpr (LPR == 1) return p;
                const uint32_t o = __shfl_xor_sync(0xFFFFFFFFu, p, 1);
                return o < p ? o : p;
            };

            // (Two chunks per builder pass -- amortising the chain barrier wake-up -> clears -> LDS -> stores -> proxy fence -> arrive
            //  over 32 columns -- was measured and is slower from 10 % density: 2.61 vs 2.52 ms, 7.5 vs 6.4 ms at 50 %.  Also dropped:
            //  keeping the (column, value) of the next entry in registers across chunks and issuing its ring loads before the current
            //  entry's stores: 2.65 vs 2.50 ms at 10 %.)
            uint32_t fbPrev = fb;                             // fb before the refill of the previous chunk
            for (uint32_t k = kb; k < ke; ++k) {
#ifdef CUSPMM_TC_DEBUG
                const long long tb0 = clock64();
#endif
                const uint32_t fbBefore = fb;
                refill(lower_of_pair());
                asm volatile("cp.async.commit_group;" ::: "memory");
                if (k == kb) { asm volatile("cp.async.wait_group 0;" ::: "memory"); lb = fb; }
                else {                                        // all groups but the last two have landed: requested >= 2 chunks ago
                    asm volatile("cp.async.wait_group 2;" ::: "memory");
                    if (fbPrev > lb) lb = fbPrev;
                }
                if constexpr (LPR == 2) __syncwarp();         // the even lane's copies are visible to the odd lane
                fbPrev = fbBefore;
#ifdef CUSPMM_TC_DEBUG
                const long long tb1 = clock64();
#endif
                if (round > 0) mbar_wait(empty + s, (round - 1) & 1);
#ifdef CUSPMM_TC_DEBUG
                const long long tb2 = clock64();
#endif
                const uint32_t aT = stA32 + s * kABytes + offT;
                if constexpr (LPR == 2) {                     // each lane clears half of the row: k groups 2h, 2h + 1 of both operand tiles
                    sts_zero16(aT + (2 * h) * 2048); sts_zero16(aT + (2 * h + 1) * 2048);
                    sts_zero16(aT + kOffP + (2 * h) * 2048); sts_zero16(aT + kOffP + (2 * h + 1) * 2048);
                    __syncwarp();                             // ... before the partner may store into them
                } else {
#pragma unroll
                    for (int g = 0; g < 4; ++g) { sts_zero16(aT + g * 2048); sts_zero16(aT + kOffP + g * 2048); }
                }
#ifdef CUSPMM_TC_DEBUG
                const long long tb3 = clock64();
#endif
                const uint32_t k0 = k * kKC;
                for (;;) {
                    const uint32_t stop = end < lb ? end : lb;
                    bool done = false;
                    // Two entries per round, all four ring loads issued before the first is used: an iteration is a chain of
                    // shared-memory round trips (~100 clocks each with the pipe loaded by the MMAs: clock64 showed ~220 clocks per
                    // entry, 1200 of the 1900 clocks a chunk takes at 10 % density), not instructions.  (10 % dense: 2.37 ms with one
                    // entry per round, 2.23 with two, 2.43 with four (the loads past the chunk are wasted).)
                    auto place = [&](uint32_t kk, uint32_t bits) {
                        const float v = __uint_as_float(bits);
                        uint32_t tb = (bits + 0x1000u) & 0xFFFFE000u;          // tf32, round to nearest (ties away)
                        float res = v - __uint_as_float(tb);                   // exact
                        uint32_t pk;                                            // low half bf16(v), high half bf16(res)
                        if ((bits << 1) >= 0xFE000000u) {                      // |v| >= 2^127, Inf, NaN
                            if ((bits << 1) < 0xFF000000u) {                   // finite: truncate instead of rounding up to infinity
                                tb = bits & 0xFFFFE000u;
                                res = v - __uint_as_float(tb);
                                pk = (bits >> 16) | ((uint32_t)bf16_bits(res) << 16);
                            } else {                                            // Inf / NaN: main product only (Inf - Inf in the corrections
                                tb = (bits & 0x007FFFFFu) ? (bits | 0x00400000u) & 0xFFFFE000u : bits;   //  would turn Inf into NaN)
                                pk = 0u;
                            }
                        } else {
                            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(res), "f"(v));
                        }
                        // tf32 tile and pair tile have the same geometry (4 k per 16-byte core-matrix row), 16 KB apart
                        const uint32_t dstT = aT + ((kk & 12u) << 9) + ((kk & 3u) << 2);
                        sts_u32(dstT, tb);
                        sts_u32(dstT + kOffP, pk);
                    };
                    while (p < stop) {
                        const uint32_t q1 = p + LPR;
                        const bool two = q1 < stop;
                        const uint32_t ea0 = ring + ((p & CF::kRingMask) << 10) + ((p & 3u) << 2);
                        const uint32_t ea1 = ring + ((q1 & CF::kRingMask) << 10) + ((q1 & 3u) << 2);
                        const uint32_t c0 = lds_u32(ea0), b0 = lds_u32(ea0 + CF::kRingVals);
                        uint32_t c1 = 0xFFFFFFFFu, b1 = 0u;
                        if (two) { c1 = lds_u32(ea1); b1 = lds_u32(ea1 + CF::kRingVals); }
                        const uint32_t kk0 = c0 - k0;                            // columns ascend and are >= k0 here
                        if (kk0 >= (uint32_t)kKC) { done = true; break; }
                        place(kk0, b0);
                        p = q1;
                        if (!two) break;                                          // (p >= stop: the outer logic decides what follows)
                        const uint32_t kk1 = c1 - k0;
                        if (kk1 >= (uint32_t)kKC) { done = true; break; }
                        place(kk1, b1);
                        p += LPR;
                    }
                    // a lane whose row has more entries that are not known to have landed (a chunk that used more than ~2 blocks)
                    // needs the whole warp: the even lanes own the copies
                    const bool need = !done && p < end;
                    if constexpr (LPR == 2) { if (!__any_sync(0xFFFFFFFFu, need)) break; }
                    else { if (!need) break; }
                    refill(lower_of_pair());
                    asm volatile("cp.async.commit_group;" ::: "memory");
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    if constexpr (LPR == 2) __syncwarp();
                    lb = fb;
                }
#ifdef CUSPMM_TC_DEBUG
                const long long tb4 = clock64();
#endif
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> tensor-core (async proxy) reads
                __syncwarp();
                if (lane == 0) mbar_arrive(full + s);
#ifdef CUSPMM_TC_DEBUG
                { const long long tb5 = clock64(); dbgRefill += tb1 - tb0; dbgWaitE += tb2 - tb1; dbgZero += tb3 - tb2; dbgLoop += tb4 - tb3; dbgFence += tb5 - tb4; ++dbgCh; }
#endif
                if (++s == kAStages) { s = 0; ++round; }
            }
        }
#ifdef CUSPMM_TC_DEBUG
        if (unit == 0 && rank == 0 && lane == 0 && (warp == 2 || warp == 5) && dbgCh)
            printf("builder warp %u: per chunk: refill+cp.async wait %lld, wait for the stage %lld, clear %lld, scatter loop %lld, fence+arrive %lld\n",
                   warp, dbgRefill / dbgCh, dbgWaitE / dbgCh, dbgZero / dbgCh, dbgLoop / dbgCh, dbgFence / dbgCh);
#endif
    } else {     // (the issuer warps were taken by the branch above)
        // ---------------------------------------------------------------------------- epilogue (warps 18..21)
        // TMEM lanes [32 q, 32 q + 32) are the ones this warp may read; it drains both M blocks of those lanes.  Every piece is
        // added to C (zeroed by the launcher) with red.global.add.v4.f32; a tile that is one single piece is stored.
        // (Load-add-store by the CTA that owns a tile was 3x slower than red.add: the C tile does not stay in L2 between two
        //  drains and the loads are serialised behind the TMEM reads -- 59 000 clocks per drain against 18 000.)
        const uint32_t q = warp & 3;
        uint32_t pieces = 0;
        if (NCTA == 2 && tmaDrain) {
            // Drain through the TMA: 32 rows x 32 columns of an accumulator go TMEM -> registers -> a 4 KB staging tile of this warp
            // (128-byte swizzle: lane = row writes its eight 16-byte pieces to piece ^ (row % 8), no bank conflicts) -> ONE
            // cp.reduce.async.bulk.tensor (.add, fp32: round-to-nearest adds in L2, rows / columns past the end of C clipped by the
            // tensor map).  The red.add path below pushes every 16 bytes through the LSU (~1.3 clocks per lane: ~20 000 clocks per
            // drain, a fifth of the kernel); here the LSU only sees the shared-memory stores.
            const uint32_t stage = smem_u32(smem + CF::kOutOff) + (warp - (2 + kBuilders / 32)) * 4096u;
            const uint32_t myrow = stage + lane * 128u;
            while (seg.next(tile, kb, ke)) {
                const uint32_t rt = tile / pl.tilesN, ct = tile % pl.tilesN;
                for (uint32_t k = kb; k < ke; k += F, ++pieces) {
                    mbar_wait(accum_full, pieces & 1);
                    tc_fence_after();
#pragma unroll 1
                    for (uint32_t em = 0; em < 2; ++em) {
                        const uint32_t row0 = rt * CF::kTileM + em * 256 + rank * 128 + q * 32;
#pragma unroll 1
                        for (uint32_t cb = 0; cb < kTileN / 32; ++cb) {
                            uint32_t v[32];
                            const uint32_t taddr = tmem_base + ((q * 32u) << 16) + em * kTileN + cb * 32;
                            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                                         "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                                         : "r"(taddr));
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                            // the previous reduction of this warp has read the staging tile
                            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                            __syncwarp();
#pragma unroll
                            for (uint32_t j = 0; j < 8; ++j)
                                asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};"
                                             ::"r"(myrow + ((j ^ (lane & 7u)) << 4)), "r"(v[4 * j]), "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3]) : "memory");
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                            __syncwarp();
                            if (lane == 0 && row0 < M && ct * kTileN + cb * 32 < N) {
                                asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];"
                                             ::"l"(&tmapC), "r"(ct * kTileN + cb * 32), "r"(row0), "r"(stage) : "memory");
                                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                            }
                        }
                    }
                    // the accumulators have been read: the next piece may start (the reductions complete on their own)
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (rank == 0) mbar_arrive(accum_empty); else mbar_arrive_remote(accum_empty, 0);
                    }
                }
            }
            if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // all of this warp's reductions have been performed
        } else
        while (seg.next(tile, kb, ke)) {
            const uint32_t rt = tile / pl.tilesN, ct = tile % pl.tilesN;
            const bool direct = (kb == 0 && ke == pl.chunks && pl.chunks <= F);     // the only piece of the tile: plain stores
            const uint32_t ncols = N - ct * kTileN < (uint32_t)kTileN ? N - ct * kTileN : (uint32_t)kTileN;
            for (uint32_t k = kb; k < ke; k += F, ++pieces) {
                mbar_wait(accum_full, pieces & 1);
                tc_fence_after();
#pragma unroll 1
                for (uint32_t em = 0; em < 2; ++em) {
                    const uint32_t er = rt * CF::kTileM + (NCTA == 1 ? em * 128 : em * 256 + rank * 128) + q * 32 + lane;
                    float *crow = C + (size_t)er * ldc + (size_t)ct * kTileN;
#pragma unroll 1
                    for (uint32_t cb = 0; cb < kTileN / 32; ++cb) {
                        uint32_t v[32];
                        const uint32_t taddr = tmem_base + ((q * 32u) << 16) + em * kTileN + cb * 32;
                        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                                     "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                                       "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                                       "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                                     : "r"(taddr));
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (er < M) {
                            const uint32_t c0 = cb * 32;
                            if (vecC) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    if (c0 + 4 * j + 4 <= ncols) {
                                        float *dst = crow + c0 + 4 * j;
                                        const float x = __uint_as_float(v[4 * j]), y = __uint_as_float(v[4 * j + 1]),
                                                    z2 = __uint_as_float(v[4 * j + 2]), w = __uint_as_float(v[4 * j + 3]);
                                        if (direct) __stcs(reinterpret_cast<float4 *>(dst), make_float4(x, y, z2, w));
                                        else red_add_v4(dst, x, y, z2, w);
                                    }
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    if (c0 + j < ncols) {
                                        if (direct) crow[c0 + j] = __uint_as_float(v[j]);
                                        else atomicAdd(crow + c0 + j, __uint_as_float(v[j]));
                                    }
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (NCTA == 1 || rank == 0) mbar_arrive(accum_empty); else mbar_arrive_remote(accum_empty, 0);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (NCTA == 2) cluster_sync_all();             // nobody frees tensor memory (or leaves) while the other CTA may still use it
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        if constexpr (NCTA == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

} // namespace csrtc

// The kernel that computes C when B holds non-finite values (it exits at once otherwise): plain fp32, a warp per row, lanes over
// the columns.  Only stored entries are multiplied, as in the reference; speed does not matter here.
template <bool SELL>
__global__ void __launch_bounds__(256)
csr_tc_fallback_kernel(const uint32_t *__restrict__ rowPtrs, const uint32_t *__restrict__ colIdxs, const float *__restrict__ vals,
                       uint32_t M, const float *__restrict__ B, uint32_t N, size_t ldb, float *__restrict__ C, size_t ldc,
                       const uint32_t *__restrict__ onlyIf) {
    if (!*onlyIf) return;
    const uint32_t warpsPerGrid = gridDim.x * (blockDim.x >> 5);
    for (uint32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < M; r += warpsPerGrid) {
        // CSR: entries p0 .. p1 - 1; sliced ELL: entry j at first + 32 j, the row ends at its first padding entry
        size_t first;
        uint32_t len, stride;
        if constexpr (SELL) {
            const uint32_t sb = rowPtrs[r >> 5];
            first = (size_t)sb + (r & 31u);
            len = (rowPtrs[(r >> 5) + 1] - sb) >> 5;
            stride = 32;
        } else {
            first = rowPtrs[r];
            len = rowPtrs[r + 1] - rowPtrs[r];
            stride = 1;
        }
        for (uint32_t n0 = 0; n0 < N; n0 += 128) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (uint32_t j = 0; j < len; ++j) {
                const uint32_t c = colIdxs[first + (size_t)j * stride];
                if (SELL && c == kPad) break;
                const float a = vals[first + (size_t)j * stride];
                const float *brow = B + (size_t)c * ldb + n0 + lane_id();
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (n0 + 32 * u + lane_id() < N) acc[u] = fmaf(a, brow[32 * u], acc[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (n0 + 32 * u + lane_id() < N) C[(size_t)r * ldc + n0 + 32 * u + lane_id()] = acc[u];
        }
    }
}

static std::mutex g_tc_pool_mu;
static cudaMemPool_t g_tc_pools[64] = {};
// gives the memory the pools hold back to the device (cuspmm_host_pipeline_release); device < 0: all
void tc_pool_trim(int dev) {
    std::lock_guard<std::mutex> lock(g_tc_pool_mu);
    for (int d = 0; d < 64; ++d)
        if (g_tc_pools[d] && (dev < 0 || dev == d)) cudaMemPoolTrimTo(g_tc_pools[d], 0);
}
static cudaMemPool_t tc_pool(int dev) {
    std::mutex &mu = g_tc_pool_mu;
    cudaMemPool_t *pools = g_tc_pools;
    std::lock_guard<std::mutex> lock(mu);
    if (dev < 0 || dev >= 64) return nullptr;
    if (!pools[dev]) {
        cudaMemPoolProps props;
        memset(&props, 0, sizeof props);
        props.allocType = cudaMemAllocationTypePinned;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        if (cudaMemPoolCreate(&pools[dev], &props) != cudaSuccess) { pools[dev] = nullptr; return nullptr; }
        uint64_t keep = ~0ull;
        cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
    }
    return pools[dev];
}

// Does the tensor-core kernel (work ~ Mpad x K x Npad, independent of nnz) beat the fp32 kernels (work ~ nnz x N)?
// r = nnz x N / (Mpad x K x Npad) is the useful fraction of the dense work.  Measured cross-over (profiles/r02_tc_sweep.jsonl,
// 117 shapes: squares 4096..25605, N 128..2048, 2..50 % dense, row panels of the BASELINE matrix, FFN shapes; final kernel):
// r ~ 0.036 at N = 512, ~0.028 for N <= 256 (where the fp32 staged kernel does not apply) and for N >= 1024; small problems pay the
// fixed cost of the launches and the partly filled last wave: x (1 + 3.7e9 / dense work).  With these thresholds the choice is
// within 2 % of the faster kernel on 115 of the 117 shapes.  Above the cross-over the gain grows quickly: 1.2-3.0x at 10 %,
// 2-4.9x at 20 %, 3-8.9x at 50 %.
// On the sliced-ELL layout (nnz = slots) the slices are compacted into CSR first (two passes over A: +0.35 ms at 10 %, +1.8 ms at
// 50 % on 25605^2): 2.58 against 4.31 ms (fp32 ELL kernels) at 10 %, 4.09 / 9.82 at 30 %, 5.65 / 15.5 at 50 %; threshold 1.25x higher.
bool tc_kernel_wins(uint32_t M, uint32_t K, uint64_t nnz, uint32_t N, bool sell) {
    if (M == 0 || K == 0 || N == 0) return false;
    const double Mpad = (double)((M + 511u) / 512u) * 512.0, Npad = (double)((N + 255u) / 256u) * 256.0;
    const double dense = Mpad * (double)K * Npad;
    const double r = (double)nnz * (double)N / dense;
    const double base = N <= 256 ? 0.028 : (N <= 512 ? 0.036 : 0.029);
    return r >= (sell ? 1.25 : 1.0) * base * (1.0 + 3.7e9 / dense);
}

size_t csr_tc_workspace_bytes(uint32_t K, uint32_t N) {
    const uint64_t tilesN = (N + csrtc::kTileN - 1) / csrtc::kTileN, chunks = (K + csrtc::kKC - 1) / csrtc::kKC;
    return (size_t)(tilesN * chunks * csrtc::kBBytes + 256);
}

template <int NCTA, bool SELL>
static int run_tc(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals, uint32_t M, uint32_t K, uint32_t nnz,
                  const float *B, uint32_t N, size_t ldb, float *C, size_t ldc, cudaMemPool_t pool, cudaStream_t st) {
    using namespace csrtc;
    using CF = Cfg<NCTA>;
    Plan pl;
    pl.tilesN = (N + kTileN - 1) / kTileN;
    pl.chunks = (K + kKC - 1) / kKC;
    const uint32_t tilesM = (M + CF::kTileM - 1) / CF::kTileM;
    const uint64_t tiles = (uint64_t)tilesM * pl.tilesN;
    const uint64_t units = tiles * pl.chunks;
    // one CTA (or pair) per SM (or two); with little work, at least 8 chunks each
    uint64_t grid = (uint64_t)sm_count() / NCTA;
    if (units / 8 < grid) grid = units / 8 ? units / 8 : 1;
    pl.grid = (uint32_t)grid;
    pl.fullWaves = (uint32_t)(tiles / grid);
    pl.remTiles = (uint32_t)(tiles - (uint64_t)pl.fullWaves * grid);
    pl.unitsPerCta = ((uint64_t)pl.remTiles * pl.chunks + grid - 1) / grid;
    static const int flushEnv = getenv("CUSPMM_TC_FLUSH") ? atoi(getenv("CUSPMM_TC_FLUSH")) : 0;      // tuning hook
    pl.flushChunks = flushEnv > 0 ? (uint32_t)flushEnv : kFlushChunks;

    const size_t bytes = csr_tc_workspace_bytes(K, N);
    unsigned char *ws = nullptr;
    CUSPMM_CUDA(cudaMallocFromPoolAsync(reinterpret_cast<void **>(&ws), bytes, pool, st));
    uint32_t *flag = reinterpret_cast<uint32_t *>(ws + bytes - 256);
    int rc = CUSPMM_OK;
    do {
        if (cudaMemsetAsync(flag, 0, sizeof(uint32_t), st) != cudaSuccess) { rc = set_error(CUSPMM_ERR_CUDA, "memset of the flag failed"); break; }
        csr_tc_prepare_B<NCTA><<<dim3(2 * pl.chunks, pl.tilesN), 256, 0, st>>>(B, K, N, ldb, ws, pl.chunks, flag);
        if (cudaGetLastError() != cudaSuccess) { rc = set_error(CUSPMM_ERR_CUDA, "launch of csr_tc_prepare_B failed"); break; }
        count_launch();
        // accumulator drains through TMA reductions (pairs; C 16-byte aligned with a 16-byte multiple as row pitch): a tensor map of
        // C with 32 x 32 boxes and the 128-byte swizzle the epilogue writes its staging tiles in
        CUtensorMap tmapC;
        memset(&tmapC, 0, sizeof tmapC);
        static const int tmaEnv = getenv("CUSPMM_TC_TMA_DRAIN") ? atoi(getenv("CUSPMM_TC_TMA_DRAIN")) : 1;   // tuning hook
        int tmaDrain = 0;
        if (NCTA == 2 && tmaEnv && (reinterpret_cast<uintptr_t>(C) & 15) == 0 && (ldc * sizeof(float)) % 16 == 0) {
            tmemk::EncodeTiledFn encode = tmemk::tensor_map_encoder();
            const cuuint64_t dims[2] = {N, M};
            const cuuint64_t strides[1] = {(cuuint64_t)ldc * sizeof(float)};
            const cuuint32_t box[2] = {32, 32};
            const cuuint32_t estr[2] = {1, 1};
            if (encode && encode(&tmapC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, C, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
                tmaDrain = 1;
        }
        // every tile is drained into C piece by piece with reductions (on the red.add path a tile that is a single piece is stored):
        // C starts at zero
        const bool allDirect = !tmaDrain && pl.chunks <= pl.flushChunks && (pl.remTiles == 0 || pl.unitsPerCta % pl.chunks == 0);
        if (!allDirect && cudaMemset2DAsync(C, ldc * sizeof(float), 0, (size_t)N * sizeof(float), M, st) != cudaSuccess) {
            rc = set_error(CUSPMM_ERR_CUDA, "memset of C failed");
            break;
        }
        const bool vecA = !SELL && ((reinterpret_cast<uintptr_t>(colIdxs) | reinterpret_cast<uintptr_t>(vals)) & 15) == 0;
        const int vecC = (N % 4 == 0) && (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
        // builder threads per row: two lanes per row pay from ~12 % density (25605^2 x 512: 50 %: 6.48 -> 4.79 ms, 10 %: 2.50 = 2.50,
        // 5 %: 2.16 -> 2.26, 2 %: 1.89 -> 2.10: below that the second set of warps only adds fixed work per chunk)
        static const int lprEnv = getenv("CUSPMM_TC_LPR") ? atoi(getenv("CUSPMM_TC_LPR")) : 0;          // tuning hook
        const int lpr = lprEnv == 1 || lprEnv == 2 ? lprEnv : ((double)nnz >= 0.12 * (double)M * (double)K ? 2 : 1);
        auto kern = lpr == 2 ? (vecA ? csr_tc_kernel<!SELL, NCTA, 2, SELL> : csr_tc_kernel<false, NCTA, 2, SELL>)
                             : (vecA ? csr_tc_kernel<!SELL, NCTA, 1, SELL> : csr_tc_kernel<false, NCTA, 1, SELL>);
        const uint32_t smemBytes = lpr == 2 ? Cfg<NCTA, 2>::kSmemTotal : Cfg<NCTA, 1>::kSmemTotal;
        if (set_smem_once(kern, smemBytes) != cudaSuccess) { rc = set_error(CUSPMM_ERR_CUDA, "cannot reserve %u bytes of shared memory", smemBytes); break; }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(pl.grid * NCTA, 1, 1);
        cfg.blockDim = dim3(threads_for(lpr), 1, 1);
        cfg.dynamicSmemBytes = smemBytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = NCTA;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const unsigned char *wsc = ws;
        const uint32_t *flagc = flag;
        if (cudaLaunchKernelEx(&cfg, kern, rowPtrs, colIdxs, vals, M, nnz, wsc, N, C, ldc, pl, flagc, vecC, tmapC, tmaDrain) != cudaSuccess) {
            rc = set_error(CUSPMM_ERR_CUDA, "launch of csr_tc_kernel failed: %s", cudaGetErrorString(cudaPeekAtLastError()));
            break;
        }
        count_launch();
        csr_tc_fallback_kernel<SELL><<<(unsigned)sm_count() * 4, 256, 0, st>>>(rowPtrs, colIdxs, vals, M, B, N, ldb, C, ldc, flag);
        if (cudaGetLastError() != cudaSuccess) { rc = set_error(CUSPMM_ERR_CUDA, "launch of csr_tc_fallback_kernel failed"); break; }
        count_launch();
    } while (0);
    cudaFreeAsync(ws, st);
    return rc;
}

// ---------------------------------------------------------------------------------------------- sliced ELL -> CSR on the device
// The tensor kernel reads sliced ELL directly (SELL = true), but a row's entries are 128 bytes apart there and the per-row ring of
// 4-byte copies fetches a 32-byte sector for each of them: 3.4 ms against 2.2 on CSR at 10 % density, 14.1 against 3.85 at 50 %.
// Compacting the slices into CSR first (two passes over A at HBM speed, tiles transposed through shared memory so that both the
// reads and the writes are 128-byte lines) and running the CSR kernel is faster at every density where the tensor kernel is chosen.
__global__ void __launch_bounds__(256)
sell_row_lengths_kernel(const uint32_t *__restrict__ slicePtrs, const uint32_t *__restrict__ colIdxs, uint32_t M, uint32_t *__restrict__ lens) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > M) return;
    if (r == M) { lens[M] = 0; return; }
    const uint32_t sb = slicePtrs[r >> 5], w = (slicePtrs[(r >> 5) + 1] - sb) >> 5;
    const size_t first = (size_t)sb + (r & 31u);
    uint32_t lo = 0, hi = w;                                  // first padding entry (columns ascend, padding = 0xFFFFFFFF)
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (colIdxs[first + 32u * (size_t)mid] != kPad) lo = mid + 1; else hi = mid;
    }
    lens[r] = lo;
}
// one CTA per slice, 4 warps; a warp moves tiles of 32 slots x 32 rows: coalesced reads along the rows of the slice, transposed in
// shared memory, coalesced writes along the entries of a row
__global__ void __launch_bounds__(128)
sell_compact_kernel(const uint32_t *__restrict__ slicePtrs, const uint32_t *__restrict__ colIdxs, const float *__restrict__ vals, uint32_t M,
                    const uint32_t *__restrict__ rowPtrs, uint32_t *__restrict__ outCols, float *__restrict__ outVals) {
    __shared__ uint32_t tc[4][32][33];
    __shared__ float tv[4][32][33];
    const uint32_t s = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t sb = slicePtrs[s], w = (slicePtrs[s + 1] - sb) >> 5;
    const uint32_t r = s * 32 + lane;
    const uint32_t myStart = r < M ? rowPtrs[r] : 0u, myLen = r < M ? rowPtrs[r + 1] - myStart : 0u;
    for (uint32_t j0 = warp * 32; j0 < w; j0 += 4 * 32) {
#pragma unroll 4
        for (uint32_t i = 0; i < 32; ++i) {
            const bool in = j0 + i < w;
            const size_t at = (size_t)sb + (size_t)(j0 + i) * 32u + lane;
            tc[warp][i][lane] = in ? colIdxs[at] : kPad;
            tv[warp][i][lane] = in ? vals[at] : 0.0f;
        }
        __syncwarp();
#pragma unroll 4
        for (uint32_t rr = 0; rr < 32; ++rr) {                // row rr of the slice: lane = entry j0 + lane
            const uint32_t start = __shfl_sync(0xFFFFFFFFu, myStart, rr), len = __shfl_sync(0xFFFFFFFFu, myLen, rr);
            if (j0 + lane < len) {
                outCols[(size_t)start + j0 + lane] = tc[warp][lane][rr];
                outVals[(size_t)start + j0 + lane] = tv[warp][lane][rr];
            }
        }
        __syncwarp();
    }
}

static int sell_to_csr_then_tc(const uint32_t *slicePtrs, const uint32_t *colIdxs, const float *vals, uint32_t M, uint32_t K, uint32_t slots,
                               const float *B, uint32_t N, size_t ldb, float *C, size_t ldc, cudaMemPool_t pool, cudaStream_t st) {
    // temporaries from the stream-ordered pool: row pointers, and column / value arrays sized by the slot count (an upper bound of
    // nnz that needs no host round trip)
    const size_t rpBytes = ((size_t)(M + 1) * 4 + 255) & ~(size_t)255, arrBytes = ((size_t)slots * 4 + 255) & ~(size_t)255;
    size_t scanBytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)(M + 1), st);
    scanBytes = (scanBytes + 255) & ~(size_t)255;
    unsigned char *tmp = nullptr;
    CUSPMM_CUDA(cudaMallocFromPoolAsync(reinterpret_cast<void **>(&tmp), 2 * rpBytes + 2 * arrBytes + scanBytes, pool, st));
    uint32_t *lens = reinterpret_cast<uint32_t *>(tmp), *rowPtrs = reinterpret_cast<uint32_t *>(tmp + rpBytes);
    uint32_t *cols = reinterpret_cast<uint32_t *>(tmp + 2 * rpBytes);
    float *v = reinterpret_cast<float *>(tmp + 2 * rpBytes + arrBytes);
    void *scanTmp = tmp + 2 * rpBytes + 2 * arrBytes;
    int rc = CUSPMM_OK;
    do {
        sell_row_lengths_kernel<<<(M + 1 + 255) / 256, 256, 0, st>>>(slicePtrs, colIdxs, M, lens);
        if (cudaGetLastError() != cudaSuccess) { rc = set_error(CUSPMM_ERR_CUDA, "launch of sell_row_lengths_kernel failed"); break; }
        count_launch();
        if (cub::DeviceScan::ExclusiveSum(scanTmp, scanBytes, lens, rowPtrs, (int)(M + 1), st) != cudaSuccess) {
            rc = set_error(CUSPMM_ERR_CUDA, "prefix sum of the row lengths failed");
            break;
        }
        sell_compact_kernel<<<(M + 31) / 32, 128, 0, st>>>(slicePtrs, colIdxs, vals, M, rowPtrs, cols, v);
        if (cudaGetLastError() != cudaSuccess) { rc = set_error(CUSPMM_ERR_CUDA, "launch of sell_compact_kernel failed"); break; }
        count_launch();
        rc = run_tc<2, false>(rowPtrs, cols, v, M, K, slots, B, N, ldb, C, ldc, pool, st);
    } while (0);
    cudaFreeAsync(tmp, st);
    return rc;
}

// variant 8 of the CSR kernels / variant 6 of the sliced-ELL kernels (sell: rowPtrs = slicePtrs, nnz = slots).  Rows must be sorted
// by column (as for variants 3, 5, 7).
int spmm_csr_tc(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals, uint32_t M, uint32_t K, uint32_t nnz,
                const float *B, uint32_t N, size_t ldb, float *C, size_t ldc, bool sell, cudaStream_t st) {
    if (K == 0) {
        CUSPMM_CUDA(cudaMemset2DAsync(C, ldc * sizeof(float), 0, (size_t)N * sizeof(float), M, st));
        return CUSPMM_OK;
    }
    int dev = 0;
    CUSPMM_CUDA(cudaGetDevice(&dev));
    cudaMemPool_t pool = tc_pool(dev);
    if (!pool) return set_error(CUSPMM_ERR_CUDA, "no memory pool for the tiled copy of B on device %d", dev);
    // tuning hook: CUSPMM_TC_PAIR=0 runs one CTA per 256-row tile (cta_group::1) instead of CTA pairs on 512-row tiles
    static const int pairEnv = getenv("CUSPMM_TC_PAIR") ? atoi(getenv("CUSPMM_TC_PAIR")) : 1;
    if (sell) {
        // tuning hook: CUSPMM_TC_SELL_DIRECT=1 lets the tensor kernel read the sliced layout itself instead of compacting it first
        static const int direct = getenv("CUSPMM_TC_SELL_DIRECT") ? atoi(getenv("CUSPMM_TC_SELL_DIRECT")) : 0;
        if (direct) return run_tc<2, true>(rowPtrs, colIdxs, vals, M, K, nnz, B, N, ldb, C, ldc, pool, st);
        return sell_to_csr_then_tc(rowPtrs, colIdxs, vals, M, K, nnz, B, N, ldb, C, ldc, pool, st);
    }
    if (pairEnv) return run_tc<2, false>(rowPtrs, colIdxs, vals, M, K, nnz, B, N, ldb, C, ldc, pool, st);
    return run_tc<1, false>(rowPtrs, colIdxs, vals, M, K, nnz, B, N, ldb, C, ldc, pool, st);
}

} // namespace cuspmm_b200
