// spmm_bsr.cu -- fp32 BSR SpMM (any block shape) for sm_100a.
//
// Replaces spmmBSRK1 (src/spmm/bsr/spmm_bsr_k1.cu:9-41: one thread per block element, one
// global atomicAdd per multiply-add).  This is the parity-exact BSR path: fp32 blocks as the
// reference stores them, terms added per C element in exactly spmmBSRCpu's order (blocks of
// the block row in storage order, then ascending column inside the block; zeros stored in a
// block are multiplied too -- src/spmm/bsr/spmm_bsr.cpp:17-39).  One warp per C row; the
// block row is flattened to a (column, value) stream that the warp reads 32 entries at a
// time and broadcasts by shuffle, B rows are read with 128-bit loads.  The tensor-core path
// for 16x16 / 32x32 bf16/fp16 blocks is spmm_bsr_tc.cu.
#include "common.cuh"

namespace cuspmm_b200 {

template <int U, bool VEC>
__global__ void __launch_bounds__(256)
bsr_f32_kernel(const uint32_t *__restrict__ blockRowPtrs, const uint32_t *__restrict__ blockColIdxs,
               const float *__restrict__ blocks, uint32_t M, uint32_t br, uint32_t bc,
               const float *__restrict__ B, uint32_t N, size_t ldb, float *__restrict__ C, size_t ldc) {
    constexpr uint32_t W = VEC ? 4u : 1u;                 // columns per lane per u
    const uint32_t lane = lane_id();
    const uint32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= M) return;
    const uint32_t R = r / br, ri = r % br;
    const uint32_t col0 = blockIdx.y * (32u * W * U) + lane * W;
    const uint32_t bstart = __ldg(blockRowPtrs + R), bend = __ldg(blockRowPtrs + R + 1);
    const uint64_t entries = (uint64_t)(bend - bstart) * bc;   // flattened (block, ac) stream

    float acc[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int x = 0; x < (int)W; ++x) acc[u][x] = 0.f;

    for (uint64_t base = 0; base < entries; base += 32) {
        const uint64_t t = base + lane;
        uint32_t mc = 0;
        float mv = 0.f;
        if (t < entries) {
            const uint32_t b = bstart + (uint32_t)(t / bc), ac = (uint32_t)(t % bc);
            mc = __ldg(blockColIdxs + b) * bc + ac;
            mv = ld_stream(blocks + ((size_t)b * br + ri) * bc + ac);
        }
        const int cnt = (int)min((uint64_t)32, entries - base);
        for (int j = 0; j < cnt; ++j) {
            const uint32_t c = __shfl_sync(0xFFFFFFFFu, mc, j);
            const float v = __shfl_sync(0xFFFFFFFFu, mv, j);
            const float *brow = B + (size_t)c * ldb + col0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (col0 + u * 32u * W < N) {
                    if constexpr (VEC) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4 *>(brow + u * 128));
                        acc[u][0] = fmaf(v, b4.x, acc[u][0]);
                        acc[u][1] = fmaf(v, b4.y, acc[u][1]);
                        acc[u][2] = fmaf(v, b4.z, acc[u][2]);
                        acc[u][3] = fmaf(v, b4.w, acc[u][3]);
                    } else {
                        acc[u][0] = fmaf(v, __ldg(brow + u * 32), acc[u][0]);
                    }
                }
            }
        }
    }
    float *crow = C + (size_t)r * ldc + col0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (col0 + u * 32u * W < N) {
            if constexpr (VEC)
                __stcs(reinterpret_cast<float4 *>(crow + u * 128), make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]));
            else
                crow[u * 32] = acc[u][0];
        }
    }
}

// ------------------------------------------------------------------ block-row kernel (br % 4 == 0, bc <= 32)
// A warp owns BR rows of one block row and 128 columns.  For every stored block it copies its BR x bc
// slice into shared memory (transposed, so the BR values that multiply one B row are contiguous), then for
// each of the bc block columns loads the B row ONCE (LDG.128 per lane) and applies it to all BR rows:
// BR*4 FMAs per 128-bit load instead of 4 -- the register-level reuse of B that unstructured formats
// cannot have.  Per C element the terms are still added in spmmBSRCpu's order (blocks in storage order,
// then ascending column inside the block, zeros included), so the result is bit-identical to bsr_f32_kernel.
template <int BR>
__global__ void __launch_bounds__(256)
bsr_f32_blockrow_kernel(const uint32_t *__restrict__ blockRowPtrs, const uint32_t *__restrict__ blockColIdxs,
                        const float *__restrict__ blocks, uint32_t numBlockRows, uint32_t br, uint32_t bc,
                        const float *__restrict__ B, uint32_t N, size_t ldb, float *__restrict__ C, size_t ldc) {
    constexpr int STRIDE = BR + 4;                       // padded: transposed stores spread over the banks
    __shared__ __align__(16) float sA[8][2][32 * STRIDE];
    const uint32_t lane = lane_id(), wib = threadIdx.x >> 5;
    const uint32_t subs = br / BR;                       // BR-row slices per block row
    const uint32_t gw = blockIdx.x * 8 + wib;
    if (gw >= numBlockRows * subs) return;
    const uint32_t R = gw / subs, rsub = gw % subs;
    const uint32_t col = blockIdx.y * 128u + lane * 4u;
    const bool valid = col < N;
    const uint32_t bstart = __ldg(blockRowPtrs + R), bend = __ldg(blockRowPtrs + R + 1);

    float4 acc[BR];
#pragma unroll
    for (int i = 0; i < BR; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    // (prefetching the next block's slice into registers was tried: 144 registers -> one CTA per SM and
    //  1.7x slower; occupancy hides the block-fetch latency better than the software pipeline did)
    for (uint32_t b = bstart; b < bend; ++b) {
        float *tile = sA[wib][(b - bstart) & 1];
        const float *src = blocks + ((size_t)b * br + (size_t)rsub * BR) * bc;       // BR consecutive block rows
        for (uint32_t e = lane; e < (uint32_t)BR * bc; e += 32) tile[(e % bc) * STRIDE + e / bc] = ld_stream(src + e);
        __syncwarp();
        const float *brow = B + (size_t)__ldg(blockColIdxs + b) * bc * ldb + col;
        for (uint32_t ac0 = 0; ac0 < bc; ac0 += 8) {
            float4 bv[8];
#pragma unroll
            for (int t = 0; t < 8; ++t)
                if (valid && ac0 + t < bc) bv[t] = __ldg(reinterpret_cast<const float4 *>(brow + (size_t)(ac0 + t) * ldb));
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                if (ac0 + t < bc) {
                    const float4 *a4 = reinterpret_cast<const float4 *>(tile + (ac0 + t) * STRIDE);
#pragma unroll
                    for (int q = 0; q < BR / 4; ++q) {
                        const float4 a = a4[q];          // same address in every lane: shared-memory broadcast
                        if (valid) {
                            fma4(acc[q * 4 + 0], a.x, bv[t]);
                            fma4(acc[q * 4 + 1], a.y, bv[t]);
                            fma4(acc[q * 4 + 2], a.z, bv[t]);
                            fma4(acc[q * 4 + 3], a.w, bv[t]);
                        }
                    }
                }
            }
        }
        // double buffered tile + this barrier: a buffer is rewritten only after every lane finished reading it
        __syncwarp();
    }
    if (valid) {
        float *crow = C + ((size_t)R * br + (size_t)rsub * BR) * ldc + col;
#pragma unroll
        for (int i = 0; i < BR; ++i) __stcs(reinterpret_cast<float4 *>(crow + (size_t)i * ldc), acc[i]);
    }
}

} // namespace cuspmm_b200

extern "C" int cuspmm_spmm_bsr_f32(const uint32_t *blockRowPtrs, const uint32_t *blockColIdxs,
                                   const float *blocks, uint32_t numBlockRows, uint32_t br, uint32_t bc,
                                   uint32_t K, const float *B, uint32_t N, size_t ldb,
                                   float *C, size_t ldc, void *stream) {
    using namespace cuspmm_b200;
    (void)K;
    CUSPMM_REQUIRE(br > 0 && bc > 0, "block shape must be positive (got %ux%u)", br, bc);
    CUSPMM_REQUIRE(ldb >= N && ldc >= N, "ldb/ldc must be >= N");
    const uint64_t M64 = (uint64_t)numBlockRows * br;
    CUSPMM_REQUIRE(M64 <= 0xFFFFFFFFull, "numBlockRows * br overflows uint32");
    const uint32_t M = (uint32_t)M64;
    if (M == 0 || N == 0) return CUSPMM_OK;
    CUSPMM_REQUIRE(blockRowPtrs && B && C, "null operand pointer");
    cudaStream_t st = as_stream(stream);
    const bool vok = (N % 4 == 0) && (ldb % 4 == 0) && (ldc % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    // blocks whose height is a multiple of 4 (and at most 32 wide): block-row kernel with register reuse of B
    if (vok && bc <= 32 && br % 4 == 0) {
        const uint32_t BR = (br % 16 == 0) ? 16 : (br % 8 == 0 ? 8 : 4);
        const uint32_t warps = numBlockRows * (br / BR);
        dim3 grid((warps + 7) / 8, (N + 127) / 128);
        if (BR == 16) bsr_f32_blockrow_kernel<16><<<grid, 256, 0, st>>>(blockRowPtrs, blockColIdxs, blocks, numBlockRows, br, bc, B, N, ldb, C, ldc);
        else if (BR == 8) bsr_f32_blockrow_kernel<8><<<grid, 256, 0, st>>>(blockRowPtrs, blockColIdxs, blocks, numBlockRows, br, bc, B, N, ldb, C, ldc);
        else bsr_f32_blockrow_kernel<4><<<grid, 256, 0, st>>>(blockRowPtrs, blockColIdxs, blocks, numBlockRows, br, bc, B, N, ldb, C, ldc);
        CUSPMM_LAUNCH_CHECK("bsr_f32_blockrow_kernel");
        return CUSPMM_OK;
    }
    const uint32_t blocksX = (M + 7) / 8;
    if (vok) {
        if (N > 256) bsr_f32_kernel<4, true><<<dim3(blocksX, (N + 511) / 512), 256, 0, st>>>(blockRowPtrs, blockColIdxs, blocks, M, br, bc, B, N, ldb, C, ldc);
        else if (N > 128) bsr_f32_kernel<2, true><<<dim3(blocksX, 1), 256, 0, st>>>(blockRowPtrs, blockColIdxs, blocks, M, br, bc, B, N, ldb, C, ldc);
        else bsr_f32_kernel<1, true><<<dim3(blocksX, 1), 256, 0, st>>>(blockRowPtrs, blockColIdxs, blocks, M, br, bc, B, N, ldb, C, ldc);
    } else {
        bsr_f32_kernel<4, false><<<dim3(blocksX, (N + 127) / 128), 256, 0, st>>>(blockRowPtrs, blockColIdxs, blocks, M, br, bc, B, N, ldb, C, ldc);
    }
    CUSPMM_LAUNCH_CHECK("bsr_f32_kernel");
    return CUSPMM_OK;
}
