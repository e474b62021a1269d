// tmem_common.cuh -- pieces shared by the kernels that stage B in tensor memory (spmm_csr_tmem.cu: dual operand path,
// spmm_csr_quad.cu: all reads from TMEM): tcgen05 fences / commit, non-blocking barrier test, the 2-D tensor-map TMA
// copy of a B chunk, the tensor-map encoder and the whole-waves grid plan.
#pragma once

#include "common.cuh"

#include <cuda.h>      // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no -lcuda)
#include <mutex>
#include <stdlib.h>
#include <string.h>

namespace cuspmm_b200 {
namespace tmemk {

using pipe::smem_u32;
using pipe::mbar_init;
using pipe::mbar_expect_tx;
using pipe::mbar_arrive;
using pipe::bulk_g2s;

#ifdef __CUDACC__
// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (woken by the arrive) or the
// hint elapses, instead of spinning
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) { pipe::mbar_wait<2000>(bar, parity); }
// one non-blocking test of a phase (true: the phase with this parity has completed)
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
// 2-D tiled TMA copy: box (c0 .. , c1 ..) of the tensor map -> shared memory, completion (full box bytes, rows past the end of
// the tensor arrive as zeros) on the mbarrier
__device__ __forceinline__ void tma_box_2d(void *dst, const CUtensorMap *map, uint32_t c0, uint32_t c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// arrives on bar once every tcgen05 operation issued so far by this thread (here: the copies) is complete
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// no-swizzle shared-memory matrix descriptor: core matrices of 8 rows x 16 B (128 B contiguous); the next 8 rows are
// SBO bytes further, the next 16-byte column LBO bytes further
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
#endif

constexpr int kNT = 512;                    // columns of B/C per CTA

// whole waves of the SM count, rows per CTA = ceil(M / panels) <= rowSlots (as the staged kernel)
struct GridPlan { uint32_t panels, rpc; };
static inline GridPlan plan_grid(uint32_t M, uint32_t ytiles, uint32_t rowSlots) {
    const uint32_t sms = (uint32_t)sm_count();
    const uint32_t minPanels = (M + rowSlots - 1) / rowSlots;
    GridPlan g;
    const uint32_t waves = (minPanels * ytiles + sms - 1) / sms;
    g.panels = (waves * sms) / ytiles;
    if (g.panels < minPanels) g.panels = minPanels;
    g.rpc = (M + g.panels - 1) / g.panels;
    if (g.rpc > rowSlots) g.rpc = rowSlots;
    if (g.rpc == 0) g.rpc = 1;
    g.panels = (M + g.rpc - 1) / g.rpc;
    return g;
}

// B as a 2-D tensor of 8-byte elements (so that a 512-column box is 256 elements, the TMA box limit): dims (N/2, K), row
// pitch ldb * 4 bytes, box (256, boxRows), no swizzle, rows past K read as zeros.  The driver's encoder is fetched once per process.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn encode = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            encode = reinterpret_cast<EncodeTiledFn>(fn);
    });
    return encode;
}
static inline bool make_tmap_B(CUtensorMap *map, const float *B, uint32_t K, uint32_t N, size_t ldb, uint32_t boxRows) {
    EncodeTiledFn encode = tensor_map_encoder();
    if (!encode || (N & 1) || ((ldb * sizeof(float)) & 15) || (reinterpret_cast<uintptr_t>(B) & 15)) return false;
    const cuuint64_t dims[2] = {N / 2, K};
    const cuuint64_t strides[1] = {(cuuint64_t)ldb * sizeof(float)};
    const cuuint32_t box[2] = {256, boxRows};
    const cuuint32_t estr[2] = {1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<float *>(B), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

} // namespace tmemk
} // namespace cuspmm_b200
