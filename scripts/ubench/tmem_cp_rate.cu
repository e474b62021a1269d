// tmem_cp_rate.cu -- cost of tcgen05.cp (shared memory -> TMEM) per shape, issued by one thread.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
template <int SHAPE> __device__ __forceinline__ void cp(uint32_t taddr, uint64_t desc) {
    if (SHAPE == 0) asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
    if (SHAPE == 1) asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
    if (SHAPE == 2) asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
    if (SHAPE == 3) asm volatile("tcgen05.cp.cta_group::1.64x128b.warpx2::02_13 [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
    if (SHAPE == 4) asm volatile("tcgen05.cp.cta_group::1.4x256b [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}
// NT issuing threads (lane 0 of NT warps), each n copies, then a commit each; the CTA waits for all
template <int SHAPE, int NT>
__global__ void __launch_bounds__(1024) cp_rate(uint32_t n, unsigned long long *cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"((uint32_t)NT));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_slot;
    const long long t0 = clock64();
    if (lane == 0 && warp < NT) {
        const uint64_t desc = make_desc(smem_u32(smem), 2048, 128);
        for (uint32_t i = 0; i < n; ++i) cp<SHAPE>(tbase + ((i * 8u + warp * 128u) & 504u), desc + (uint64_t)((i & 15u) * 256u));
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
    }
}
template <int SHAPE, int NT>
static void run(const char *name, int srcBytes, int dstBytes, unsigned long long *cyc) {
    const uint32_t n = 2048;
    CK(cudaFuncSetAttribute(cp_rate<SHAPE, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    cp_rate<SHAPE, NT><<<148, (NT < 4 ? 4 : NT) * 32, 128 * 1024>>>(64, cyc);
    cp_rate<SHAPE, NT><<<148, (NT < 4 ? 4 : NT) * 32, 128 * 1024>>>(n, cyc);
    CK(cudaDeviceSynchronize());
    unsigned long long h[148], mx = 0;
    CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    for (auto c : h) mx = c > mx ? c : mx;
    const double per = (double)mx / (n * NT);
    printf("%-26s issuers=%d : %7.1f clk per copy, smem read %6.1f B/clk, TMEM write %6.1f B/clk\n", name, NT, per, srcBytes / per, dstBytes / per);
}
int main() {
    unsigned long long *cyc;
    CK(cudaMalloc(&cyc, 148 * 8));
    run<0, 1>("128x256b", 4096, 4096, cyc);
    run<0, 4>("128x256b", 4096, 4096, cyc);
    run<1, 1>("128x128b", 2048, 2048, cyc);
    run<1, 4>("128x128b", 2048, 2048, cyc);
    run<2, 1>("32x128b.warpx4", 512, 2048, cyc);
    run<2, 4>("32x128b.warpx4", 512, 2048, cyc);
    run<3, 1>("64x128b.warpx2::02_13", 1024, 2048, cyc);
    run<3, 4>("64x128b.warpx2::02_13", 1024, 2048, cyc);
    run<4, 1>("4x256b", 128, 128, cyc);
    run<2, 8>("32x128b.warpx4", 512, 2048, cyc);
    run<2, 16>("32x128b.warpx4", 512, 2048, cyc);
    run<2, 32>("32x128b.warpx4", 512, 2048, cyc);
    run<3, 8>("64x128b.warpx2::02_13", 1024, 2048, cyc);
    run<3, 16>("64x128b.warpx2::02_13", 1024, 2048, cyc);
    run<0, 8>("128x256b", 4096, 4096, cyc);
    run<1, 8>("128x128b", 2048, 2048, cyc);
    return 0;
}
