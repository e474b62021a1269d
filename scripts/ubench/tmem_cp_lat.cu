// tmem_cp_lat.cu -- latency of a batch of tcgen05.cp + commit as seen by a polling thread of another warp:
// NI issuer warps each issue `per` copies (32x128b.warpx4) and commit on one mbarrier (count NI).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t a) {
    return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ bool try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
template <int NI>
__global__ void __launch_bounds__(1024) lat_kernel(uint32_t per, uint32_t rounds, unsigned long long *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t go, done;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&go)), "r"(1u));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&done)), "r"((uint32_t)NI));
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
    unsigned long long total = 0;
    if (warp < NI) {
        if (lane == 0) {
            const uint64_t desc = make_desc(smem_u32(smem));
            for (uint32_t r = 0; r < rounds; ++r) {
                while (!try_wait(&go, r & 1)) {}
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (uint32_t i = 0; i < per; ++i)
                    asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(tbase + ((i * NI + warp) & 127u) * 4u), "l"(desc + (uint64_t)(((i * NI + warp) & 63u) * 32u)) : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
            }
        }
    } else if (warp == NI) {
        if (lane == 0) {
            for (uint32_t r = 0; r < rounds; ++r) {
                const long long t0 = clock64();
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&go)) : "memory");
                while (!try_wait(&done, r & 1)) {}
                total += (unsigned long long)(clock64() - t0);
            }
            out[blockIdx.x] = total / rounds;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
    }
}
template <int NI> static void run(uint32_t per, unsigned long long *d) {
    CK(cudaFuncSetAttribute(lat_kernel<NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    lat_kernel<NI><<<148, (NI + 1) * 32, 64 * 1024>>>(per, 200, d);
    CK(cudaDeviceSynchronize());
    unsigned long long h[148], s = 0;
    CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    for (auto c : h) s += c;
    printf("issuers=%2d copies each=%2u (total %3u x 512 B): go -> all commits seen = %5llu clk\n", NI, per, per * NI, s / 148);
}
int main() {
    unsigned long long *d;
    CK(cudaMalloc(&d, 148 * 8));
    run<1>(0, d); run<1>(1, d); run<1>(8, d); run<8>(1, d); run<8>(8, d); run<16>(4, d); run<8>(4, d); run<4>(16, d); run<16>(1, d);
    return 0;
}
