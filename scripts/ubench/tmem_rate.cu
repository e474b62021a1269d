// tmem_rate.cu -- tcgen05.ld issue rate / bandwidth per shape (x4 .. x32) with a dynamic (shuffled) column address.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X> __device__ __forceinline__ void tld(uint32_t a, uint32_t (&v)[X]);
template <> __device__ __forceinline__ void tld<4>(uint32_t a, uint32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(a));
}
template <> __device__ __forceinline__ void tld<8>(uint32_t a, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(a));
}
template <> __device__ __forceinline__ void tld<16>(uint32_t a, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(a));
}

// J loads of X columns per wait::ld; FM = FMAs per loaded float (0: one add per load, 1: full fma on every element)
template <int X, int J, int FM>
__global__ void __launch_bounds__(1024, 1) rate_kernel(uint32_t iters, float *out, unsigned long long *cycles) {
    __shared__ uint32_t tmem_slot;
    const uint32_t warp = __shfl_sync(0xFFFFFFFFu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_slot + (((warp & 3u) * 32u) << 16);
    uint32_t baddr[J];
    float bval[J];
    float acc[J][X];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        baddr[j] = tbase + (((lane * 7u + warp * 3u + j * 11u) % (512u / X)) * X);
        bval[j] = 1.f + 0.01f * j + 1e-3f * lane;
#pragma unroll
        for (int x = 0; x < X; ++x) acc[j][x] = 0.f;
    }
    const long long t0 = clock64();
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll 2
        for (uint32_t t = 0; t < 32; ++t) {
            uint32_t v[J][X];
            float s[J];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const uint32_t a = __shfl_sync(0xFFFFFFFFu, baddr[j], t);
                s[j] = __shfl_sync(0xFFFFFFFFu, bval[j], t);
                tld<X>(a, v[j]);
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if (FM == 0) acc[j][0] += __uint_as_float(v[j][0]) * s[j];
                else {
#pragma unroll
                    for (int x = 0; x < X; ++x) acc[j][x] = fmaf(__uint_as_float(v[j][x]), s[j], acc[j][x]);
                }
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
        for (int x = 0; x < X; ++x) s += acc[j][x];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    }
}

template <int X, int J, int FM>
static void run(int warps, float *out, unsigned long long *cyc) {
    const uint32_t iters = 300;
    auto k = rate_kernel<X, J, FM>;
    k<<<148, warps * 32>>>(10, out, cyc);
    k<<<148, warps * 32>>>(iters, out, cyc);
    CK(cudaDeviceSynchronize());
    std::vector<unsigned long long> h(148);
    CK(cudaMemcpy(h.data(), cyc, 148 * 8, cudaMemcpyDeviceToHost));
    unsigned long long mx = 0;
    for (auto c : h) mx = c > mx ? c : mx;
    const double loads = (double)warps * iters * 32.0 * J;
    printf("x%-2d J=%d fma=%d warps=%2d : %6.2f clk per LDTM per SM, %6.1f B/clk/SM\n", X, J, FM, warps, (double)mx / loads, loads * X * 128.0 / (double)mx);
}

int main() {
    float *out; unsigned long long *cyc;
    CK(cudaMalloc(&out, 148 * 1024 * 4)); CK(cudaMalloc(&cyc, 148 * 8));
    for (int warps : {8, 16, 28}) {
        run<4, 4, 0>(warps, out, cyc);  run<4, 4, 1>(warps, out, cyc);
        run<8, 4, 0>(warps, out, cyc);  run<8, 4, 1>(warps, out, cyc);
        run<16, 2, 0>(warps, out, cyc); run<16, 2, 1>(warps, out, cyc);
        run<16, 1, 1>(warps, out, cyc);
    }
    return 0;
}
