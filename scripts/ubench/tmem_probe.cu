// tmem_probe.cu -- micro-benchmark behind the TMEM-staged CSR kernel:
//   1. bandwidth of tcgen05.ld (TMEM -> registers) with a dynamic column address, against LDS.128 with a
//      dynamic row address, and the two mixed, at 4..32 warps per SM;
//   2. layout produced by tcgen05.cp.128x256b from a row-major fp32 tile of B in shared memory
//      (no-swizzle descriptor, SBO = 128 B, LBO = one row of the tile).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float &a, float &b, float &c, float &d) {
    uint32_t x, y, z, w;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(taddr));
    a = __uint_as_float(x); b = __uint_as_float(y); c = __uint_as_float(z); d = __uint_as_float(w);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// mode 0: TMEM only, 1: LDS only, 2: alternate TMEM / LDS.  J loads in flight per wait.
template <int MODE, int J>
__global__ void __launch_bounds__(1024, 1) bw_kernel(uint32_t iters, float *out, unsigned long long *cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float4 *tile = reinterpret_cast<float4 *>(smem);          // 128 rows x 128 float4 (2 KB rows) = 256 KB?  use 64 rows
    for (uint32_t i = threadIdx.x; i < 64 * 128; i += blockDim.x) tile[i] = make_float4(1e-3f * i, 1.f, 2.f, 3.f);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_slot + (((warp & 3u) * 32u) << 16);
    float4 acc[4];
    for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t k = warp * 7u + 3u;
    const long long t0 = clock64();
    for (uint32_t it = 0; it < iters; ++it) {
        float4 b[J];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            k = (k * 13u + 5u) & 63u;                        // warp-uniform pseudo-random row of the chunk
            const bool useT = MODE == 0 || (MODE == 2 && (j & 1) == 0);
            if (useT) tmem_ld4(tbase + ((k & 31u) * 4u + (j & 1) * 128u), b[j].x, b[j].y, b[j].z, b[j].w);
            else b[j] = tile[k * 128u + (j & 3) * 32u + lane];
        }
        if (MODE != 1) tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < J; ++j) {
            acc[j & 3].x = fmaf(b[j].x, 1.0001f, acc[j & 3].x);
            acc[j & 3].y = fmaf(b[j].y, 1.0001f, acc[j & 3].y);
            acc[j & 3].z = fmaf(b[j].z, 1.0001f, acc[j & 3].z);
            acc[j & 3].w = fmaf(b[j].w, 1.0001f, acc[j & 3].w);
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
    for (int u = 0; u < 4; ++u) s += acc[u].x + acc[u].y + acc[u].z + acc[u].w;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    }
}

template <int MODE, int J>
static void run_bw(const char *name, int warps, float *out, unsigned long long *cyc) {
    const uint32_t iters = 20000;
    auto k = bw_kernel<MODE, J>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 2048));
    k<<<148, warps * 32, 64 * 2048>>>(100, out, cyc);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    k<<<148, warps * 32, 64 * 2048>>>(iters, out, cyc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<unsigned long long> h(148);
    CK(cudaMemcpy(h.data(), cyc, 148 * 8, cudaMemcpyDeviceToHost));
    unsigned long long mx = 0;
    for (auto c : h) mx = c > mx ? c : mx;
    const double bytes = (double)warps * iters * J * 512.0;
    printf("%-12s J=%d warps=%2d  %8.1f B/clk/SM  (%.3f ms, %llu cyc, %.0f MHz)\n", name, J, warps, bytes / (double)mx, ms, mx,
           (double)mx / ms / 1e3);
}

// ------------------------------------------------------------------ tcgen05.cp layout check
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// tile: KC rows x 512 fp32 (2 KB per row).  One thread copies it into TMEM with KC/2 128x256b copies:
// expected TMEM[lane L][col 4*kk + j] = tile[kk][4*L + j].
template <int KC>
__global__ void __launch_bounds__(128) cp_kernel(const float *__restrict__ src, float *__restrict__ dst) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    float *tile = reinterpret_cast<float *>(smem);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x; i < KC * 512; i += blockDim.x) tile[i] = src[i];
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_slot;
    if (threadIdx.x == 0) {
        for (int kk = 0; kk < KC; kk += 2) {
            const uint64_t desc = make_desc(smem_u32(tile + kk * 512), 2048, 128);
            asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tbase + kk * 4), "l"(desc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int kk = 0; kk < KC; ++kk) {
        float a, b, c, d;
        tmem_ld4(tbase + ((warp * 32u) << 16) + kk * 4, a, b, c, d);
        tmem_wait_ld();
        float *o = dst + kk * 512 + (warp * 32 + lane) * 4;
        o[0] = a; o[1] = b; o[2] = c; o[3] = d;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
    }
}

int main() {
    float *out;
    unsigned long long *cyc;
    CK(cudaMalloc(&out, 148 * 1024 * 4));
    CK(cudaMalloc(&cyc, 148 * 8));
    {   // layout check first
        constexpr int KC = 32;
        std::vector<float> h(KC * 512), g(KC * 512, -1.f);
        for (int i = 0; i < KC * 512; ++i) h[i] = (float)i;
        float *s, *d;
        CK(cudaMalloc(&s, KC * 512 * 4)); CK(cudaMalloc(&d, KC * 512 * 4));
        CK(cudaMemcpy(s, h.data(), KC * 512 * 4, cudaMemcpyHostToDevice));
        CK(cudaFuncSetAttribute(cp_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, KC * 2048));
        cp_kernel<KC><<<1, 128, KC * 2048>>>(s, d);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(g.data(), d, KC * 512 * 4, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int i = 0; i < KC * 512; ++i) if (g[i] != h[i]) { if (bad < 8) printf("  cp mismatch at k=%d n=%d: got %.0f want %.0f\n", i / 512, i % 512, g[i], h[i]); ++bad; }
        printf("tcgen05.cp.128x256b layout check: %d mismatches of %d\n", bad, KC * 512);
    }
    for (int warps : {4, 8, 16, 32}) {
        run_bw<1, 4>("LDS.128", warps, out, cyc);
        run_bw<0, 4>("TMEM.x4", warps, out, cyc);
        run_bw<0, 8>("TMEM.x4", warps, out, cyc);
        run_bw<2, 8>("mixed 1:1", warps, out, cyc);
    }
    return 0;
}
