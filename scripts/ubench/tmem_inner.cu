// tmem_inner.cu -- inner-loop study for the TMEM-assisted staged CSR kernel: the consumer loop of
// csr_staged (shuffle-broadcast of (col, val) from a register window, B row fetched per non-zero, FMAs)
// with three operand paths:
//   A  today's loop: warp = RW rows x 512 columns, 4 x LDS.128 per non-zero
//   B  warp = RW rows x 128 columns (its TMEM lane quarter); non-zeros alternate between tcgen05.ld.x4
//      (TMEM) and LDS.128 (shared memory)
//   C  warp = RW rows x 512 columns; B rows with k % 4 == quarter live in TMEM (one tcgen05.ld.x16),
//      the others come from shared memory (4 x LDS.128)
// Reports SM cycles per non-zero of a 512-column row (the staged kernel measures 18.6 on large_25605).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fma4(float4 &acc, float a, const float4 &b) {
    acc.x = fmaf(a, b.x, acc.x); acc.y = fmaf(a, b.y, acc.y); acc.z = fmaf(a, b.z, acc.z); acc.w = fmaf(a, b.w, acc.w);
}
__device__ __forceinline__ float4 tmem_ld4(uint32_t taddr) {
    uint32_t x, y, z, w;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(taddr));
    return make_float4(__uint_as_float(x), __uint_as_float(y), __uint_as_float(z), __uint_as_float(w));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float4 (&b)[4]) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
#pragma unroll
    for (int u = 0; u < 4; ++u)
        b[u] = make_float4(__uint_as_float(v[4 * u]), __uint_as_float(v[4 * u + 1]), __uint_as_float(v[4 * u + 2]), __uint_as_float(v[4 * u + 3]));
}
__device__ __forceinline__ void fma4x2(float4 &acc, float a, const float4 &b) {
    // two packed fp32x2 FMAs (FFMA2): same IEEE result per component as four FFMAs, half the issue slots
    unsigned long long a2, b01, b23, c01, c23;
    asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(a));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b01) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b23) : "f"(b.z), "f"(b.w));
    asm("mov.b64 %0, {%1, %2};" : "=l"(c01) : "f"(acc.x), "f"(acc.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(c23) : "f"(acc.z), "f"(acc.w));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c01) : "l"(a2), "l"(b01));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(c23) : "l"(a2), "l"(b23));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.x), "=f"(acc.y) : "l"(c01));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc.z), "=f"(acc.w) : "l"(c23));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// MODE 0 = A, 1 = B, 2 = C, 3 = B with TMEM only (no LDS)
template <int MODE, int RW>
__global__ void __launch_bounds__(1024, 1) inner_kernel(uint32_t iters, float *out, unsigned long long *cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    const uint32_t warp = __shfl_sync(0xFFFFFFFFu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    float4 *tile = reinterpret_cast<float4 *>(smem);          // 32 rows x 128 float4 = 64 KB
    for (uint32_t i = threadIdx.x; i < 32 * 128; i += blockDim.x) tile[i] = make_float4(1e-3f * i, 1.f, 2.f, 3.f);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t q = warp & 3u;
    const uint32_t tbase = tmem_slot + ((q * 32u) << 16);
    constexpr int U = (MODE == 1 || MODE == 3) ? 1 : 4;
    float4 acc[RW][U];
    uint32_t bcol[RW];
    float bval[RW];
#pragma unroll
    for (int i = 0; i < RW; ++i) {
        bcol[i] = (lane * 7u + warp * 3u + i * 11u) & 31u;     // chunk-local B row of window entry `lane`
        bval[i] = 1.0f + 1e-3f * lane;
#pragma unroll
        for (int u = 0; u < U; ++u) acc[i][u] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float4 *tl = tile + lane;
    const long long t0 = clock64();
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll (U == 4 ? 1 : 4)
        for (uint32_t t = 0; t < 32; ++t) {
            float4 b[RW][U];
            float v[RW];
            bool anyT = false;
#pragma unroll
            for (int i = 0; i < RW; ++i) {
                const uint32_t c = __shfl_sync(0xFFFFFFFFu, bcol[i], t);
                v[i] = __shfl_sync(0xFFFFFFFFu, bval[i], t);
                if constexpr (MODE == 0) {
#pragma unroll
                    for (int u = 0; u < U; ++u) b[i][u] = tl[c * 128u + u * 32u];
                } else if constexpr (MODE == 1) {
                    if ((i & 1) == 0) b[i][0] = tmem_ld4(tbase + c * 4u);
                    else b[i][0] = tl[c * 128u + q * 32u];
                } else if constexpr (MODE == 3) {
                    b[i][0] = tmem_ld4(tbase + c * 4u);
                } else {
                    if ((c & 3u) == q) { tmem_ld16(tbase + (c >> 2) * 16u, b[i]); anyT = true; }
                    else {
#pragma unroll
                        for (int u = 0; u < U; ++u) b[i][u] = tl[c * 128u + u * 32u];
                    }
                }
            }
            if (MODE == 1 || MODE == 3) tmem_wait_ld();
            if (MODE == 2 && anyT) tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < RW; ++i)
#pragma unroll
                for (int u = 0; u < U; ++u) fma4(acc[i][u], v[i], b[i][u]);
        }
#pragma unroll
        for (int i = 0; i < RW; ++i) bcol[i] = (bcol[i] + 5u) & 31u;
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < RW; ++i)
#pragma unroll
        for (int u = 0; u < U; ++u) s += acc[i][u].x + acc[i][u].y + acc[i][u].z + acc[i][u].w;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    }
}

// MODE 4: TMEM only, window holds the ready-made TMEM address; HB rows per wait::ld; FFMA2 or FFMA.  MODE 5: loads only.
template <int RW, int HB, bool F2, bool NOFMA>
__global__ void __launch_bounds__(1024, 1) inner2_kernel(uint32_t iters, float *out, unsigned long long *cycles) {
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
    float4 acc[RW];
    uint32_t baddr[RW];
    float bval[RW];
#pragma unroll
    for (int i = 0; i < RW; ++i) {
        baddr[i] = tbase + (((lane * 7u + warp * 3u + i * 11u) & 127u) << 2);
        bval[i] = 1.0f + 1e-3f * lane + 0.01f * i;
        acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const long long t0 = clock64();
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll 2
        for (uint32_t t = 0; t < 32; ++t) {
#pragma unroll
            for (int h = 0; h < RW; h += HB) {
                float4 b[HB];
                float v[HB];
#pragma unroll
                for (int i = 0; i < HB; ++i) {
                    const uint32_t a = __shfl_sync(0xFFFFFFFFu, baddr[h + i], t);
                    v[i] = __shfl_sync(0xFFFFFFFFu, bval[h + i], t);
                    b[i] = tmem_ld4(a);
                }
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < HB; ++i) {
                    if (NOFMA) { acc[h + i].x += b[i].x * v[i]; }
                    else if (F2) fma4x2(acc[h + i], v[i], b[i]);
                    else fma4(acc[h + i], v[i], b[i]);
                }
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < RW; ++i) s += acc[i].x + acc[i].y + acc[i].z + acc[i].w;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    }
}

template <int RW, int HB, bool F2, bool NOFMA>
static void run2(const char *name, int warps, float *out, unsigned long long *cyc) {
    const uint32_t iters = 400;
    auto k = inner2_kernel<RW, HB, F2, NOFMA>;
    k<<<148, warps * 32>>>(10, out, cyc);
    k<<<148, warps * 32>>>(iters, out, cyc);
    CK(cudaDeviceSynchronize());
    std::vector<unsigned long long> h(148);
    CK(cudaMemcpy(h.data(), cyc, 148 * 8, cudaMemcpyDeviceToHost));
    unsigned long long mx = 0;
    for (auto c : h) mx = c > mx ? c : mx;
    const double nnz = (double)warps * iters * 32.0 * RW * 0.25;
    printf("%-34s RW=%d HB=%d warps=%2d  %6.2f clk per non-zero (512 cols)  = %5.1f B/clk TMEM\n", name, RW, HB, warps, (double)mx / nnz, 2048.0 * nnz / (double)mx);
}

template <int MODE, int RW>
static void run(const char *name, int warps, float *out, unsigned long long *cyc) {
    const uint32_t iters = 400;
    auto k = inner_kernel<MODE, RW>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    k<<<148, warps * 32, 64 * 1024>>>(10, out, cyc);
    k<<<148, warps * 32, 64 * 1024>>>(iters, out, cyc);
    CK(cudaDeviceSynchronize());
    std::vector<unsigned long long> h(148);
    CK(cudaMemcpy(h.data(), cyc, 148 * 8, cudaMemcpyDeviceToHost));
    unsigned long long mx = 0;
    for (auto c : h) mx = c > mx ? c : mx;
    const double slices = (MODE == 1 || MODE == 3) ? 0.25 : 1.0;      // a mode-B warp handles a quarter of a 512-column row
    const double nnz = (double)warps * iters * 32.0 * RW * slices;
    printf("%-34s RW=%d warps=%2d  %6.2f clk per non-zero (512 cols)\n", name, RW, warps, (double)mx / nnz);
}

int main() {
    float *out;
    unsigned long long *cyc;
    CK(cudaMalloc(&out, 148 * 1024 * 4));
    CK(cudaMalloc(&cyc, 148 * 8));
    for (int warps : {16, 32}) {
        run<0, 2>("A  LDS only, 512 cols/warp", warps, out, cyc);
        run<1, 2>("B  TMEM.x4 / LDS 1:1, 128 cols/warp", warps, out, cyc);
        run<1, 4>("B  TMEM.x4 / LDS 1:1, 128 cols/warp", warps, out, cyc);
        run<3, 2>("B' TMEM.x4 only, 128 cols/warp", warps, out, cyc);
        run<3, 4>("B' TMEM.x4 only, 128 cols/warp", warps, out, cyc);
        run<2, 2>("C  TMEM.x16 (k%4==q) / LDS, 512 cols", warps, out, cyc);
    }
    for (int warps : {16, 24, 28, 32}) {
        run2<4, 4, false, false>("D  TMEM only, addr window, FFMA", warps, out, cyc);
        run2<4, 4, true, false>("D  TMEM only, addr window, FFMA2", warps, out, cyc);
        run2<8, 4, true, false>("D  TMEM only, addr window, FFMA2", warps, out, cyc);
        run2<8, 8, true, false>("D  TMEM only, addr window, FFMA2", warps, out, cyc);
        run2<8, 8, true, true>("E  TMEM loads, 1 FMA each", warps, out, cyc);
    }
    return 0;
}
