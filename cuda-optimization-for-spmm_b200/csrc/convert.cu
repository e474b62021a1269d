// convert.cu -- format conversion, padding and row-panel partitioning ON THE DEVICE
// (north_star (b)).  The reference has no conversion code in C++ (SparseMatrixBSR::fromDense
// throws, src/formats/sparse_bsr.cu:259); offline it is utils/python_utils/convert_mtx.py
// (scipy tocsr/tocoo/tocsc/tobsr).  Every converter here is checked bit for bit against the
// numpy restatements in oracle/oracle.py.  CSR -> BSR is sort-free (a shared-memory bitmap per block row);
// column-ELL -> CSR, and CSR -> BSR for matrices with more than 393 216 block columns, sort with cub::DeviceRadixSort
// (part of the CUDA toolkit); everything else is hand-written.
#include "common.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <stdlib.h>

namespace cuspmm_b200 {

// ------------------------------------------------------------------ small utilities
// Exclusive scan of n uint32 by ONE block (n is a slice / block-row count: <= a few 100k).
__global__ void __launch_bounds__(1024) scan_exclusive_1block(const uint32_t *in, uint32_t *out, uint32_t n,
                                                             uint32_t scale) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = (i < n) ? in[i] * scale : 0u;
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
            if (lane >= (uint32_t)d) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t s = warp_sums[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t y = __shfl_up_sync(0xFFFFFFFFu, s, d);
                if (lane >= (uint32_t)d) s += y;
            }
            warp_sums[lane] = s;     // inclusive over warps
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        const uint32_t incl = x + (warp ? warp_sums[warp - 1] : 0u) + carry;
        if (i < n) out[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_s = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry_s;
}

// ------------------------------------------------------------------ COO -> CSR row pointers
// rowPtrs[r] = first index whose row is >= r.  Thread i looks at the boundary between
// entries i-1 and i and writes every row pointer that falls on it.
__global__ void coo_rowptr_kernel(const uint32_t *__restrict__ rowIdxs, uint32_t M, uint32_t nnz,
                                  uint32_t *__restrict__ rowPtrs) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nnz) return;
    const int64_t prev = (i == 0) ? -1 : (int64_t)rowIdxs[i - 1];
    const int64_t cur = (i == nnz) ? (int64_t)M : (int64_t)rowIdxs[i];
    for (int64_t r = prev + 1; r <= cur; ++r) rowPtrs[r] = (uint32_t)i;
}

// The same pointers by M + 1 independent binary searches over the sorted row indices: ~log2(nnz) dependent loads per row
// instead of streaming all of rowIdxs (25605^2 at 10 %: 0.7 M probes against 262 MB; 0.17 ms -> ~0.02 ms).
__global__ void coo_rowptr_search_kernel(const uint32_t *__restrict__ rowIdxs, uint32_t M, uint32_t nnz,
                                         uint32_t *__restrict__ rowPtrs) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > M) return;
    uint32_t lo = 0, hi = nnz;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(rowIdxs + mid) < r) lo = mid + 1;
        else hi = mid;
    }
    rowPtrs[r] = lo;
}

int coo_to_csr_rowptrs(const uint32_t *rowIdxs, uint32_t M, uint32_t nnz, uint32_t *rowPtrs, cudaStream_t st) {
    if ((uint64_t)nnz >= 64ull * ((uint64_t)M + 1)) {     // long rows: searching is cheaper than streaming
        coo_rowptr_search_kernel<<<(M + 1 + 127) / 128, 128, 0, st>>>(rowIdxs, M, nnz, rowPtrs);
        CUSPMM_LAUNCH_CHECK("coo_rowptr_search_kernel");
        return CUSPMM_OK;
    }
    const uint64_t n = (uint64_t)nnz + 1;
    coo_rowptr_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rowIdxs, M, nnz, rowPtrs);
    CUSPMM_LAUNCH_CHECK("coo_rowptr_kernel");
    return CUSPMM_OK;
}

// Row pointers of a PANEL of a row-sorted COO matrix: entries [idxBase, idxBase + cnt) hold the rows [rowBase, rowBase + rows);
// rowPtrs[j] = idxBase + (first entry of the panel whose row is >= rowBase + j), j = 0 .. rows, i.e. ABSOLUTE offsets into
// the full colIdxs / vals arrays (the host pipeline runs the CSR kernels on such views).  rowIdxs points at the panel's
// first entry.  One binary search per row.
__global__ void coo_panel_rowptr_kernel(const uint32_t *__restrict__ rowIdxs, uint32_t rowBase, uint32_t rows, uint32_t cnt,
                                        uint32_t idxBase, uint32_t *__restrict__ rowPtrs) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j > rows) return;
    const uint32_t r = rowBase + j;
    uint32_t lo = 0, hi = cnt;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(rowIdxs + mid) < r) lo = mid + 1;
        else hi = mid;
    }
    rowPtrs[j] = idxBase + lo;
}

int coo_panel_rowptrs(const uint32_t *rowIdxs, uint32_t rowBase, uint32_t rows, uint32_t cnt, uint32_t idxBase,
                      uint32_t *rowPtrs, cudaStream_t st) {
    coo_panel_rowptr_kernel<<<(rows + 1 + 127) / 128, 128, 0, st>>>(rowIdxs, rowBase, rows, cnt, idxBase, rowPtrs);
    CUSPMM_LAUNCH_CHECK("coo_panel_rowptr_kernel");
    return CUSPMM_OK;
}

// ------------------------------------------------------------------ sortedness check (precondition of the staged kernels)
__global__ void csr_check_sorted_kernel(const uint32_t *__restrict__ rowPtrs, const uint32_t *__restrict__ colIdxs, uint32_t M,
                                        uint32_t K, unsigned int *__restrict__ bad) {
    // one warp per row: entry i must be > entry i - 1 (strictly ascending) and < K.  The predecessor comes from the
    // neighbouring lane; only lane 0 re-reads the last entry of the previous 32.  Iterations are independent (no carried
    // value: a version that carried the last lane's entry serialised every load behind a shuffle and ran 20x slower).
    const uint32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= M) return;
    const uint32_t lane = lane_id();
    const uint32_t p0 = __ldg(rowPtrs + r), p1 = __ldg(rowPtrs + r + 1);
    bool wrong = p1 < p0;
#pragma unroll 4
    for (uint32_t base = p0; base < p1; base += 32) {
        const uint32_t i = base + lane;
        const bool live = i < p1;
        const uint32_t c = live ? __ldg(colIdxs + i) : 0xFFFFFFFFu;
        uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, c, 1);
        if (lane == 0 && base > p0) prev = __ldg(colIdxs + i - 1);
        if (live && (c >= K || (i > p0 && prev >= c))) wrong = true;
    }
    if (__any_sync(0xFFFFFFFFu, wrong) && lane == 0) atomicAdd(bad, 1u);
}

// ------------------------------------------------------------------ nnz-balanced row panels
__global__ void partition_kernel(const uint32_t *__restrict__ rowPtrs, uint32_t M, uint32_t nnz, uint32_t parts,
                                 uint32_t *__restrict__ splits) {
    const uint32_t g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g > parts) return;
    uint32_t s;
    if (g == 0) s = 0;
    else if (g == parts) s = M;
    else {
        const uint64_t target = ((uint64_t)g * nnz) / parts;
        auto key = [&](uint32_t p) -> uint64_t { return (uint64_t)__ldg(rowPtrs + p); };
        s = warp_lower_bound(M, target, key);
    }
    if (lane_id() == 0) splits[g] = s;
}

// ------------------------------------------------------------------ CSR -> sliced ELL
__global__ void sell_width_kernel(const uint32_t *__restrict__ rowPtrs, uint32_t M, uint32_t *__restrict__ widths) {
    // one warp per slice of 32 rows
    const uint32_t s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t slices = (M + 31) / 32;
    if (s >= slices) return;
    const uint32_t r = s * 32 + lane_id();
    uint32_t len = 0;
    if (r < M) len = rowPtrs[r + 1] - rowPtrs[r];
    len = __reduce_max_sync(0xFFFFFFFFu, len);
    if (lane_id() == 0) widths[s] = len;
}

__global__ void __launch_bounds__(256)
sell_fill_kernel(const uint32_t *__restrict__ rowPtrs, const uint32_t *__restrict__ colIdxs,
                 const float *__restrict__ vals, uint32_t M, const uint32_t *__restrict__ slicePtrs,
                 uint32_t *__restrict__ outCols, float *__restrict__ outVals) {
    const uint32_t s = blockIdx.x;
    const uint32_t base = slicePtrs[s];
    const uint32_t width = (slicePtrs[s + 1] - base) / 32;
    const uint32_t i = threadIdx.x & 31, jj = threadIdx.x >> 5;
    const uint32_t r = s * 32 + i;
    uint32_t start = 0, len = 0;
    if (r < M) { start = rowPtrs[r]; len = rowPtrs[r + 1] - start; }
    for (uint32_t j = jj; j < width; j += 8) {
        uint32_t c = kPad;
        float v = 0.f;
        if (j < len) { c = colIdxs[start + j]; v = vals[start + j]; }
        outCols[base + (size_t)j * 32 + i] = c;      // coalesced: 32 rows' j-th entries are adjacent
        outVals[base + (size_t)j * 32 + i] = v;
    }
}

// ------------------------------------------------------------------ CSR -> BSR
__global__ void bsr_keys_kernel(const uint32_t *__restrict__ rowPtrs, const uint32_t *__restrict__ colIdxs,
                                uint32_t M, uint32_t br, uint32_t bc, uint32_t nbc,
                                uint64_t *__restrict__ keys, uint32_t *__restrict__ idx) {
    // one warp per row
    const uint32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= M) return;
    const uint32_t start = rowPtrs[r], end = rowPtrs[r + 1];
    const uint64_t R = r / br;
    for (uint32_t i = start + lane_id(); i < end; i += 32) {
        keys[i] = R * nbc + colIdxs[i] / bc;
        idx[i] = i;
    }
}

// flag[i] = 1 where sorted key i starts a new block; also counts blocks per block row
__global__ void bsr_flag_kernel(const uint64_t *__restrict__ keys, uint32_t nnz, uint32_t nbc,
                                uint32_t *__restrict__ flags, uint32_t *__restrict__ rowCounts) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const bool first = (i == 0) || (keys[i] != keys[i - 1]);
    flags[i] = first ? 1u : 0u;
    if (first && rowCounts) atomicAdd(rowCounts + (uint32_t)(keys[i] / nbc), 1u);   // integer count: order-free
}

// multi-block exclusive scan of flags (nnz can be 10^8): per-block sums, 1-block scan, add back
__global__ void __launch_bounds__(1024) block_sums_kernel(const uint32_t *__restrict__ in, uint64_t n,
                                                         uint32_t *__restrict__ sums) {
    __shared__ uint32_t ws[32];
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t v = (i < n) ? in[i] : 0u;
    v = __reduce_add_sync(0xFFFFFFFFu, v);
    if (lane_id() == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t s = ws[threadIdx.x];
        s = __reduce_add_sync(0xFFFFFFFFu, s);
        if (threadIdx.x == 0) sums[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(1024) block_scan_kernel(const uint32_t *__restrict__ in, uint64_t n,
                                                         const uint32_t *__restrict__ blockOffsets,
                                                         uint32_t *__restrict__ out) {
    __shared__ uint32_t ws[32];
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t v = (i < n) ? in[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d);
        if (lane >= (uint32_t)d) x += y;
    }
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t s = ws[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xFFFFFFFFu, s, d);
            if (lane >= (uint32_t)d) s += y;
        }
        ws[lane] = s;
    }
    __syncthreads();
    if (i < n) out[i] = x - v + (warp ? ws[warp - 1] : 0u) + blockOffsets[blockIdx.x];
}

__global__ void bsr_scatter_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ idx,
                                   const uint32_t *__restrict__ flags, const uint32_t *__restrict__ blockOf,
                                   const uint32_t *__restrict__ rowPtrs, const uint32_t *__restrict__ colIdxs,
                                   const float *__restrict__ vals, uint32_t M, uint32_t nnz,
                                   uint32_t br, uint32_t bc, uint32_t nbc,
                                   uint32_t *__restrict__ blockColIdxs, float *__restrict__ blocks) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    // blockOf is the EXCLUSIVE scan of flags: block id of sorted entry i = blockOf[i] + flags[i] - 1
    const uint32_t b = blockOf[i] + flags[i] - 1u;
    const uint64_t key = keys[i];
    if (flags[i]) blockColIdxs[b] = (uint32_t)(key % nbc);
    const uint32_t src = idx[i];
    const uint32_t c = colIdxs[src];
    // row of the source entry: binary search in rowPtrs (src is an index into the CSR arrays)
    uint32_t lo = 0, hi = M;             // largest r with rowPtrs[r] <= src
    while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (rowPtrs[mid] <= src) lo = mid; else hi = mid;
    }
    // skip empty rows that share the same pointer: lo is the LAST row with rowPtrs[lo] <= src,
    // and rowPtrs[lo+1] > src, so src belongs to row lo.
    const uint32_t r = lo;
    blocks[(size_t)b * br * bc + (size_t)(r % br) * bc + (c % bc)] = vals[src];
}

// ------------------------------------------------------------------ CSR -> BSR without a sort (per-block-row bitmap)
// The entries of a block row are one contiguous range of the CSR arrays (br consecutive rows), so no global sort is
// needed: a CTA per block row marks the block columns it meets in a shared-memory bitmap (nbc bits); the bitmap IS the
// ascending blockColIdxs of that block row, and "number of set bits below bit cb" is the position of block cb inside it.
//   count: blocks per block row = popcount of the bitmap
//   fill : rebuild the bitmap, prefix-popcount its words, emit blockColIdxs, scatter every entry into its block
// Reads colIdxs twice and vals once, writes the blocks once: HBM-bound streaming instead of two 64-bit-key radix sorts of
// all non-zeros (the sort path stays for matrices whose bitmap would not fit into shared memory).
constexpr uint32_t kBsrBitmapMaxWords = 12288;      // x2 arrays (bits + prefix) = 96 KB: nbc <= 393 216 block columns

template <bool FILL>
__global__ void __launch_bounds__(256)
bsr_bitmap_kernel(const uint32_t *__restrict__ rowPtrs, const uint32_t *__restrict__ colIdxs, const float *__restrict__ vals,
                  uint32_t M, uint32_t br, uint32_t bc, uint32_t words, uint32_t *__restrict__ counts,
                  const uint32_t *__restrict__ blockRowPtrs, uint32_t *__restrict__ blockColIdxs, float *__restrict__ blocks) {
    extern __shared__ uint32_t bm[];                 // [words] bits, then (FILL) [words] exclusive prefix popcounts
    __shared__ uint32_t red[8];
    uint32_t *pre = bm + words;
    const uint32_t R = blockIdx.x;
    const uint32_t rBeg = R * br, rEnd = min(M, rBeg + br);
    const uint32_t i0 = __ldg(rowPtrs + rBeg), i1 = __ldg(rowPtrs + rEnd);
    for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) bm[w] = 0u;
    __syncthreads();
    for (uint32_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const uint32_t cb = __ldg(colIdxs + i) / bc;
        atomicOr(bm + (cb >> 5), 1u << (cb & 31u));
    }
    __syncthreads();
    // popcount of the bitmap, as exclusive prefix per word when filling: each thread owns a contiguous run of words
    const uint32_t per = (words + blockDim.x - 1) / blockDim.x;
    const uint32_t w0 = min(words, threadIdx.x * per), w1 = min(words, w0 + per);
    uint32_t mine = 0;
    for (uint32_t w = w0; w < w1; ++w) mine += (uint32_t)__popc(bm[w]);
    uint32_t incl = mine;                            // inclusive scan over the 256 threads
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane_id() >= (uint32_t)d) incl += y;
    }
    if (lane_id() == 31) red[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t warpBase = 0, total = 0;
    for (uint32_t w = 0; w < (blockDim.x >> 5); ++w) {
        if (w < (threadIdx.x >> 5)) warpBase += red[w];
        total += red[w];
    }
    if constexpr (!FILL) {
        if (threadIdx.x == 0) counts[R] = total;
        return;
    } else {
        uint32_t run = warpBase + incl - mine;
        for (uint32_t w = w0; w < w1; ++w) {
            pre[w] = run;
            run += (uint32_t)__popc(bm[w]);
        }
        __syncthreads();
        const uint32_t base = __ldg(blockRowPtrs + R);
        for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) {        // ascending block columns of this block row
            uint32_t bits = bm[w], k = pre[w];
            while (bits) {
                const uint32_t bit = (uint32_t)__ffs((int)bits) - 1u;
                bits &= bits - 1u;
                blockColIdxs[base + k++] = w * 32u + bit;
            }
        }
        for (uint32_t r = rBeg; r < rEnd; ++r) {                            // every entry into its block (blocks are pre-zeroed)
            const uint32_t p0 = __ldg(rowPtrs + r), p1 = __ldg(rowPtrs + r + 1);
            for (uint32_t i = p0 + threadIdx.x; i < p1; i += blockDim.x) {
                const uint32_t c = __ldg(colIdxs + i), cb = c / bc;
                const uint32_t blk = base + pre[cb >> 5] + (uint32_t)__popc(bm[cb >> 5] & ((1u << (cb & 31u)) - 1u));
                blocks[(size_t)blk * br * bc + (size_t)(r - rBeg) * bc + (c - cb * bc)] = ld_stream(vals + i);
            }
        }
    }
}

static bool bsr_bitmap_fits(uint32_t nbc) {
    static const bool forceSort = getenv("CUSPMM_BSR_CONVERT_SORT") != nullptr;       // tuning hook: the radix-sort path
    return !forceSort && (nbc + 31) / 32 <= kBsrBitmapMaxWords;
}

// ------------------------------------------------------------------ column-ELL -> CSR
__global__ void iota_kernel(uint32_t *idx, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) idx[i] = (uint32_t)i;
}

__global__ void count_nonpad_kernel(const uint32_t *__restrict__ rows, uint64_t n, unsigned long long *count) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t v = (i < n && rows[i] != kPad) ? 1u : 0u;
    v = __reduce_add_sync(0xFFFFFFFFu, v);
    if (lane_id() == 0 && v) atomicAdd(count, (unsigned long long)v);
}

__global__ void colell_gather_kernel(const uint32_t *__restrict__ sortedIdx, const float *__restrict__ ellVals,
                                     uint32_t W, uint32_t nnz, uint32_t *__restrict__ colIdxs,
                                     float *__restrict__ vals) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const uint32_t slot = sortedIdx[i];
    colIdxs[i] = slot / W;
    vals[i] = ellVals[slot];
}

// ------------------------------------------------------------------ dense transpose
__global__ void __launch_bounds__(256) transpose_kernel(const float *__restrict__ in, uint32_t rows, uint32_t cols,
                                                        float *__restrict__ out) {
    __shared__ float tile[32][33];
    const uint32_t c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (uint32_t j = ty; j < 32; j += 8)
        if (r0 + j < rows && c0 + tx < cols) tile[j][tx] = in[(size_t)(r0 + j) * cols + c0 + tx];
    __syncthreads();
    for (uint32_t j = ty; j < 32; j += 8)
        if (c0 + j < cols && r0 + tx < rows) out[(size_t)(c0 + j) * rows + r0 + tx] = tile[tx][j];
}

struct AsyncBuf {   // stream-ordered temporary
    void *p = nullptr;
    cudaStream_t st;
    explicit AsyncBuf(cudaStream_t s) : st(s) {}
    cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes ? bytes : 1, st); }
    ~AsyncBuf() { if (p) cudaFreeAsync(p, st); }
    template <class T> T *as() { return static_cast<T *>(p); }
};

static int sort_bsr_keys(const uint32_t *rowPtrs, const uint32_t *colIdxs, uint32_t M, uint32_t nnz,
                         uint32_t br, uint32_t bc, uint32_t nbr, uint32_t nbc, AsyncBuf &keysOut, AsyncBuf &idxOut,
                         cudaStream_t st) {
    AsyncBuf keysIn(st), idxIn(st), tmp(st);
    CUSPMM_CUDA(keysIn.alloc((size_t)nnz * 8));
    CUSPMM_CUDA(idxIn.alloc((size_t)nnz * 4));
    CUSPMM_CUDA(keysOut.alloc((size_t)nnz * 8));
    CUSPMM_CUDA(idxOut.alloc((size_t)nnz * 4));
    bsr_keys_kernel<<<(M + 7) / 8, 256, 0, st>>>(rowPtrs, colIdxs, M, br, bc, nbc, keysIn.as<uint64_t>(), idxIn.as<uint32_t>());
    CUSPMM_LAUNCH_CHECK("bsr_keys_kernel");
    int bits = 1;
    while (bits < 64 && (((uint64_t)nbr * nbc - 1) >> bits)) ++bits;
    size_t tmpBytes = 0;
    CUSPMM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keysIn.as<uint64_t>(), keysOut.as<uint64_t>(),
                                                idxIn.as<uint32_t>(), idxOut.as<uint32_t>(), (int)nnz, 0, bits, st));
    CUSPMM_CUDA(tmp.alloc(tmpBytes));
    CUSPMM_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmpBytes, keysIn.as<uint64_t>(), keysOut.as<uint64_t>(),
                                                idxIn.as<uint32_t>(), idxOut.as<uint32_t>(), (int)nnz, 0, bits, st));
    count_launch();
    return CUSPMM_OK;
}

} // namespace cuspmm_b200

using namespace cuspmm_b200;

extern "C" int cuspmm_transpose_f32(const float *in, uint32_t rows, uint32_t cols, float *out, void *stream) {
    CUSPMM_REQUIRE(in && out, "null pointer");
    if (!rows || !cols) return CUSPMM_OK;
    dim3 grid((cols + 31) / 32, (rows + 31) / 32);
    transpose_kernel<<<grid, 256, 0, as_stream(stream)>>>(in, rows, cols, out);
    CUSPMM_LAUNCH_CHECK("transpose_kernel");
    return CUSPMM_OK;
}

extern "C" int cuspmm_coo_to_csr_rowptrs(const uint32_t *rowIdxs, uint32_t M, uint32_t nnz, uint32_t *rowPtrs,
                                         void *stream) {
    CUSPMM_REQUIRE(rowPtrs && (nnz == 0 || rowIdxs), "null pointer");
    return coo_to_csr_rowptrs(rowIdxs, M, nnz, rowPtrs, as_stream(stream));
}

extern "C" int cuspmm_csr_check_sorted(const uint32_t *rowPtrs, const uint32_t *colIdxs, uint32_t M, uint32_t K,
                                       uint32_t *bad_rows_host, void *stream) {
    CUSPMM_REQUIRE(bad_rows_host && (M == 0 || (rowPtrs && colIdxs)), "null pointer");
    *bad_rows_host = 0;
    if (M == 0) return CUSPMM_OK;
    cudaStream_t st = as_stream(stream);
    unsigned int *bad = nullptr;
    CUSPMM_CUDA(cudaMallocAsync(&bad, sizeof(unsigned int), st));
    CUSPMM_CUDA(cudaMemsetAsync(bad, 0, sizeof(unsigned int), st));
    csr_check_sorted_kernel<<<(M + 7) / 8, 256, 0, st>>>(rowPtrs, colIdxs, M, K, bad);
    CUSPMM_LAUNCH_CHECK("csr_check_sorted_kernel");
    CUSPMM_CUDA(cudaMemcpyAsync(bad_rows_host, bad, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    CUSPMM_CUDA(cudaFreeAsync(bad, st));
    CUSPMM_CUDA(cudaStreamSynchronize(st));
    return CUSPMM_OK;
}

extern "C" int cuspmm_partition_rows_by_nnz(const uint32_t *rowPtrs, uint32_t M, uint32_t nnz, uint32_t parts,
                                            uint32_t *splits_host, void *stream) {
    CUSPMM_REQUIRE(parts >= 1 && splits_host && rowPtrs, "bad arguments");
    cudaStream_t st = as_stream(stream);
    AsyncBuf d(st);
    CUSPMM_CUDA(d.alloc((size_t)(parts + 1) * 4));
    partition_kernel<<<(parts + 1 + 7) / 8, 256, 0, st>>>(rowPtrs, M, nnz, parts, d.as<uint32_t>());
    CUSPMM_LAUNCH_CHECK("partition_kernel");
    CUSPMM_CUDA(cudaMemcpyAsync(splits_host, d.p, (size_t)(parts + 1) * 4, cudaMemcpyDeviceToHost, st));
    CUSPMM_CUDA(cudaStreamSynchronize(st));
    return CUSPMM_OK;
}

extern "C" int cuspmm_csr_to_sell_count(const uint32_t *rowPtrs, uint32_t M, uint32_t sliceH, uint32_t *slicePtrs,
                                        uint32_t *slots_host, void *stream) {
    CUSPMM_REQUIRE(sliceH == 32, "slice height must be 32");
    CUSPMM_REQUIRE(rowPtrs && slicePtrs && slots_host, "null pointer");
    cudaStream_t st = as_stream(stream);
    const uint32_t slices = (M + 31) / 32;
    AsyncBuf widths(st);
    CUSPMM_CUDA(widths.alloc((size_t)slices * 4));
    if (slices) {
        sell_width_kernel<<<(slices + 7) / 8, 256, 0, st>>>(rowPtrs, M, widths.as<uint32_t>());
        CUSPMM_LAUNCH_CHECK("sell_width_kernel");
    }
    scan_exclusive_1block<<<1, 1024, 0, st>>>(widths.as<uint32_t>(), slicePtrs, slices, 32u);
    CUSPMM_LAUNCH_CHECK("scan_exclusive_1block");
    CUSPMM_CUDA(cudaMemcpyAsync(slots_host, slicePtrs + slices, 4, cudaMemcpyDeviceToHost, st));
    CUSPMM_CUDA(cudaStreamSynchronize(st));
    return CUSPMM_OK;
}

extern "C" int cuspmm_csr_to_sell_fill(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals,
                                       uint32_t M, uint32_t sliceH, const uint32_t *slicePtrs,
                                       uint32_t *sellCols, float *sellVals, void *stream) {
    CUSPMM_REQUIRE(sliceH == 32, "slice height must be 32");
    const uint32_t slices = (M + 31) / 32;
    if (!slices) return CUSPMM_OK;
    sell_fill_kernel<<<slices, 256, 0, as_stream(stream)>>>(rowPtrs, colIdxs, vals, M, slicePtrs, sellCols, sellVals);
    CUSPMM_LAUNCH_CHECK("sell_fill_kernel");
    return CUSPMM_OK;
}

extern "C" int cuspmm_csr_to_bsr_count(const uint32_t *rowPtrs, const uint32_t *colIdxs, uint32_t M, uint32_t K,
                                       uint32_t nnz, uint32_t br, uint32_t bc, uint32_t *blockRowPtrs,
                                       uint32_t *numBlocks_host, void *stream) {
    CUSPMM_REQUIRE(br && bc && blockRowPtrs && numBlocks_host, "bad arguments");
    cudaStream_t st = as_stream(stream);
    const uint32_t nbr = (M + br - 1) / br, nbc = (K + bc - 1) / bc;
    AsyncBuf counts(st);
    CUSPMM_CUDA(counts.alloc((size_t)(nbr + 1) * 4));
    CUSPMM_CUDA(cudaMemsetAsync(counts.p, 0, (size_t)(nbr + 1) * 4, st));
    if (nnz && bsr_bitmap_fits(nbc)) {
        const uint32_t words = (nbc + 31) / 32;
        auto kern = bsr_bitmap_kernel<false>;
        CUSPMM_CUDA(set_smem_once(kern, (size_t)words * 4));
        kern<<<nbr, 256, (size_t)words * 4, st>>>(rowPtrs, colIdxs, nullptr, M, br, bc, words, counts.as<uint32_t>(), nullptr, nullptr, nullptr);
        CUSPMM_LAUNCH_CHECK("bsr_bitmap_kernel<count>");
    } else if (nnz) {
        AsyncBuf keys(st), idx(st), flags(st);
        int rc = sort_bsr_keys(rowPtrs, colIdxs, M, nnz, br, bc, nbr, nbc, keys, idx, st);
        if (rc) return rc;
        CUSPMM_CUDA(flags.alloc((size_t)nnz * 4));
        bsr_flag_kernel<<<(unsigned)(((uint64_t)nnz + 255) / 256), 256, 0, st>>>(keys.as<uint64_t>(), nnz, nbc,
                                                                                flags.as<uint32_t>(), counts.as<uint32_t>());
        CUSPMM_LAUNCH_CHECK("bsr_flag_kernel");
    }
    scan_exclusive_1block<<<1, 1024, 0, st>>>(counts.as<uint32_t>(), blockRowPtrs, nbr, 1u);
    CUSPMM_LAUNCH_CHECK("scan_exclusive_1block");
    CUSPMM_CUDA(cudaMemcpyAsync(numBlocks_host, blockRowPtrs + nbr, 4, cudaMemcpyDeviceToHost, st));
    CUSPMM_CUDA(cudaStreamSynchronize(st));
    return CUSPMM_OK;
}

extern "C" int cuspmm_csr_to_bsr_fill(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals,
                                      uint32_t M, uint32_t K, uint32_t nnz, uint32_t br, uint32_t bc,
                                      uint32_t numBlocks, uint32_t *blockColIdxs, float *blocks, void *stream) {
    CUSPMM_REQUIRE(br && bc, "bad block shape");
    cudaStream_t st = as_stream(stream);
    const uint32_t nbr = (M + br - 1) / br, nbc = (K + bc - 1) / bc;
    if (numBlocks) CUSPMM_CUDA(cudaMemsetAsync(blocks, 0, (size_t)numBlocks * br * bc * sizeof(float), st));
    if (!nnz) return CUSPMM_OK;
    if (bsr_bitmap_fits(nbc)) {
        // block row pointers again (the caller keeps only numBlocks between the two calls): count + scan, then fill
        const uint32_t words = (nbc + 31) / 32;
        AsyncBuf counts(st), brp(st);
        CUSPMM_CUDA(counts.alloc((size_t)(nbr + 1) * 4));
        CUSPMM_CUDA(brp.alloc((size_t)(nbr + 1) * 4));
        auto kc = bsr_bitmap_kernel<false>;
        auto kf = bsr_bitmap_kernel<true>;
        CUSPMM_CUDA(set_smem_once(kc, (size_t)words * 4));
        CUSPMM_CUDA(set_smem_once(kf, (size_t)words * 8));
        kc<<<nbr, 256, (size_t)words * 4, st>>>(rowPtrs, colIdxs, nullptr, M, br, bc, words, counts.as<uint32_t>(), nullptr, nullptr, nullptr);
        CUSPMM_LAUNCH_CHECK("bsr_bitmap_kernel<count>");
        scan_exclusive_1block<<<1, 1024, 0, st>>>(counts.as<uint32_t>(), brp.as<uint32_t>(), nbr, 1u);
        CUSPMM_LAUNCH_CHECK("scan_exclusive_1block");
        kf<<<nbr, 256, (size_t)words * 8, st>>>(rowPtrs, colIdxs, vals, M, br, bc, words, nullptr, brp.as<uint32_t>(), blockColIdxs, blocks);
        CUSPMM_LAUNCH_CHECK("bsr_bitmap_kernel<fill>");
        return CUSPMM_OK;
    }
    AsyncBuf keys(st), idx(st), flags(st), blockOf(st), sums(st), offs(st);
    int rc = sort_bsr_keys(rowPtrs, colIdxs, M, nnz, br, bc, nbr, nbc, keys, idx, st);
    if (rc) return rc;
    CUSPMM_CUDA(flags.alloc((size_t)nnz * 4));
    CUSPMM_CUDA(blockOf.alloc((size_t)nnz * 4));
    const unsigned nb = (unsigned)(((uint64_t)nnz + 1023) / 1024);
    CUSPMM_CUDA(sums.alloc((size_t)nb * 4));
    CUSPMM_CUDA(offs.alloc((size_t)(nb + 1) * 4));
    bsr_flag_kernel<<<(unsigned)(((uint64_t)nnz + 255) / 256), 256, 0, st>>>(keys.as<uint64_t>(), nnz, nbc,
                                                                            flags.as<uint32_t>(), nullptr);
    CUSPMM_LAUNCH_CHECK("bsr_flag_kernel");
    block_sums_kernel<<<nb, 1024, 0, st>>>(flags.as<uint32_t>(), nnz, sums.as<uint32_t>());
    CUSPMM_LAUNCH_CHECK("block_sums_kernel");
    scan_exclusive_1block<<<1, 1024, 0, st>>>(sums.as<uint32_t>(), offs.as<uint32_t>(), nb, 1u);
    CUSPMM_LAUNCH_CHECK("scan_exclusive_1block");
    block_scan_kernel<<<nb, 1024, 0, st>>>(flags.as<uint32_t>(), nnz, offs.as<uint32_t>(), blockOf.as<uint32_t>());
    CUSPMM_LAUNCH_CHECK("block_scan_kernel");
    bsr_scatter_kernel<<<(unsigned)(((uint64_t)nnz + 255) / 256), 256, 0, st>>>(
        keys.as<uint64_t>(), idx.as<uint32_t>(), flags.as<uint32_t>(), blockOf.as<uint32_t>(), rowPtrs, colIdxs, vals,
        M, nnz, br, bc, nbc, blockColIdxs, blocks);
    CUSPMM_LAUNCH_CHECK("bsr_scatter_kernel");
    return CUSPMM_OK;
}

extern "C" int cuspmm_colell_to_csr(const uint32_t *ellRowIdxs, const float *ellVals, uint32_t M, uint32_t K,
                                    uint32_t W, uint32_t nnz, uint32_t *rowPtrs, uint32_t *colIdxs, float *vals,
                                    void *stream) {
    CUSPMM_REQUIRE(rowPtrs, "null rowPtrs");
    cudaStream_t st = as_stream(stream);
    const uint64_t slots = (uint64_t)K * W;
    CUSPMM_REQUIRE(slots <= 0x7FFFFFFFull, "column-ELL with %llu slots is too large for the device sort",
                   (unsigned long long)slots);
    if (slots == 0 || nnz == 0) {
        CUSPMM_CUDA(cudaMemsetAsync(rowPtrs, 0, (size_t)(M + 1) * 4, st));
        CUSPMM_CUDA(cudaStreamSynchronize(st));
        return CUSPMM_OK;
    }
    AsyncBuf idxIn(st), idxOut(st), keysOut(st), tmp(st), cnt(st);
    CUSPMM_CUDA(idxIn.alloc(slots * 4));
    CUSPMM_CUDA(idxOut.alloc(slots * 4));
    CUSPMM_CUDA(keysOut.alloc(slots * 4));
    CUSPMM_CUDA(cnt.alloc(8));
    CUSPMM_CUDA(cudaMemsetAsync(cnt.p, 0, 8, st));
    const unsigned gb = (unsigned)((slots + 255) / 256);
    iota_kernel<<<gb, 256, 0, st>>>(idxIn.as<uint32_t>(), slots);
    CUSPMM_LAUNCH_CHECK("iota_kernel");
    count_nonpad_kernel<<<gb, 256, 0, st>>>(ellRowIdxs, slots, cnt.as<unsigned long long>());
    CUSPMM_LAUNCH_CHECK("count_nonpad_kernel");
    // stable LSD radix sort by row: inside a row the slots stay in ascending (col, slot) order.
    // Padding (0xFFFFFFFF) sorts to the end.
    size_t tmpBytes = 0;
    CUSPMM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, ellRowIdxs, keysOut.as<uint32_t>(),
                                                idxIn.as<uint32_t>(), idxOut.as<uint32_t>(), (int)slots, 0, 32, st));
    CUSPMM_CUDA(tmp.alloc(tmpBytes));
    CUSPMM_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmpBytes, ellRowIdxs, keysOut.as<uint32_t>(),
                                                idxIn.as<uint32_t>(), idxOut.as<uint32_t>(), (int)slots, 0, 32, st));
    count_launch();
    unsigned long long found = 0;
    CUSPMM_CUDA(cudaMemcpyAsync(&found, cnt.p, 8, cudaMemcpyDeviceToHost, st));
    CUSPMM_CUDA(cudaStreamSynchronize(st));
    CUSPMM_REQUIRE(found == nnz, "column-ELL holds %llu non-padding slots but the header says nnz = %u", found, nnz);
    int rc = coo_to_csr_rowptrs(keysOut.as<uint32_t>(), M, nnz, rowPtrs, st);
    if (rc) return rc;
    colell_gather_kernel<<<(unsigned)(((uint64_t)nnz + 255) / 256), 256, 0, st>>>(idxOut.as<uint32_t>(), ellVals, W, nnz,
                                                                                 colIdxs, vals);
    CUSPMM_LAUNCH_CHECK("colell_gather_kernel");
    CUSPMM_CUDA(cudaStreamSynchronize(st));
    return CUSPMM_OK;
}
