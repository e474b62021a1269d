/*
 * cuspmm_b200.h -- C ABI of the B200-native SpMM engine (libcuspmm_b200.so).
 *
 * C = A * B, A sparse (CSR / COO / ELL / BSR), B and C dense row-major fp32.
 * Every entry point is what the reference's host layer would bind for this path;
 * the reference interface each one replaces is cited as file:line relative to
 * mli43/Cuda-Optimization-for-SpMM.  Conventions:
 *
 *   - plain pointers and sizes only; no C++ types, no exceptions cross the ABI;
 *   - all `*_dev` pointers are DEVICE pointers on the current device; work is
 *     enqueued on `stream` (a cudaStream_t passed as void*, NULL = legacy default
 *     stream) and the call returns without synchronising unless stated;
 *   - C is OVERWRITTEN (beta = 0); it does not have to be zeroed first;
 *   - indices are uint32 (the reference's MT), values float (its DT);
 *   - ldb / ldc are row strides in ELEMENTS (>= N);
 *   - return value: 0 = CUSPMM_OK, otherwise a cuspmmStatus; the message is in
 *     cuspmm_last_error() (thread local).  There is NO CPU fallback: without a
 *     usable CUDA device every compute entry point fails with CUSPMM_ERR_CUDA.
 */
#ifndef CUSPMM_B200_H
#define CUSPMM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CUSPMM_B200_VERSION 210

typedef enum {
    CUSPMM_OK = 0,
    CUSPMM_ERR_INVALID = 1,     /* bad argument / shape / alignment */
    CUSPMM_ERR_CUDA = 2,        /* CUDA runtime or driver error */
    CUSPMM_ERR_UNSUPPORTED = 3, /* variant cannot run this shape (cf. spmm_csr_k4.cu:97-101) */
    CUSPMM_ERR_WORKSPACE = 4,   /* workspace too small */
    CUSPMM_ERR_CUSPARSE = 5     /* vendor baseline failed */
} cuspmmStatus;

/* BSR tensor-core block value types */
typedef enum { CUSPMM_BLK_BF16 = 0, CUSPMM_BLK_FP16 = 1 } cuspmmBlockType;

/* ------------------------------------------------------------------ misc ---- */
int cuspmm_version(void);
const char *cuspmm_last_error(void);
/* Number of launches of THIS library's kernels issued by the calling thread since
 * the last cuspmm_reset_launch_count() (bench.py's gpu_launches). */
unsigned long long cuspmm_launch_count(void);
void cuspmm_reset_launch_count(void);
int cuspmm_device_count(int *count);
int cuspmm_device_info(int device, int *sm_count, int *cc_major, int *cc_minor,
                       size_t *l2_bytes, size_t *total_mem_bytes);

/* ------------------------------------------------------------------- CSR ---- */
/* Replaces spmmCSRWrapper1..4 / spmmCSRK1..4 (include/engine/engine_csr.hpp:15-25,
 * src/spmm/csr/spmm_csr_k{1,2,3,4}.cu).  Variants (Engine<CSR>::runKernel numbers):
 *   0  auto (selector: DESIGN.md "kernel selection")
 *   1  row-split, warp per row, 128-bit B loads, nnz-balanced row ranges
 *   2  vector-per-row: sub-warp per row for narrow N / short rows
 *   3  staged: row panel x K-chunks, B tiles staged in shared memory by TMA bulk copies
 *   4  scalar generic (any N, any alignment)
 *   5  staged, dual operand path: as 3, and the first rows of every B chunk are copied on
 *      into tensor memory (tcgen05.cp) and gathered from there with tcgen05.ld, which
 *      takes those non-zeros off the shared-memory pipe; fp32 FMA in CSR order like 1..4
 *      (N % 512 == 0, else CUSPMM_ERR_UNSUPPORTED)
 *   6  nnz split that cuts rows between warps (merge-path style) with an ordered carry fix-up, for few / skewed rows;
 *      no atomics, fixed order; cut rows are rounded differently from 1..5 (partial sums first), uncut rows identically.
 *      Needs workspace: only through cuspmm_spmm_csr_ws.
 *   7  every B read from tensor memory: TMA -> shared-memory ring -> tcgen05.cp.128x256b -> a 128-row TMEM ring of B; a warp
 *      owns 8 rows x the 128 columns of its TMEM lane quarter and fetches B with tcgen05.ld.x4 (no shared-memory operand
 *      reads at all); fp32 FMA in CSR order like 1..5 (N % 512 == 0, else CUSPMM_ERR_UNSUPPORTED)
 *   8  tensor cores: tiles of A (512 rows x 16 columns per CTA pair) are made dense in shared memory and multiplied with
 *      tcgen05.mma (cta_group::2) against a pre-tiled copy of B (built per call, in a stream-ordered pool allocation of 8
 *      bytes per element of B).  fp32-grade result from a three-product split: tf32(a)*tf32(b) + bf16(a)*bf16(b - tf32(b)) +
 *      bf16(a - tf32(a))*bf16(b), fp32 accumulation in tensor memory, drained into C every 64 chunks of 16 columns of A by
 *      reductions that round to nearest (the tensor core truncates when it accumulates); per-product error <= 2^-17 |a||b|
 *      (7.6e-6) in the worst case, ~5e-7 of sum|a||b| on long sums.  Work does not depend on nnz (M*K*N*2 tensor MACs): it
 *      wins from ~5 % density upwards (1.85x the fp32 kernels at 10 %, 3.3x at 50 % on 25605^2 x 512).  Any N, any alignment.
 *      If B holds a non-finite value (a dense product would spread it to rows that never reference it) a device-side flag
 *      reroutes the call to a plain fp32 kernel without host synchronisation.  C is cleared and then accumulated into
 *      (memset + reductions): it must not alias B, and tiles that two CTA pairs share (the last tiles of a grid that is not
 *      a multiple of the SM count) differ run to run in the last bits.
 *
 * PRECONDITION for variants 0, 3, 5, 7, 8 (and everything built on them: COO variant 2, sliced ELL, the host-buffer and multi-GPU
 * entry points): column indices ascend strictly inside every row, as the reference's converter writes them
 * (convert_mtx.py:127-143, scipy CSR with sorted indices).  Variants 1, 2, 4, 6 accept any order.  cuspmm_csr_check_sorted
 * verifies it on the device; with the environment variable CUSPMM_CHECK_SORTED set, every CSR call that is about to run a
 * staged kernel performs that check first and fails with CUSPMM_ERR_INVALID on unsorted rows (a debug guard: it costs a pass
 * over colIdxs and a stream synchronisation). */
#define CUSPMM_CSR_NUM_VARIANTS 8
int cuspmm_spmm_csr(const uint32_t *rowPtrs_dev, const uint32_t *colIdxs_dev, const float *vals_dev,
                    uint32_t M, uint32_t K, uint32_t nnz,
                    const float *B_dev, uint32_t N, size_t ldb,
                    float *C_dev, size_t ldc, int variant, void *stream);
/* Tensor-core mode of the CSR selector (and of everything built on it: COO variant 0 / 2, the host-buffer and multi-GPU CSR
 * entries).  1 (default): variant 0 may resolve to variant 8 where it is expected to be faster (from ~5 % density on large
 * matrices; results then agree with the fp32 kernels to ~5e-7 of sum|a||b|, not bit for bit).  0: fp32 FMA kernels only, every
 * variant 0 result bit-identical to variants 1..5.  The environment variable CUSPMM_TENSOR=0 sets the initial mode to 0.
 * Process-wide; returns the previous mode. */
int cuspmm_set_csr_tensor_mode(int mode);
/* The kernel variant 0 resolves to for this shape on the current device, assuming 16-byte aligned operands (sliced_ell != 0:
 * the same question for cuspmm_spmm_sell, answered in CSR variant numbers).  bench.py records it beside every measurement. */
int cuspmm_csr_selected_variant(uint32_t M, uint32_t K, uint32_t nnz, uint32_t N, int sliced_ell);
/* The same with caller-provided device workspace (16-byte aligned, cuspmm_spmm_csr_workspace bytes; 0 for variants
 * 1..5): variant 6 runs, and variant 0 may select it (few rows with >= 16 non-zeros each). */
size_t cuspmm_spmm_csr_workspace(uint32_t M, uint32_t K, uint32_t nnz, uint32_t N, int variant);
int cuspmm_spmm_csr_ws(const uint32_t *rowPtrs_dev, const uint32_t *colIdxs_dev, const float *vals_dev,
                       uint32_t M, uint32_t K, uint32_t nnz,
                       const float *B_dev, uint32_t N, size_t ldb,
                       float *C_dev, size_t ldc, int variant,
                       void *workspace_dev, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------- COO ---- */
/* Replaces spmmCOOWrapper1 / spmmCOOK1 (include/engine/engine_coo.hpp:14-15,
 * src/spmm/coo/spmm_coo_k1.cu).  Entries must be sorted by (row, col) as the
 * reference's converter writes them (convert_mtx.py:181-185).  Variants:
 *   0 auto, 1 row-aligned nnz split (no atomics, no workspace),
 *   2 COO->CSR row pointers on device, then the staged CSR kernel
 *     (needs (M+1)*4 bytes of workspace). */
#define CUSPMM_COO_NUM_VARIANTS 2
size_t cuspmm_spmm_coo_workspace(uint32_t M, uint32_t nnz, uint32_t N, int variant);
int cuspmm_spmm_coo(const uint32_t *rowIdxs_dev, const uint32_t *colIdxs_dev, const float *vals_dev,
                    uint32_t M, uint32_t K, uint32_t nnz,
                    const float *B_dev, uint32_t N, size_t ldb,
                    float *C_dev, size_t ldc, int variant,
                    void *workspace_dev, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------- ELL ---- */
/* Sliced ELL (the engine's native ELL layout): slices of `sliceH` (=32) rows, slice
 * s is W_s slots wide, slot-major: entry j of row s*32+i at slicePtrs[s] + j*32 + i;
 * padding colIdx 0xFFFFFFFF / value 0.  Replaces spmmELLWrapper1/2 / spmmELLK1/2
 * (include/engine/engine_ell.hpp:15-19, src/spmm/ell/spmm_ell_k{1,2}.cu). */
#define CUSPMM_ELL_NUM_VARIANTS 6 /* 1 row kernels (warp / sub-warp per row); 2 staged (B tiles via TMA bulk
                                     copies); 3 slice per CTA with the slots staged through shared memory;
                                     4 staged with the dual operand path (CSR variant 5 on the sliced layout);
                                     5 every B read from tensor memory (CSR variant 7 on the sliced layout);
                                     6 tensor cores (CSR variant 8 on the sliced layout: fp32-grade, not bit-identical
                                     to 1..5; what variant 0 resolves to from ~5 % density, see cuspmm_set_csr_tensor_mode) */
int cuspmm_spmm_sell(const uint32_t *slicePtrs_dev, const uint32_t *colIdxs_dev, const float *vals_dev,
                     uint32_t M, uint32_t K, uint32_t sliceH, uint32_t numSlots /* = slicePtrs[numSlices] */,
                     const float *B_dev, uint32_t N, size_t ldb,
                     float *C_dev, size_t ldc, int variant, void *stream);

/* The reference's column-ELL storage (include/formats/sparse_ell.hpp:12-37:
 * rowIdxs/vals are [K x maxColNnz], padding row -1) -> CSR, on the device: a stable
 * radix sort of the slots by row index.  nnz is the count from the ELL header
 * (SparseMatrixELL::numNonZero); the call fails with CUSPMM_ERR_INVALID if the
 * number of non-padding slots differs.  Outputs: rowPtrs_dev[M+1], colIdxs_dev[nnz],
 * vals_dev[nnz] in (row, col) order.  Temporaries come from the stream-ordered
 * allocator (cudaMallocAsync); the call synchronises the stream. */
int cuspmm_colell_to_csr(const uint32_t *ellRowIdxs_dev, const float *ellVals_dev,
                         uint32_t M, uint32_t K, uint32_t maxColNnz, uint32_t nnz,
                         uint32_t *rowPtrs_dev, uint32_t *colIdxs_dev, float *vals_dev, void *stream);

/* CSR -> sliced ELL on the device (north_star (b)).  _count fills
 * slicePtrs_dev[numSlices+1] (numSlices = ceil(M/32)) and returns the total slot
 * count through *slots_host (synchronises); _fill writes padded colIdxs/vals. */
int cuspmm_csr_to_sell_count(const uint32_t *rowPtrs_dev, uint32_t M, uint32_t sliceH,
                             uint32_t *slicePtrs_dev, uint32_t *slots_host, void *stream);
int cuspmm_csr_to_sell_fill(const uint32_t *rowPtrs_dev, const uint32_t *colIdxs_dev, const float *vals_dev,
                            uint32_t M, uint32_t sliceH, const uint32_t *slicePtrs_dev,
                            uint32_t *sellCols_dev, float *sellVals_dev, void *stream);

/* ------------------------------------------------------------------- BSR ---- */
/* fp32 blocks, any block shape; bit-for-bit the summation order of spmmBSRCpu.
 * Replaces spmmBSRWrapper1 / spmmBSRK1 (include/engine/engine_bsr.hpp:15-16,
 * src/spmm/bsr/spmm_bsr_k1.cu).  M = numBlockRows * br. */
#define CUSPMM_BSR_NUM_VARIANTS 3 /* 1 fp32 SIMT, 2 bf16 tcgen05, 3 fp16 tcgen05 */
int cuspmm_spmm_bsr_f32(const uint32_t *blockRowPtrs_dev, const uint32_t *blockColIdxs_dev,
                        const float *blocks_dev, uint32_t numBlockRows, uint32_t br, uint32_t bc,
                        uint32_t K, const float *B_dev, uint32_t N, size_t ldb,
                        float *C_dev, size_t ldc, void *stream);

/* CSR -> BSR on the device (the reference leaves SparseMatrixBSR::fromDense
 * unimplemented, src/formats/sparse_bsr.cu:259; offline it is scipy tobsr).
 * M and K are zero-padded up to multiples of br / bc.  _count fills
 * blockRowPtrs_dev[ceil(M/br)+1] and returns numBlocks (synchronises); _fill writes
 * ascending blockColIdxs and zero-filled row-major fp32 blocks.  No sort: a CTA per block row marks the block columns of
 * its (contiguous) CSR range in a shared-memory bitmap, whose prefix popcount is every block's position; matrices with more
 * than 393 216 block columns fall back to a device radix sort of (blockRow, blockCol) keys.  Temporaries come from the
 * stream-ordered allocator.  Column indices inside a row may be in any order. */
int cuspmm_csr_to_bsr_count(const uint32_t *rowPtrs_dev, const uint32_t *colIdxs_dev,
                            uint32_t M, uint32_t K, uint32_t nnz, uint32_t br, uint32_t bc,
                            uint32_t *blockRowPtrs_dev, uint32_t *numBlocks_host, void *stream);
int cuspmm_csr_to_bsr_fill(const uint32_t *rowPtrs_dev, const uint32_t *colIdxs_dev, const float *vals_dev,
                           uint32_t M, uint32_t K, uint32_t nnz, uint32_t br, uint32_t bc,
                           uint32_t numBlocks, uint32_t *blockColIdxs_dev, float *blocks_dev, void *stream);

/* Tensor-core BSR (north_star: tcgen05 MMA, TMEM accumulators, TMA-fed, fp32
 * accumulate).  Square blocks of 16 or 32.  The plan owns device copies of the
 * blocks in bf16/fp16 laid out as UMMA core matrices and a scratch for the
 * converted B; `prepare_B` casts + re-tiles B (K x N fp32) once per B, `run`
 * multiplies.  Results equal the fp32 product of the ROUNDED operands up to fp32
 * accumulation order (tolerances: DESIGN.md "parity").
 * Lifetime / device: the plan BORROWS blockRowPtrs_dev and blockColIdxs_dev (they must stay valid and unchanged until
 * plan_destroy); blocks_dev is read only during plan_create.  A plan belongs to the device that was current at plan_create:
 * prepare_B and run fail with CUSPMM_ERR_INVALID when another device is current. */
typedef struct cuspmmBsrTcPlan_s *cuspmmBsrTcPlan;
int cuspmm_bsr_tc_plan_create(cuspmmBsrTcPlan *plan,
                              const uint32_t *blockRowPtrs_dev, const uint32_t *blockColIdxs_dev,
                              const float *blocks_dev, uint32_t numBlockRows, uint32_t numBlocks,
                              uint32_t blockSize, uint32_t K, uint32_t maxN,
                              cuspmmBlockType type, void *stream);
int cuspmm_bsr_tc_prepare_B(cuspmmBsrTcPlan plan, const float *B_dev, uint32_t N, size_t ldb, void *stream);
int cuspmm_bsr_tc_run(cuspmmBsrTcPlan plan, float *C_dev, size_t ldc, void *stream);
int cuspmm_bsr_tc_plan_destroy(cuspmmBsrTcPlan plan);

/* --------------------------------------------------------- partitioning ---- */
/* nnz-balanced contiguous row panels (north_star (b)/(c)): splits_host[0] = 0,
 * splits_host[parts] = M, splits_host[g] = first row r with rowPtrs[r] >= g*nnz/parts.
 * Runs a device binary search on rowPtrs and synchronises the stream. */
int cuspmm_partition_rows_by_nnz(const uint32_t *rowPtrs_dev, uint32_t M, uint32_t nnz,
                                 uint32_t parts, uint32_t *splits_host, void *stream);
/* COO -> CSR row pointers on the device (entries sorted by row). */
int cuspmm_coo_to_csr_rowptrs(const uint32_t *rowIdxs_dev, uint32_t M, uint32_t nnz,
                              uint32_t *rowPtrs_dev, void *stream);

/* Dense transpose on the device: out[c * rows + r] = in[r * cols + c].  Replaces the host
 * double loop of DenseMatrix::toOrdering (src/formats/dense.cu:140-191). */
int cuspmm_transpose_f32(const float *in_dev, uint32_t rows, uint32_t cols, float *out_dev, void *stream);

/* Verifies the precondition of the staged kernels: column indices strictly ascending inside every row and < K.
 * *bad_rows_host = number of rows that violate it (0 = fine).  Synchronises the stream. */
int cuspmm_csr_check_sorted(const uint32_t *rowPtrs_dev, const uint32_t *colIdxs_dev, uint32_t M, uint32_t K,
                            uint32_t *bad_rows_host, void *stream);

/* ------------------------------------------------- host-buffer entry points ---- */
/* What runEngine + spmm<FMT>Wrapper<k> do end to end (src/engine/engine.cpp:20-44,
 * src/spmm/csr/spmm_csr_k3.cu:59-105): operands in HOST memory, H2D, kernel, D2H of
 * C -- for every format, as the reference runs all four through one runEngine (src/engine/engine.cpp:63-80).  A is cut
 * into balanced panels (rows / slices / block rows) that are pipelined over three streams so the copies overlap the
 * kernels.  Host buffers should be pinned (cudaHostAlloc / cuspmm_host_alloc) for the copies to be asynchronous.
 * Synchronises before returning, on every exit path.  `device_ms` (optional) receives the device time of the whole
 * pipeline measured with CUDA events.  Device staging buffers are cached per device between calls and freed by
 * cuspmm_host_pipeline_release(device) (device < 0: all devices), which also trims the stream-ordered pools CSR variant 8 keeps
 * for its per-call copy of B.  CSR / ELL: column indices ascending inside a row. */
int cuspmm_spmm_csr_host(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals,
                         uint32_t M, uint32_t K, uint32_t nnz,
                         const float *B, uint32_t N, float *C, int variant, float *device_ms);
/* The same with B already resident on the current device (row stride ldb): the kernels wait for the work enqueued so far
 * on `b_ready_stream` (e.g. the stream an NCCL all-gather of B runs on).  Used when one matrix is sharded over several
 * processes: every rank uploads its own row panel of A, B arrives over NVLink. */
int cuspmm_spmm_csr_host_devB(const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals,
                              uint32_t M, uint32_t K, uint32_t nnz,
                              const float *B_dev, size_t ldb, void *b_ready_stream,
                              uint32_t N, float *C, int variant, float *device_ms);
/* COO (entries sorted by (row, col)); variant as cuspmm_spmm_coo (0 / 2: panel-wise device row pointers + CSR selector;
 * 1: the row-aligned COO kernel, one launch). */
int cuspmm_spmm_coo_host(const uint32_t *rowIdxs, const uint32_t *colIdxs, const float *vals,
                         uint32_t M, uint32_t K, uint32_t nnz,
                         const float *B, uint32_t N, float *C, int variant, float *device_ms);
/* Sliced ELL (layout above); panels end on slice boundaries, balanced by slots. */
int cuspmm_spmm_sell_host(const uint32_t *slicePtrs, const uint32_t *colIdxs, const float *vals,
                          uint32_t M, uint32_t K, uint32_t sliceH, uint32_t numSlots,
                          const float *B, uint32_t N, float *C, int variant, float *device_ms);
/* BSR, fp32 blocks on the host; C is (numBlockRows * br) x N.  variant 0 / 1: fp32 kernels, panels of block rows balanced
 * by blocks; 2 / 3: bf16 / fp16 tensor-core plan built on the device after the upload (cast + re-tiling of the blocks and of
 * B are inside the timed region), one launch. */
int cuspmm_spmm_bsr_host(const uint32_t *blockRowPtrs, const uint32_t *blockColIdxs, const float *blocks,
                         uint32_t numBlockRows, uint32_t br, uint32_t bc, uint32_t K,
                         const float *B, uint32_t N, float *C, int variant, float *device_ms);
int cuspmm_host_pipeline_release(int device);
int cuspmm_host_alloc(void **ptr, size_t bytes); /* pinned, cf. cudaMallocHost in src/formats/dense.cu:244 */
int cuspmm_host_free(void *ptr);

/* ------------------------------------------------------ multi-GPU engine ---- */
/* Row-panel data parallelism over `ngpus` devices of one node (north_star (c);
 * the reference is single-device, src/main.cu:176), for every format.  create_<fmt>(): A (HOST arrays) is split
 * into balanced contiguous row panels -- CSR: by non-zeros (split points from the device partitioner); COO: the same rule,
 * moved to row boundaries; sliced ELL: at slice boundaries, by slots; BSR: at block-row boundaries, by blocks -- and panel g
 * is uploaded to device devices[g];
 * set_B(): B (HOST) is uploaded to device 0 and replicated to the peers over
 * NVLink (cudaMemcpyPeerAsync);  run(): every device multiplies its panel on its
 * own stream (variant numbering of the format; BSR: 1 fp32, 2 bf16, 3 fp16 tensor cores); with gather != 0 each device's
 * kernel writes its C rows straight
 * into device 0's C through peer memory (no separate collective), otherwise C
 * stays sharded.  Device time = max over devices, CUDA events.  */
typedef struct cuspmmMgpuPlan_s *cuspmmMgpuPlan;
int cuspmm_mgpu_create_csr(cuspmmMgpuPlan *plan, int ngpus, const int *devices,
                           const uint32_t *rowPtrs, const uint32_t *colIdxs, const float *vals,
                           uint32_t M, uint32_t K, uint32_t nnz, uint32_t maxN);
int cuspmm_mgpu_create_coo(cuspmmMgpuPlan *plan, int ngpus, const int *devices,
                           const uint32_t *rowIdxs, const uint32_t *colIdxs, const float *vals,
                           uint32_t M, uint32_t K, uint32_t nnz, uint32_t maxN);
int cuspmm_mgpu_create_sell(cuspmmMgpuPlan *plan, int ngpus, const int *devices,
                            const uint32_t *slicePtrs, const uint32_t *colIdxs, const float *vals,
                            uint32_t M, uint32_t K, uint32_t sliceH, uint32_t numSlots, uint32_t maxN);
int cuspmm_mgpu_create_bsr(cuspmmMgpuPlan *plan, int ngpus, const int *devices,
                           const uint32_t *blockRowPtrs, const uint32_t *blockColIdxs, const float *blocks,
                           uint32_t numBlockRows, uint32_t br, uint32_t bc, uint32_t K, uint32_t maxN);
int cuspmm_mgpu_set_B(cuspmmMgpuPlan plan, const float *B, uint32_t N);
int cuspmm_mgpu_run(cuspmmMgpuPlan plan, int variant, int gather, int iters, float *max_device_ms);
int cuspmm_mgpu_get_splits(cuspmmMgpuPlan plan, uint32_t *splits /* ngpus+1, in rows of C */);
int cuspmm_mgpu_get_counts(cuspmmMgpuPlan plan, uint32_t *counts /* ngpus: non-zeros / slots / blocks per panel */);
int cuspmm_mgpu_get_C(cuspmmMgpuPlan plan, float *C /* host, M x N */);
int cuspmm_mgpu_destroy(cuspmmMgpuPlan plan);

/* ------------------------------------------------------ vendor baseline ---- */
/* cusparseSpMM, fp32 compute, row-major B and C, alpha 1 beta 0: the reference's
 * cusparseTest (src/engine/cusparse.cu:10-57) with its algorithm choices
 * (CSR_ALG2: sparse_csr.cu:183-185, COO_ALG4: sparse_coo.cu:98-100), timed with
 * CUDA events over `iters` launches after `warmup`; handle, descriptors, buffer and
 * cusparseSpMM_preprocess are outside the timed region.  fmt: 0 CSR, 1 COO.
 * alg: 0 = the reference's choice, otherwise a cusparseSpMMAlg_t value. */
int cuspmm_cusparse_spmm(int fmt, const uint32_t *rowOrPtr_dev, const uint32_t *colIdxs_dev,
                         const float *vals_dev, uint32_t M, uint32_t K, uint32_t nnz,
                         const float *B_dev, uint32_t N, float *C_dev,
                         int alg, int warmup, int iters, float *avg_ms, float *min_ms);

/* cuSPARSE BSR SpMM (fp32 blocks, CUSPARSE_SPMM_ALG_DEFAULT): the descriptor the reference builds in
 * src/formats/sparse_bsr.cu:139-155 but never runs (engine_bsr.hpp:24 SUPPORT_CUSPARSE = false). */
int cuspmm_cusparse_spmm_bsr(const uint32_t *blockRowPtrs_dev, const uint32_t *blockColIdxs_dev, const float *blocks_dev,
                             uint32_t numBlockRows, uint32_t numBlockCols, uint32_t numBlocks, uint32_t blockSize,
                             const float *B_dev, uint32_t N, float *C_dev, int warmup, int iters,
                             float *avg_ms, float *min_ms);

/* cuSPARSE Blocked-ELL SpMM (fp32, CUSPARSE_SPMM_BLOCKED_ELL_ALG1): the ELL descriptor the reference leaves unimplemented
 * (src/formats/sparse_ell.cu:92-105 throws "not implemented").  Blocked-ELL is cuSPARSE's only ELL
 * flavour: every block row holds the same number of bs x bs blocks, so the BSR operand given here is padded on the device to
 * its longest block row (column index -1, zero values); that re-layout is outside the timed region.  *ell_width_blocks
 * (optional) receives the padded width in blocks. */
int cuspmm_cusparse_spmm_blockedell(const uint32_t *blockRowPtrs_dev, const uint32_t *blockColIdxs_dev, const float *blocks_dev,
                                    uint32_t numBlockRows, uint32_t numBlockCols, uint32_t numBlocks, uint32_t blockSize,
                                    const float *B_dev, uint32_t N, float *C_dev, int warmup, int iters,
                                    float *avg_ms, float *min_ms, uint32_t *ell_width_blocks);

#ifdef __cplusplus
}
#endif
#endif /* CUSPMM_B200_H */
