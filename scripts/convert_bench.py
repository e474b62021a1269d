#!/usr/bin/env python
"""The device converters (north_star (b)) measured as the HBM-bound kernels they are: milliseconds and GB/s of compulsory
bytes against the measured HBM peak, on the headline shape (25605^2, 65.6 M non-zeros).

  CSR -> sliced ELL                      unstructured d = 0.10
  CSR -> BSR 16x16 / 32x32               block-sparse pattern (10 % of the blocks, every stored block dense: the same 65.6 M
                                         non-zeros; unstructured 10 % would fill every block = 2.6 GB of fp32 blocks)
                                         sort-free bitmap path, and the radix-sort path (CUSPMM_BSR_CONVERT_SORT=1, other process)
  column-ELL -> CSR                      the reference's ELL storage (K x maxColNnz slots), stable radix sort by row
  COO -> CSR row pointers, partition     searches (not bandwidth figures)
One JSON line per converter."""
import importlib
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

b = load_package().binding
wl = importlib.import_module("cuspmm_b200.workloads")
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0


_burn = torch.randn((8192, 8192), device="cuda", dtype=torch.bfloat16)


def timed(fn, iters=5):
    """Average over `iters` back-to-back calls between two CUDA events, after a burn that brings the clocks up (a converter
    timed cold, one call after an idle period, shows 4-5x the time: the GPU sits in a low power state)."""
    fn()
    for _ in range(20):
        torch.mm(_burn, _burn)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def emit(name, ms, nbytes, **kw):
    print(json.dumps({"converter": name, "ms": round(ms, 4), "algorithmic_bytes": int(nbytes), "GBs": round(nbytes / ms / 1e6, 1),
                      "frac_of_hbm_peak": round(nbytes / ms / 1e6 / PEAK, 4), "hbm_peak_GBs": PEAK, **kw}), flush=True)


M = K = 25605
rp, ci, va = wl.gen_csr_device(M, K, 0.10, seed=618)
nnz = int(ci.numel())
sp, sc, sv = b.csr_to_sell(rp, ci, va, M)
slots = int(sc.numel())
emit("csr_to_sell32", timed(lambda: b.csr_to_sell(rp, ci, va, M)), 2 * 4 * (M + 1) + 8 * nnz + 8 * slots, nnz=nnz, slots=slots)
rows = torch.repeat_interleave(torch.arange(M, device="cuda", dtype=torch.int32), (rp[1:] - rp[:-1]).to(torch.int64))
emit("coo_to_csr_rowptrs", timed(lambda: b.coo_to_csr_rowptrs(rows, M)), 4 * (M + 1), note="M+1 binary searches over rowIdxs")
emit("partition_rows_by_nnz_x8", timed(lambda: b.partition_rows_by_nnz(rp, M, nnz, 8)), 4 * (M + 1), note="32-ary searches, includes the stream sync")
emit("csr_check_sorted", timed(lambda: b.csr_check_sorted(rp, ci, M, K)), 4 * (M + 1) + 4 * nnz)
del sp, sc, sv, rows
# column-ELL of a smaller unstructured matrix (the slot array is K x maxColNnz: at 25605^2 that is 69 M slots)
Ms = Ks = 12000
rps, cis, vas = wl.gen_csr_device(Ms, Ks, 0.10, seed=5)
dense_cols = torch.zeros(Ks, dtype=torch.int64, device="cuda")
dense_cols.index_add_(0, cis.to(torch.int64), torch.ones_like(cis, dtype=torch.int64))
W = int(dense_cols.max().item())
order = torch.argsort(cis.to(torch.int64) * Ms + torch.repeat_interleave(torch.arange(Ms, device="cuda"), (rps[1:] - rps[:-1]).to(torch.int64)))
colsorted = cis[order].to(torch.int64)
rowsorted = torch.repeat_interleave(torch.arange(Ms, device="cuda"), (rps[1:] - rps[:-1]).to(torch.int64))[order]
start = torch.zeros(Ks + 1, dtype=torch.int64, device="cuda"); start[1:] = torch.cumsum(dense_cols, 0)
pos = torch.arange(colsorted.numel(), device="cuda") - start[colsorted]
ellR = torch.full((Ks * W,), -1, dtype=torch.int32, device="cuda")
ellV = torch.zeros(Ks * W, dtype=torch.float32, device="cuda")
ellR[colsorted * W + pos] = rowsorted.to(torch.int32)
ellV[colsorted * W + pos] = vas[order]
nnzs = int(cis.numel())
rp2, ci2, va2 = b.colell_to_csr(ellR, ellV, Ms, Ks, W, nnzs)
ok = bool((rp2 == rps).all().item() and (ci2 == cis).all().item() and (va2 == vas).all().item())
emit("colell_to_csr", timed(lambda: b.colell_to_csr(ellR, ellV, Ms, Ks, W, nnzs)), 8 * Ks * W + 4 * (Ms + 1) + 8 * nnzs,
     shape=f"{Ms}x{Ks}", slots=Ks * W, nnz=nnzs, bit_exact_round_trip=ok, note="stable cub radix sort of all slots by row")
del rps, cis, vas, ellR, ellV, order, colsorted, rowsorted
torch.cuda.empty_cache()
# CSR -> BSR on a block-sparse pattern
for bs in (16, 32):
    g = torch.Generator(device="cuda"); g.manual_seed(618)
    nbr, nbc = (M + bs - 1) // bs, (K + bs - 1) // bs
    mask = torch.rand((nbr, nbc), generator=g, device="cuda") < 0.10
    # CSR of the expanded pattern: row r has the columns of its block row's blocks
    bcols = [None] * 0
    cnt = mask.sum(dim=1, dtype=torch.int64)                               # blocks per block row
    bci = mask.nonzero(as_tuple=False)[:, 1]
    brp = torch.zeros(nbr + 1, dtype=torch.int64, device="cuda"); brp[1:] = torch.cumsum(cnt, 0)
    Mp = nbr * bs
    row_len = torch.repeat_interleave(cnt * bs, bs)                        # entries per CSR row
    rpb = torch.zeros(Mp + 1, dtype=torch.int64, device="cuda"); rpb[1:] = torch.cumsum(row_len, 0)
    nz = int(rpb[-1].item())
    rows_of = torch.repeat_interleave(torch.arange(Mp, device="cuda"), row_len)
    off = torch.arange(nz, device="cuda") - rpb[rows_of]
    cib = (bci[brp[rows_of // bs] + off // bs] * bs + off % bs).to(torch.int32)
    vab = torch.rand(nz, generator=g, device="cuda") * 2 - 1
    rpb32 = rpb.to(torch.int32)
    del rows_of, off, row_len
    r1, c1, b1 = b.csr_to_bsr(rpb32, cib, vab, Mp, nbc * bs, bs, bs)
    ok = bool((r1.to(torch.int64) == brp).all().item() and (c1.to(torch.int64) == bci).all().item())
    nb = int(c1.numel())
    nbytes = 2 * (4 * (Mp + 1) + 4 * nz) + 4 * nz + 4 * nb + 4 * (nbr + 1) + 2 * 4 * nb * bs * bs      # colIdxs twice, vals, memset + scatter of blocks
    emit(f"csr_to_bsr{bs}" + ("_sort" if os.environ.get("CUSPMM_BSR_CONVERT_SORT") else "_bitmap"),
         timed(lambda: b.csr_to_bsr(rpb32, cib, vab, Mp, nbc * bs, bs, bs), iters=3), nbytes, nnz=nz, blocks=nb, pattern_round_trip=ok)
    del rpb, rpb32, cib, vab, r1, c1, b1
    torch.cuda.empty_cache()
