#!/usr/bin/env python
"""Sweep of the BASELINE.json configs on one GPU: every format / variant beside same-run cuSPARSE.
Writes one JSON line per (config, kernel) to stdout: ms (median of --iters, CUDA events), GFLOP/s,
algorithmic HBM GB/s and its fraction of the measured peak, the three bounds of BASELINE.md section 3,
and the speed-up over cuSPARSE.  Matrices <= L2 in footprint are timed with an L2 flush between
iterations (a 256 MB memset), stated per line."""
import argparse
import importlib
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

CONFIGS = [
    # name, M, K, density, N
    ("cfg2 medium_4096 s90", 4096, 4096, 0.10, 512),
    ("cfg3 medium_4000 s99 N128", 4000, 4000, 0.01, 128),
    ("cfg3 medium_4000 s99 N512", 4000, 4000, 0.01, 512),
    ("cfg3 medium_4000 s99 N2048", 4000, 4000, 0.01, 2048),
    ("cfg3 medium_4000 s95 N512", 4000, 4000, 0.05, 512),
    ("cfg3 medium_4000 s90 N128", 4000, 4000, 0.10, 128),
    ("cfg3 medium_4000 s90 N512", 4000, 4000, 0.10, 512),
    ("cfg3 medium_4000 s90 N2048", 4000, 4000, 0.10, 2048),
    ("cfg3 medium_4000 s70 N512", 4000, 4000, 0.30, 512),
    ("cfg3 medium_4000 s50 N128", 4000, 4000, 0.50, 128),
    ("cfg3 medium_4000 s50 N512", 4000, 4000, 0.50, 512),
    ("cfg3 medium_4000 s50 N2048", 4000, 4000, 0.50, 2048),
    ("cfg4 large_25605 s90 N512", 25605, 25605, 0.10, 512),
    ("cfg4 large_25605 s70 N512", 25605, 25605, 0.30, 512),
    ("cfg4 large_25605 s50 N512", 25605, 25605, 0.50, 512),
    ("cfg5 ffn_11008x4096 s90 N4096", 11008, 4096, 0.10, 4096),
    ("cfg5 ffn_11008x4096 s70 N4096", 11008, 4096, 0.30, 4096),
    ("cfg5 ffn_11008x4096 s50 N4096", 11008, 4096, 0.50, 4096),
    ("cfg5 large_20000 s90 N512", 20000, 20000, 0.10, 512),
    ("real-like large_20000 d7e-4 N512", 20000, 20000, 0.00069, 512),
]
HBM = 6451.8


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=7)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    b = load_package().binding
    wl = importlib.import_module("cuspmm_b200.workloads")
    flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")     # 256 MB, read (not written) to evict:
    # a written flush buffer leaves ~126 MB of dirty lines whose write-back is charged to the next kernel

    def timeit(fn, cold):
        for _ in range(2):
            fn()
        ts = []
        for _ in range(a.iters):
            if cold:
                flush.sum()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts), min(ts)

    for name, M, K, d, N in CONFIGS:
        if a.only and a.only not in name:
            continue
        rp, ci, va = wl.gen_csr_device(M, K, d, seed=618)
        Bd = wl.gen_dense_device(K, N, seed=619)
        nnz = int(ci.numel())
        Cd = torch.empty((M, N), device="cuda")
        rows = torch.repeat_interleave(torch.arange(M, device="cuda", dtype=torch.int32), (rp[1:] - rp[:-1]).to(torch.int64))
        sp, sc, sv = b.csr_to_sell(rp, ci, va, M)
        flops = 2.0 * nnz * N
        byts = {"csr": wl.csr_bytes(M, K, N, nnz), "coo": wl.coo_bytes(M, K, N, nnz),
                "ell": wl.sell_bytes(M, K, N, int(sc.numel()), int(sp.numel()) - 1)}
        cold = byts["csr"] < 200e6
        tmp = torch.empty_like(Cd)
        try:
            if cold:     # same protocol as ours: one timed launch after a read-flush of L2, median of iters
                b.cusparse_spmm(0, rp, ci, va, M, K, Bd, tmp, warmup=2, iters=1)
                c1, c2 = [], []
                for _ in range(a.iters):
                    flush.sum(); c1.append(b.cusparse_spmm(0, rp, ci, va, M, K, Bd, tmp, warmup=0, iters=1)[0])
                    flush.sum(); c2.append(b.cusparse_spmm(1, rows, ci, va, M, K, Bd, tmp, warmup=0, iters=1)[0])
                cus_avg, coo_avg = statistics.median(c1), statistics.median(c2)
            else:
                cus_avg, cus_min = b.cusparse_spmm(0, rp, ci, va, M, K, Bd, tmp, warmup=2, iters=a.iters)
                coo_avg, _ = b.cusparse_spmm(1, rows, ci, va, M, K, Bd, tmp, warmup=2, iters=a.iters)
        except Exception as ex:
            cus_avg = coo_avg = float("nan")
            print(json.dumps({"config": name, "cusparse_error": str(ex)[:100]}), flush=True)
        kernels = [("csr", v, (lambda v=v: b.spmm_csr(rp, ci, va, M, K, Bd, variant=v, out=Cd))) for v in (0, 1, 2, 3, 5, 7, 8)]
        kernels += [("coo", v, (lambda v=v: b.spmm_coo(rows, ci, va, M, K, Bd, variant=v, out=Cd))) for v in (1, 2)]
        kernels += [("ell", v, (lambda v=v: b.spmm_sell(sp, sc, sv, M, K, Bd, variant=v, out=Cd))) for v in (0, 1, 2, 3, 4, 5, 6)]
        for fmt, v, fn in kernels:
            try:
                med, mn = timeit(fn, cold)
            except Exception as ex:
                print(json.dumps({"config": name, "format": fmt, "variant": v, "declined": str(ex)[-80:]}), flush=True)
                continue
            base = cus_avg if fmt != "coo" else coo_avg
            rec = {"config": name, "M": M, "K": K, "N": N, "density": d, "nnz": nnz, "format": fmt, "variant": v,
                   "ms": round(med, 4), "ms_min": round(mn, 4), "gflops": round(flops / med / 1e6, 1),
                   "alg_MB": round(byts[fmt] / 1e6, 1), "hbm_GBs": round(byts[fmt] / med / 1e6, 1),
                   "hbm_frac_of_measured": round(byts[fmt] / med / 1e6 / HBM, 4),
                   "t_hbm_ms": round(byts[fmt] / HBM / 1e6, 4), "t_fp32_ms": round(flops / 74.4e9, 4),
                   "t_smem_ms": round(4.0 * nnz * N / 37.2e9, 4),
                   "cusparse_ms": round(base, 4), "vs_cusparse": round(base / med, 2),
                   "timing": "L2 evicted by reading 256 MB between iterations (ours and cuSPARSE)" if cold else "inputs larger than L2"}
            print(json.dumps(rec), flush=True)
        del rp, ci, va, Bd, Cd, rows, sp, sc, sv, tmp
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
