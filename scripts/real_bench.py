#!/usr/bin/env python
"""The reference's own test matrices (tests/golden/real/*.npz = data/medium_*, data/large_* as its converter wrote them)
x a seeded dense B (N = 512 by default): kernel time of the selector's choice and of every CSR variant, algorithmic
HBM GB/s and its fraction of the measured peak, same-run cuSPARSE CSR_ALG2.  L2 is evicted between iterations (all of
these problems are smaller than L2).  One JSON line per matrix."""
import argparse
import glob
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

HBM = 6451.8


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=512)
    ap.add_argument("--iters", type=int, default=9)
    ap.add_argument("--variants", type=int, nargs="+", default=[0, 1, 2, 4])
    a = ap.parse_args()
    b = load_package().binding
    flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")

    def timeit(fn):
        for _ in range(2):
            fn()
        ts = []
        for _ in range(a.iters):
            flush.sum()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "real", "*.npz"))):
        z = np.load(path)
        M, K, N = int(z["M"]), int(z["K"]), a.N
        rp, ci, va = b.dev_u32(z["rowPtrs"].astype(np.uint32)), b.dev_u32(z["colIdxs"].astype(np.uint32)), b.dev_f32(z["vals"].astype(np.float32))
        nnz = int(ci.numel())
        g = torch.Generator(device="cuda"); g.manual_seed(619)
        Bd = torch.rand((K, N), generator=g, device="cuda") * 2 - 1
        Cd = torch.empty((M, N), device="cuda")
        ms = {}
        for v in a.variants:
            ms[v] = timeit(lambda v=v: b.spmm_csr(rp, ci, va, M, K, Bd, variant=v, out=Cd, allow_split=True))
        sel = ms[a.variants[0]]
        tmp = torch.empty_like(Cd)
        b.cusparse_spmm(0, rp, ci, va, M, K, Bd, tmp, warmup=2, iters=1)
        cs = []
        for _ in range(a.iters):
            flush.sum(); cs.append(b.cusparse_spmm(0, rp, ci, va, M, K, Bd, tmp, warmup=0, iters=1)[0])
        cus = statistics.median(cs)
        byts = 8.0 * nnz + 4.0 * (M + 1) + 4.0 * K * N + 4.0 * M * N
        lens = np.diff(z["rowPtrs"].astype(np.int64))
        print(json.dumps({"dir": os.path.basename(path)[:-4], "matrix": str(z["name"]), "M": M, "K": K, "N": N, "nnz": nnz,
                          "row_nnz_max": int(lens.max()), "row_nnz_mean": round(float(lens.mean()), 1),
                          "ms_selector": round(sel, 4), "ms_by_variant": {str(k): round(v, 4) for k, v in ms.items()},
                          "gflops": round(2.0 * nnz * N / sel / 1e6, 1), "alg_MB": round(byts / 1e6, 1),
                          "hbm_GBs": round(byts / sel / 1e6, 1), "hbm_frac_of_measured": round(byts / sel / 1e6 / HBM, 3),
                          "cusparse_ms": round(cus, 4), "vs_cusparse": round(cus / sel, 2),
                          "max_abs_diff_vs_cusparse": float((tmp - Cd).abs().max().item()),
                          "timing": "L2 evicted by reading 256 MB between iterations (ours and cuSPARSE)"}), flush=True)


if __name__ == "__main__":
    main()
