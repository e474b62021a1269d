#!/usr/bin/env python
"""Kernel-only timing helper for tuning (CUDA events, device-resident inputs).
    python scripts/ktime.py --workload large_25605 --format csr --variant 3 --iters 10
Prints one line per run: workload format variant ms_min ms_median GFLOP/s(median)."""
import argparse
import importlib
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="large_25605")
    ap.add_argument("--format", default="csr")
    ap.add_argument("--variant", type=int, nargs="+", default=[0])
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--N", type=int, default=0)
    ap.add_argument("--density", type=float, default=0.0)
    ap.add_argument("--cusparse", action="store_true")
    ap.add_argument("--M", type=int, default=0)
    a = ap.parse_args()
    b = load_package().binding
    wl = importlib.import_module("cuspmm_b200.workloads")
    M, K, d, N = wl.NAMED[a.workload]
    if a.M:
        M = a.M
    if a.N:
        N = a.N
    if a.density:
        d = a.density
    rp, ci, va = wl.gen_csr_device(M, K, d, seed=618)
    Bd = wl.gen_dense_device(K, N, seed=619)
    nnz = int(ci.numel())
    Cd = torch.empty((M, N), device="cuda")
    for fmt in a.format.split(","):
        if fmt == "coo":
            rows = torch.repeat_interleave(torch.arange(M, device="cuda", dtype=torch.int32), (rp[1:] - rp[:-1]).to(torch.int64))
        if fmt == "ell":
            sp, sc, sv = b.csr_to_sell(rp, ci, va, M)
        for v in a.variant:
            if fmt == "csr":
                step = lambda: b.spmm_csr(rp, ci, va, M, K, Bd, variant=v, out=Cd)
            elif fmt == "coo":
                step = lambda: b.spmm_coo(rows, ci, va, M, K, Bd, variant=v, out=Cd)
            else:
                step = lambda: b.spmm_sell(sp, sc, sv, M, K, Bd, variant=v, out=Cd)
            try:
                for _ in range(3):
                    step()
            except Exception as ex:
                print(f"{a.workload} N={N} d={d} {fmt} v{v}: {str(ex)[:120]}")
                continue
            ts = []
            for _ in range(a.iters):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); step(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            med = statistics.median(ts)
            print(f"{a.workload} M={M} K={K} N={N} d={d} nnz={nnz} {fmt} v{v}: min {min(ts):.4f} ms  median {med:.4f} ms  "
                  f"{2.0 * nnz * N / med / 1e6:.1f} GFLOP/s", flush=True)
    if a.cusparse:
        tmp = torch.empty_like(Cd)
        avg, mn = b.cusparse_spmm(0, rp, ci, va, M, K, Bd, tmp, warmup=2, iters=a.iters)
        print(f"{a.workload} N={N} d={d} cusparse CSR_ALG2: min {mn:.4f} ms avg {avg:.4f} ms  {2.0 * nnz * N / avg / 1e6:.1f} GFLOP/s")


if __name__ == "__main__":
    main()
