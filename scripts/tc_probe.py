#!/usr/bin/env python
"""CSR variant 8 (tensor cores, three-product split) against the fp32 kernels: norm-wise error and time.
    python scripts/tc_probe.py [quick]      -> JSON lines"""
import sys, json; sys.path.insert(0, ".")
import torch, importlib
from __graft_entry__ import load_package
b = load_package().binding
wl = importlib.import_module("cuspmm_b200.workloads")
from scripts.quad_probe import timed

def check(M, K, N, d, seed, time_it=False, ref_variant=1):
    rp, ci, va = wl.gen_csr_device(M, K, d, seed=seed)
    Bd = wl.gen_dense_device(K, N, seed=seed + 1)
    nnz = int(ci.numel())
    C8 = torch.full((M, N), float("nan"), device="cuda")
    b.spmm_csr(rp, ci, va, M, K, Bd, variant=8, out=C8)
    C1 = b.spmm_csr(rp, ci, va, M, K, Bd, variant=ref_variant)
    Cabs = b.spmm_csr(rp, ci, va.abs(), M, K, Bd.abs(), variant=ref_variant)
    torch.cuda.synchronize()
    err = ((C8 - C1).abs() / Cabs.clamp_min(1e-30)).max().item()
    rec = {"M": M, "K": K, "N": N, "d": d, "nnz": nnz, "nan": bool(torch.isnan(C8).any().item()), "max_normwise_err_vs_fp32": err,
           "max_abs": (C8 - C1).abs().max().item()}
    if time_it:
        for v in (8, 0):
            out = torch.empty((M, N), device="cuda")
            ms = timed(lambda: b.spmm_csr(rp, ci, va, M, K, Bd, variant=v, out=out), iters=7)[0]
            rec[f"v{v}_ms"] = round(ms, 4)
        rec["selected"] = b.csr_selected_variant(M, K, nnz, N)
        rec["v8_tflops_useful"] = round(2.0 * nnz * N / rec["v8_ms"] / 1e9, 2)
    print(json.dumps(rec), flush=True)

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
mid = len(sys.argv) > 1 and sys.argv[1] == "mid"
check(256, 64, 256, 0.2, 1)
check(300, 100, 512, 0.1, 2)
check(1000, 777, 512, 0.1, 3)
check(1000, 777, 130, 0.3, 4)
check(5000, 3000, 512, 0.05, 5, time_it=True)
if mid:
    check(25605, 25605, 512, 0.10, 618, time_it=True, ref_variant=5)
    check(25605, 25605, 512, 0.02, 701, time_it=True, ref_variant=0)
    check(25605, 25605, 512, 0.50, 700, time_it=True, ref_variant=5)
elif not quick:
    check(25605, 25605, 512, 0.10, 618, time_it=True, ref_variant=5)
    check(25605, 25605, 512, 0.50, 700, time_it=True, ref_variant=5)
    check(25605, 25605, 512, 0.02, 701, time_it=True, ref_variant=0)
    check(3200, 25605, 512, 0.10, 702, time_it=True, ref_variant=0)
    check(4096, 4096, 4096, 0.10, 703, time_it=True, ref_variant=0)
