#!/usr/bin/env python
"""Staged CSR kernel (variant 3) in thread-block clusters with multicast TMA (CUSPMM_STAGED_CLUSTER = 2 / 4 / 8, read once per
process): bit-identity against variant 1 and time per row-panel height of large_25605 (the strong-scaling shapes)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402
import importlib  # noqa: E402

pkg = load_package()
b = pkg.binding
wl = importlib.import_module("cuspmm_b200.workloads")
from scripts.quad_probe import timed  # noqa: E402

cs = os.environ.get("CUSPMM_STAGED_CLUSTER", "0")
K, N = 25605, 512
for M in (3200, 3201, 6401, 12803, 25605, 100, 31):
    rp, ci, va = wl.gen_csr_device(M, K, 0.10, seed=618)
    Bd = wl.gen_dense_device(K, N, seed=619)
    Cd = torch.empty((M, N), dtype=torch.float32, device="cuda")
    ref = b.spmm_csr(rp, ci, va, M, K, Bd, variant=1)
    got = b.spmm_csr(rp, ci, va, M, K, Bd, variant=3)
    torch.cuda.synchronize()
    same = bool((ref == got).all().item())
    med, mn = timed(lambda: b.spmm_csr(rp, ci, va, M, K, Bd, variant=3, out=Cd), iters=7)
    print(json.dumps({"cluster": cs, "M": M, "bit_identical": same, "ms": round(med, 4), "ms_min": round(mn, 4),
                      "tflops": round(2.0 * int(ci.numel()) * N / (med * 1e-3) / 1e12, 2)}), flush=True)
