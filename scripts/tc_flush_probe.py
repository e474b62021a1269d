#!/usr/bin/env python
"""variant 8 time against the flush period (CUSPMM_TC_FLUSH, read once per process): python scripts/tc_flush_probe.py <density>"""
import sys, json, os; sys.path.insert(0, ".")
import torch, importlib
from __graft_entry__ import load_package
b = load_package().binding
wl = importlib.import_module("cuspmm_b200.workloads")
from scripts.quad_probe import timed
d = float(sys.argv[1]); M = K = 25605; N = 512
rp, ci, va = wl.gen_csr_device(M, K, d, seed=618)
Bd = wl.gen_dense_device(K, N, seed=619)
C = torch.empty((M, N), device="cuda")
ms = timed(lambda: b.spmm_csr(rp, ci, va, M, K, Bd, variant=8, out=C), iters=7)[0]
print(json.dumps({"d": d, "flush": os.environ.get("CUSPMM_TC_FLUSH", "default"), "v8_ms": round(ms, 4)}), flush=True)
