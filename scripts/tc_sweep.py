#!/usr/bin/env python
"""variant 8 (tensor cores) against the fp32 kernel the selector picks without it, over shapes and densities -> JSON lines
(profiles/r02_tc_sweep.jsonl; the thresholds in csr_select_variant come from here).  CUSPMM_TENSOR=0 must be set so that
variant 0 resolves to the fp32 choice."""
import sys, json; sys.path.insert(0, ".")
import torch, importlib
from __graft_entry__ import load_package
b = load_package().binding
wl = importlib.import_module("cuspmm_b200.workloads")
from scripts.quad_probe import timed

cases = []
for MK in (4096, 8192, 16384, 25605):
    for N in (128, 256, 512, 1024, 2048):
        for d in (0.02, 0.05, 0.1, 0.2, 0.5):
            if MK * MK * d * 8 > 3.0e9 or MK * N * 4 * 3 > 2e9: continue
            cases.append((MK, MK, N, d))
for M in (1600, 3200, 6401, 12803):          # row panels of the BASELINE matrix (16 / 8 / 4 / 2 GPUs)
    for d in (0.05, 0.1, 0.3):
        cases.append((M, 25605, 512, d))
for (M, K, N, d) in [(11008, 4096, 4096, 0.1), (11008, 4096, 4096, 0.5), (4096, 11008, 2048, 0.1), (50000, 2048, 512, 0.1), (2048, 50000, 512, 0.1)]:
    cases.append((M, K, N, d))
only = sys.argv[1:] and sys.argv[1]
for i, (M, K, N, d) in enumerate(cases):
    rp, ci, va = wl.gen_csr_device(M, K, d, seed=900 + i)
    Bd = wl.gen_dense_device(K, N, seed=901 + i)
    nnz = int(ci.numel())
    C = torch.empty((M, N), device="cuda")
    rec = {"M": M, "K": K, "N": N, "d": d, "nnz": nnz, "fp32_variant": b.csr_selected_variant(M, K, nnz, N)}
    rec["fp32_ms"] = round(timed(lambda: b.spmm_csr(rp, ci, va, M, K, Bd, variant=0, out=C), iters=5)[0], 4)
    rec["v8_ms"] = round(timed(lambda: b.spmm_csr(rp, ci, va, M, K, Bd, variant=8, out=C), iters=5)[0], 4)
    rec["speedup"] = round(rec["fp32_ms"] / rec["v8_ms"], 3)
    print(json.dumps(rec), flush=True)
    del rp, ci, va, Bd, C
    torch.cuda.empty_cache()
