#!/usr/bin/env python
"""Strong scaling of ONE matrix over 1/2/4/8 GPUs of a node (BASELINE configs[4]: large_20000 row-sharded).

Uses the single-process multi-GPU engine (cuspmm_mgpu_*): A is split into nnz-balanced contiguous row
panels, B is replicated over NVLink (peer copies), every GPU multiplies its panel on its own stream;
with --gather each kernel stores its C rows straight into GPU 0's C through peer memory.  Time =
max over the devices' CUDA-event times (per iteration).  One JSON line per GPU count."""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="large_20000")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--gather", action="store_true")
    ap.add_argument("--gpus", type=int, nargs="+", default=None)
    a = ap.parse_args()
    b = load_package().binding
    wl = importlib.import_module("cuspmm_b200.workloads")
    M, K, d, N = wl.NAMED[a.workload]
    rp, ci, va = wl.gen_csr_device(M, K, d, seed=618)
    Bd = wl.gen_dense_device(K, N, seed=619)
    rp_h = rp.cpu().numpy().view(np.uint32).copy()
    ci_h = ci.cpu().numpy().view(np.uint32).copy()
    va_h = va.cpu().numpy().copy()
    B_h = Bd.cpu().numpy().copy()
    nnz = int(ci_h.shape[0])
    ref = b.spmm_csr(rp, ci, va, M, K, Bd, variant=0).cpu().numpy()
    del rp, ci, va, Bd
    torch.cuda.empty_cache()
    ndev = torch.cuda.device_count()
    base = None
    for n in (a.gpus or [g for g in (1, 2, 4, 8) if g <= ndev]):
        plan = b.MgpuPlan(n, rp_h, ci_h, va_h, M, K, N)
        try:
            plan.set_B(B_h)
            plan.run(variant=0, gather=a.gather, iters=3)
            ms = plan.run(variant=0, gather=a.gather, iters=a.iters)
            same = bool((plan.get_C() == ref).all())
            splits = plan.splits()
            per = np.diff(rp_h[splits].astype(np.int64))
        finally:
            plan.close()
        base = base or ms
        print(json.dumps({"workload": a.workload, "M": M, "K": K, "N": N, "nnz": nnz, "n_gpus": n, "gather": a.gather,
                          "ms": round(ms, 4), "gflops": round(2.0 * nnz * N / ms / 1e6, 1), "speedup_vs_1": round(base / ms, 2),
                          "efficiency": round(base / ms / n, 3), "nnz_imbalance": round(float(per.max() / per.mean()), 4),
                          "bit_identical_to_single_gpu": same}), flush=True)
    torch.cuda.set_device(0)


if __name__ == "__main__":
    main()
