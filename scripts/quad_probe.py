#!/usr/bin/env python
"""Correctness + timing probe of the all-TMEM CSR kernel (variant 7) against variants 1 / 3 / 5.

    python scripts/quad_probe.py check            # bit-identity vs variant 1 on many shapes (+ sliced ELL)
    python scripts/quad_probe.py time [workload]  # ms per variant on the named workloads

CUSPMM_QUAD_SHAPE selects the kernel shape (read once per process), so shapes are compared by running this script
once per value.  Prints one JSON line per case."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
b = pkg.binding
import importlib  # noqa: E402

wl = importlib.import_module("cuspmm_b200.workloads")


def timed(fn, iters=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def check():
    shape = os.environ.get("CUSPMM_QUAD_SHAPE", "714")
    bad = 0
    cases = []
    for M in (1, 7, 8, 9, 55, 56, 57, 300, 1000, 4096):
        for K in (1, 31, 32, 33, 127, 128, 129, 257, 1000):
            cases.append((M, K, 512, 0.1))
    cases += [(300, 500, 512, 0.0), (300, 500, 512, 0.01), (300, 500, 512, 0.5), (300, 500, 512, 1.0), (64, 4096, 512, 0.9),
              (4096, 4096, 1024, 0.1), (777, 900, 1536, 0.3), (25605, 2000, 512, 0.1), (20000, 1500, 512, 0.02)]
    for (M, K, N, d) in cases:
        rp, ci, va = wl.gen_csr_device(M, K, d, seed=7 + M + K)
        Bd = wl.gen_dense_device(K, N, seed=11)
        ref = b.spmm_csr(rp, ci, va, M, K, Bd, variant=1)
        got = b.spmm_csr(rp, ci, va, M, K, Bd, variant=7)
        torch.cuda.synchronize()
        same = bool((ref == got).all().item())
        rec = {"case": "csr", "shape": shape, "M": M, "K": K, "N": N, "d": d, "nnz": int(ci.numel()), "bit_identical": same}
        if not same:
            bad += 1
            diff = (ref - got).abs()
            rec["max_abs_diff"] = float(diff.max().item())
            rec["bad_elems"] = int((ref != got).sum().item())
            rows = (ref != got).any(dim=1).nonzero()[:8, 0].tolist()
            cols = (ref != got).any(dim=0).nonzero()[:8, 0].tolist()
            rec["first_bad_rows"], rec["first_bad_cols"] = rows, cols
        # strided B / C (ldb > N): column tile narrower than B -> tensor-map TMA
        if N == 512 and M in (300, 4096) and K in (257, 1000):
            Bw = torch.zeros((K, 1024), dtype=torch.float32, device="cuda")
            Bw[:, 256:768] = Bd
            Cw = torch.zeros((M, 1024), dtype=torch.float32, device="cuda")
            b.spmm_csr(rp, ci, va, M, K, Bw[:, 256:768], variant=7, out=Cw[:, 128:640])
            torch.cuda.synchronize()
            rec["strided_bit_identical"] = bool((Cw[:, 128:640] == ref).all().item())
            bad += 0 if rec["strided_bit_identical"] else 1
        print(json.dumps(rec), flush=True)
        # sliced ELL, same kernel on the other row layout
        if M >= 55 and K in (33, 257, 500, 1000, 2000):
            sp, sc, sv = b.csr_to_sell(rp, ci, va, M)
            e1 = b.spmm_sell(sp, sc, sv, M, K, Bd, variant=1)
            e5 = b.spmm_sell(sp, sc, sv, M, K, Bd, variant=5)
            torch.cuda.synchronize()
            same = bool((e1 == e5).all().item()) and bool((e1 == ref).all().item())
            bad += 0 if same else 1
            print(json.dumps({"case": "sell", "shape": shape, "M": M, "K": K, "N": N, "d": d, "bit_identical": same}), flush=True)
    print(json.dumps({"summary": "check", "shape": shape, "failed": bad}), flush=True)
    return bad


def time_variants(names):
    shape = os.environ.get("CUSPMM_QUAD_SHAPE", "714")
    for name in names:
        if name.startswith("panel"):                       # a 1/8 row panel of large_25605 (strong-scaling shape)
            M, K, d, N = 25605 // int(name[5:]), 25605, 0.10, 512
        else:
            M, K, d, N = wl.NAMED[name]
        rp, ci, va = wl.gen_csr_device(M, K, d, seed=618)
        Bd = wl.gen_dense_device(K, N, seed=619)
        Cd = torch.empty((M, N), dtype=torch.float32, device="cuda")
        nnz = int(ci.numel())
        rec = {"case": "time", "workload": name, "shape": shape, "M": M, "K": K, "N": N, "d": d, "nnz": nnz}
        variants = (7,) if os.environ.get("QUAD_ONLY") else (7, 5, 3)
        outs = {}
        for v in variants:
            try:
                med, mn = timed(lambda: b.spmm_csr(rp, ci, va, M, K, Bd, variant=v, out=Cd))
                rec[f"v{v}_ms"] = round(med, 4)
                rec[f"v{v}_min_ms"] = round(mn, 4)
                rec[f"v{v}_tflops"] = round(2.0 * nnz * N / (med * 1e-3) / 1e12, 2)
                outs[v] = Cd.clone()
            except Exception as ex:
                rec[f"v{v}_err"] = str(ex)[:120]
        if 7 in outs and 5 in outs:
            rec["v7_eq_v5"] = bool((outs[7] == outs[5]).all().item())
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "check"
    if mode == "check":
        sys.exit(1 if check() else 0)
    names = sys.argv[2:] or ["large_25605", "large_25605_s50", "large_20000", "medium_4096", "panel8", "ffn_11008x4096_s90"]
    time_variants(names)
