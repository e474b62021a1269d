#!/usr/bin/env python
"""Randomised stress of the staged kernels (CSR 3 / 5 / 6, ELL 2 / 4) against the warp-per-row kernel on device-generated
matrices: bit-identity (3, 5, ELL) or tolerance (6) over many shapes, densities, K tails and strides.  Prints one line per
case and a summary; exits non-zero on the first mismatch."""
import importlib
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402


def main():
    b = load_package().binding
    wl = importlib.import_module("cuspmm_b200.workloads")
    rnd = random.Random(20261018)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    bad = 0
    for it in range(n):
        M = rnd.choice([1, 57, 58, 59, 116, 500, 3000, 8584, 8585, 20011])
        K = rnd.choice([1, 31, 32, 33, 96, 1000, 4097, 9999])
        N = rnd.choice([128, 512, 512, 1024, 1536])
        d = rnd.choice([0.003, 0.03, 0.1, 0.2, 0.45, 0.9])
        rp, ci, va = wl.gen_csr_device(M, K, d, seed=1000 + it)
        pad = rnd.choice([0, 0, 64])
        Bbig = torch.rand((K, N + pad), device="cuda") * 2 - 1
        Bd = Bbig[:, :N]
        ref = b.spmm_csr(rp, ci, va, M, K, Bd, variant=1)
        ok = True
        for v in ((3, 5, 0) if N % 512 == 0 else (3, 0)):
            ok &= bool((b.spmm_csr(rp, ci, va, M, K, Bd, variant=v) == ref).all().item())
        v6 = b.spmm_csr(rp, ci, va, M, K, Bd, variant=6)
        scale = torch.zeros_like(ref)
        ok6 = bool(((v6 - ref).abs() <= 1e-5 * (ref.abs() + 1.0) * max(1.0, d * K) ** 0.5 + 1e-6).all().item())
        sp, sc, sv = b.csr_to_sell(rp, ci, va, M)
        for v in ((2, 4, 0) if N % 512 == 0 else (2, 0)):
            ok &= bool((b.spmm_sell(sp, sc, sv, M, K, Bd, variant=v) == ref).all().item())
        torch.cuda.synchronize()
        print(f"{it:3d} M={M:6d} K={K:5d} N={N:4d} d={d:<5} ldb={N + pad:5d} nnz={int(ci.numel()):9d} bit-identical={ok} v6-close={ok6}", flush=True)
        if not (ok and ok6):
            bad += 1
    print("mismatches:", bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
