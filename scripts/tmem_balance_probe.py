#!/usr/bin/env python
"""Diagnostic for the TMEM-staged kernel (CSR variant 5): the same 25605^2, 10 % dense, N = 512 problem with a
perfectly regular pattern (row r has columns r%10, r%10 + 10, ...: every warp has the same work in every K-chunk)
against the Bernoulli pattern of the benchmark.  The gap is the cost of per-chunk imbalance between warps."""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402


def main():
    b = load_package().binding
    M = K = 25605
    N = 512
    per = (K + 9) // 10
    r = torch.arange(M, device="cuda", dtype=torch.int64)
    cols = (r % 10)[:, None] + 10 * torch.arange(per, device="cuda", dtype=torch.int64)[None, :]
    ok = cols < K
    lens = ok.sum(dim=1)
    rp = torch.zeros(M + 1, dtype=torch.int64, device="cuda")
    rp[1:] = torch.cumsum(lens, 0)
    ci = cols[ok].to(torch.int32)
    nnz = int(ci.numel())
    va = torch.rand(nnz, device="cuda") * 2 - 1
    rp = rp.to(torch.int32)
    Bd = torch.rand((K, N), device="cuda") * 2 - 1
    Cd = torch.empty((M, N), device="cuda")
    ref = None
    for v in (3, 5):
        for _ in range(3):
            b.spmm_csr(rp, ci, va, M, K, Bd, variant=v, out=Cd)
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); b.spmm_csr(rp, ci, va, M, K, Bd, variant=v, out=Cd); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        if ref is None:
            ref = Cd.clone()
        same = bool((ref == Cd).all().item())
        print(f"regular pattern nnz={nnz} v{v}: min {min(ts):.4f} ms median {statistics.median(ts):.4f} ms  bit-identical to v3: {same}")


if __name__ == "__main__":
    main()
