#!/usr/bin/env python
"""BASELINE configs[3]: large_25605 BSR, block 16 / 32, 10 % of blocks stored, bf16/fp16 blocks on
tcgen05 tensor cores (fp32 accumulate) beside the fp32 SIMT BSR kernel.  One JSON line per kernel."""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

HBM, TENSOR = 6451.8, 1682.6   # MEASURED_PEAKS.json: GB/s, bf16 TFLOP/s (burst)


def gen_block_sparse(M, K, bs, frac, seed=618):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    nbr, nbc = (M + bs - 1) // bs, (K + bs - 1) // bs
    mask = torch.rand((nbr, nbc), generator=g, device="cuda") < frac
    counts = mask.sum(dim=1, dtype=torch.int64)
    rp = torch.zeros(nbr + 1, dtype=torch.int64, device="cuda"); rp[1:] = torch.cumsum(counts, 0)
    ci = mask.nonzero(as_tuple=False)[:, 1].to(torch.int32)
    nb = int(ci.numel())
    blocks = torch.rand(nb * bs * bs, generator=g, device="cuda") * 2 - 1
    return rp.to(torch.int32), ci, blocks, nbr, nbc, nb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", type=int, default=25605)
    ap.add_argument("--N", type=int, default=512)
    ap.add_argument("--frac", type=float, default=0.10)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--bs", type=int, nargs="+", default=[16, 32])
    a = ap.parse_args()
    b = load_package().binding
    M = K = a.M
    N = a.N

    def timeit(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(a.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts), min(ts)

    for bs in a.bs:
        rp, ci, blocks, nbr, nbc, nb = gen_block_sparse(M, K, bs, a.frac)
        Kp, Mp = nbc * bs, nbr * bs
        g = torch.Generator(device="cuda"); g.manual_seed(619)
        Bd = torch.rand((Kp, N), generator=g, device="cuda") * 2 - 1
        Cd = torch.empty((Mp, N), device="cuda")
        flops = 2.0 * nb * bs * bs * N
        base = {"config": f"cfg4 large_25605 BSR {bs}x{bs}, {a.frac:.0%} of blocks, N={N}", "numBlocks": nb, "M_padded": Mp}
        med, mn = timeit(lambda: b.spmm_bsr_f32(rp, ci, blocks, nbr, bs, bs, Kp, Bd, out=Cd))
        byts = nb * bs * bs * 4 + 4 * nb + 4 * (nbr + 1) + 4 * Kp * N + 4 * Mp * N
        ref = Cd.clone()
        print(json.dumps({**base, "kernel": "bsr_f32_simt", "ms": round(med, 4), "tflops_executed": round(flops / med / 1e9, 2),
                          "alg_MB": round(byts / 1e6, 1), "hbm_frac": round(byts / med / 1e6 / HBM, 4)}), flush=True)
        try:
            tmp = torch.empty_like(Cd)
            cavg, cmin = b.cusparse_spmm_bsr(rp, ci, blocks, nbr, nbc, bs, Bd, tmp, warmup=2, iters=a.iters)
            cerr = ((tmp - ref).abs().max() / ref.abs().max()).item()
            print(json.dumps({**base, "kernel": "cusparse_bsr_f32 (CUSPARSE_SPMM_ALG_DEFAULT)", "ms": round(cavg, 4), "ms_min": round(cmin, 4),
                              "tflops_executed": round(flops / cavg / 1e9, 2), "max_abs_diff_vs_f32_kernel_over_max": cerr}), flush=True)
            del tmp
        except Exception as ex:
            cavg = None
            print(json.dumps({**base, "kernel": "cusparse_bsr_f32", "error": str(ex)[:160]}), flush=True)
        for dt in ("bf16", "fp16"):
            plan = b.BsrTcPlan(rp, ci, blocks, nbr, bs, Kp, N, dtype=dt)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); plan.prepare_B(Bd); e1.record(); torch.cuda.synchronize()
            prep = e0.elapsed_time(e1)
            med, mn = timeit(lambda: plan.run(out=Cd))
            byts = nb * bs * bs * 2 + 4 * nb + 4 * (nbr + 1) + 2 * Kp * N + 4 * Mp * N
            err = ((Cd - ref).abs().max() / ref.abs().max()).item()
            print(json.dumps({**base, "kernel": f"bsr_tcgen05_{dt}", "ms": round(med, 4), "ms_min": round(mn, 4),
                              "tflops_executed": round(flops / med / 1e9, 2), "tensor_frac_of_measured_bf16": round(flops / med / 1e9 / TENSOR, 4),
                              "alg_MB": round(byts / 1e6, 1), "hbm_GBs": round(byts / med / 1e6, 1), "hbm_frac": round(byts / med / 1e6 / HBM, 4),
                              "l2_gather_GB": round(nb * bs * N * 2 / 1e9, 2), "prepare_B_ms": round(prep, 4),
                              "max_abs_diff_vs_f32_kernel_over_max": err,
                              "vs_cusparse_bsr_f32": round(cavg / med, 2) if cavg else None}), flush=True)
            plan.close()
        del rp, ci, blocks, Bd, Cd, ref
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
