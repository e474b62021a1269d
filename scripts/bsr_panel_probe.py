#!/usr/bin/env python
"""tcgen05 BSR: block-row kernel vs panel kernel (union walk) on BASELINE configs[3] (25605^2 padded, 10 % of the blocks), per
panel height.  CUSPMM_BSR_PANEL / CUSPMM_BSR_P are read once per process: run once per setting."""
import json
import os
import statistics
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

M = K = 25605
N = int(os.environ.get("BSR_N", "512"))
dens = float(os.environ.get("BSR_DENSITY", "0.10"))
for bs in (16, 32):
    g = torch.Generator(device="cuda"); g.manual_seed(618)
    nbr, nbc = (M + bs - 1) // bs, (K + bs - 1) // bs
    mask = torch.rand((nbr, nbc), generator=g, device="cuda") < dens
    brp = torch.zeros(nbr + 1, dtype=torch.int64, device="cuda"); brp[1:] = torch.cumsum(mask.sum(dim=1, dtype=torch.int64), 0)
    bci = mask.nonzero(as_tuple=False)[:, 1].to(torch.int32)
    nb = int(bci.numel())
    blocks = torch.rand(nb * bs * bs, generator=g, device="cuda") * 2 - 1
    brp = brp.to(torch.int32)
    Bd = wl.gen_dense_device(nbc * bs, N, seed=619)
    Cd = torch.empty((nbr * bs, N), dtype=torch.float32, device="cuda")
    plan = b.BsrTcPlan(brp, bci, blocks, nbr, bs, nbc * bs, N, dtype="bf16")
    plan.prepare_B(Bd)
    med, mn = timed(lambda: plan.run(out=Cd), iters=9, warmup=3)
    flops = 2.0 * nb * bs * bs * N
    print(json.dumps({"bs": bs, "N": N, "block_density": dens, "panel": os.environ.get("CUSPMM_BSR_PANEL", "auto"),
                      "P": os.environ.get("CUSPMM_BSR_P", "auto"), "ms": round(med, 4), "ms_min": round(mn, 4),
                      "tflops_executed": round(flops / (med * 1e-3) / 1e12, 1), "checksum": float(Cd.double().sum().item())}), flush=True)
    plan.close()
