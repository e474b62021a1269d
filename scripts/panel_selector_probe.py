import sys, json; sys.path.insert(0, ".")
import torch, importlib
from __graft_entry__ import load_package
b = load_package().binding
wl = importlib.import_module("cuspmm_b200.workloads")
from scripts.quad_probe import timed
K, N = 25605, 512
for M in (25605, 17070, 12803, 8535, 6401, 4268, 3200):
    rp, ci, va = wl.gen_csr_device(M, K, 0.10, seed=618)
    Bd = wl.gen_dense_device(K, N, seed=619)
    Cd = torch.empty((M, N), device="cuda")
    rec = {"M": M, "selected": b.csr_selected_variant(M, K, int(ci.numel()), N)}
    for v in (3, 5):
        rec[f"v{v}"] = round(timed(lambda: b.spmm_csr(rp, ci, va, M, K, Bd, variant=v, out=Cd), iters=7)[0], 4)
    print(json.dumps(rec), flush=True)
