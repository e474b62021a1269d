import sys; sys.path.insert(0, ".")
import torch, importlib
from __graft_entry__ import load_package
b = load_package().binding
wl = importlib.import_module("cuspmm_b200.workloads")
d = float(sys.argv[1]) if len(sys.argv) > 1 else 0.10
M = K = 25605; N = 512
rp, ci, va = wl.gen_csr_device(M, K, d, seed=618)
Bd = wl.gen_dense_device(K, N, seed=619)
sp, sc, sv = b.csr_to_sell(rp, ci, va, M)
C = torch.empty((M, N), device="cuda")
for _ in range(2):
    b.spmm_sell(sp, sc, sv, M, K, Bd, variant=6, out=C)
torch.cuda.synchronize()
print("ok")
