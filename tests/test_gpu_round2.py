"""Round-2 GPU parity tests, all through the C ABI:

  * the all-TMEM kernel (CSR variant 7 / ELL variant 5) against the oracle and bit for bit against variant 1;
  * host-buffer entry points of COO / sliced ELL / BSR (fp32 and tcgen05) and the device-B CSR entry;
  * multi-GPU plans of COO / ELL / BSR (n = 1 on a one-GPU box, every n the box has otherwise);
  * BASELINE configs at their stated sizes: cfg2 (4096^2), the cfg3 grid {50, 90, 99 %} x N {128, 512, 2048} for CSR and
    ELL (sampled rows vs the oracle + checksum), cfg4 BSR tcgen05 at 25616^2 / 25632^2, bf16 and fp16 (sampled block rows vs
    the oracle on rounded operands + fp64 checksum);
  * cuspmm_csr_check_sorted and the selector query.

Tolerances as in test_gpu_parity.py: fp32 kernels max |C - Cref| / (|A||B|) <= 1e-5; tensor-core BSR <= 2e-5 against the oracle on
the ROUNDED operands."""
import importlib

import numpy as np
import pytest

from conftest import random_csr
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

TOL = 1e-5
TC_TOL = 2e-5


@pytest.fixture(scope="module")
def b():
    import torch
    assert torch.cuda.is_available()
    from __graft_entry__ import load_package
    pkg = load_package()
    pkg.lib()
    # this module asserts that the fp32 kernels agree BIT FOR BIT (same terms, same order, same FMA), variant 0 included:
    # keep the selector on the fp32 family.  The tensor-core kernel (variant 8) and the selector with it are tested in
    # test_gpu_tensor.py, to the north_star tolerance.
    pkg.binding.set_csr_tensor_mode(0)
    return pkg.binding


@pytest.fixture(scope="module")
def wl(b):
    return importlib.import_module("cuspmm_b200.workloads")


def dev_csr(b, a):
    return b.dev_u32(a.rowPtrs), b.dev_u32(a.colIdxs), b.dev_f32(a.vals)


def rel_err(got, ref, denom):
    return orc.max_rel_err(got.cpu().numpy() if hasattr(got, "cpu") else got, ref, denom)


# ------------------------------------------------------------------ all-TMEM kernel
@pytest.mark.parametrize("M,K,N,d,skew", [(1, 1, 512, 1.0, False), (57, 129, 512, 0.10, False), (300, 500, 512, 0.5, False),
                                          (1000, 4096, 512, 0.01, False), (777, 900, 1536, 0.30, False),
                                          (2000, 1500, 512, 0.03, True), (300, 257, 1024, 0.0, False)])
def test_quad_kernel_vs_oracle(b, M, K, N, d, skew):
    a = random_csr(M, K, d, seed=100 + M, skew=skew)
    B = np.random.default_rng(3).uniform(-1, 1, (K, N)).astype(np.float32)
    ref, den = orc.spmm_csr(a, B, omp=True), orc.absprod_csr(a, B)
    rp, ci, va = dev_csr(b, a)
    Bd = b.dev_f32(B)
    got = b.spmm_csr(rp, ci, va, M, K, Bd, variant=7)
    assert rel_err(got, ref, den) <= TOL
    assert (got == b.spmm_csr(rp, ci, va, M, K, Bd, variant=1)).all().item()          # same summation order
    sp, sc, sv = b.csr_to_sell(rp, ci, va, M)
    assert (b.spmm_sell(sp, sc, sv, M, K, Bd, variant=5) == got).all().item()


def test_quad_kernel_rejects_other_widths(b):
    a = random_csr(64, 64, 0.2, seed=5)
    rp, ci, va = dev_csr(b, a)
    Bd = b.dev_f32(np.ones((64, 256), np.float32))
    with pytest.raises(b.CuspmmError):
        b.spmm_csr(rp, ci, va, 64, 64, Bd, variant=7)


# ------------------------------------------------------------------ precondition check + selector query
def test_check_sorted_and_selector(b):
    a = random_csr(500, 400, 0.1, seed=9)
    rp, ci, va = dev_csr(b, a)
    assert b.csr_check_sorted(rp, ci, a.M, a.K) == 0
    bad = a.colIdxs.copy()
    p0 = int(a.rowPtrs[17])
    bad[p0], bad[p0 + 1] = bad[p0 + 1], bad[p0]          # one row out of order
    assert b.csr_check_sorted(rp, b.dev_u32(bad), a.M, a.K) == 1
    assert b.csr_selected_variant(25605, 25605, 65571195, 512) in (3, 5, 7)
    assert b.csr_selected_variant(120, 210, 840, 21) == 4


# ------------------------------------------------------------------ host-buffer entry points
def _host_case():
    a = random_csr(9000, 2500, 0.04, seed=41)
    B = np.random.default_rng(42).uniform(-1, 1, (2500, 256)).astype(np.float32)
    return a, B, orc.spmm_csr(a, B, omp=True), orc.absprod_csr(a, B)


def test_coo_host_entry(b):
    import torch
    a, B, ref, den = _host_case()
    coo = orc.csr_to_coo(a)
    for variant in (0, 1, 2):
        C_h = torch.empty((a.M, 256), dtype=torch.float32).pin_memory()
        ms = b.spmm_coo_host(b.pinned(coo.rowIdxs), b.pinned(coo.colIdxs), b.pinned(coo.vals), a.M, a.K, b.pinned(B), C_h,
                             variant=variant)
        assert ms > 0
        assert rel_err(C_h, ref, den) <= TOL, variant


def test_sell_host_entry(b):
    import torch
    a, B, ref, den = _host_case()
    s = orc.csr_to_sell(a)
    C_h = torch.empty((a.M, 256), dtype=torch.float32).pin_memory()
    ms = b.spmm_sell_host(b.pinned(s.slicePtrs), b.pinned(s.colIdxs), b.pinned(s.vals), a.M, a.K, b.pinned(B), C_h)
    assert ms > 0
    assert rel_err(C_h, ref, den) <= TOL


def test_csr_host_devB_entry(b):
    import torch
    a, B, ref, den = _host_case()
    Bd = b.dev_f32(B)
    C_h = torch.empty((a.M, 256), dtype=torch.float32).pin_memory()
    ms = b.spmm_csr_host_devB(b.pinned(a.rowPtrs), b.pinned(a.colIdxs), b.pinned(a.vals), a.M, a.K, Bd, C_h)
    assert ms > 0
    assert rel_err(C_h, ref, den) <= TOL
    assert b.lib().cuspmm_host_pipeline_release(-1) == 0         # cached staging buffers can be given back ...
    ms = b.spmm_csr_host_devB(b.pinned(a.rowPtrs), b.pinned(a.colIdxs), b.pinned(a.vals), a.M, a.K, Bd, C_h)   # ... and come back
    assert rel_err(C_h, ref, den) <= TOL


def test_csr_host_many_panels(b, wl):
    """Enough rows and bytes for the MAXIMUM panel count of the host pipeline (16: >= 512 MiB of (col, val) and >= 16 x 6400
    rows); the 'one more tapered panel' rule used to push it to 17, past the event arrays."""
    import torch
    M, K, N = 110000, 1500, 128
    rp, ci, va = wl.gen_csr_device(M, K, 0.41, seed=51)
    nnz = int(ci.numel())
    assert nnz * 8 >= 16 * (32 << 20)
    Bd = wl.gen_dense_device(K, N, seed=52)
    C_h = torch.empty((M, N), dtype=torch.float32).pin_memory()
    ms = b.spmm_csr_host(rp.cpu().pin_memory(), ci.cpu().pin_memory(), va.cpu().pin_memory(), M, K, Bd.cpu().pin_memory(), C_h)
    assert ms > 0
    Bh = Bd.cpu().numpy()
    for r0 in (0, M // 2, M - 64):
        srp, sci, sva = wl.csr_sample_to_host(rp, ci, va, r0, r0 + 64)
        sub = orc.CSR(64, K, srp, sci, sva)
        assert orc.max_rel_err(C_h[r0:r0 + 64].numpy(), orc.spmm_csr(sub, Bh), orc.absprod_csr(sub, Bh)) <= TOL


@pytest.mark.parametrize("bs,variant", [(16, 1), (16, 2), (32, 3), (4, 1)])
def test_bsr_host_entry(b, bs, variant):
    import torch
    a = random_csr(1024, 768, 0.02, seed=61)
    bsr = orc.csr_to_bsr(a, bs, bs)
    N = 256
    B = np.random.default_rng(62).uniform(-1, 1, (bsr.K, N)).astype(np.float32)
    C_h = torch.empty((bsr.M, N), dtype=torch.float32).pin_memory()
    ms = b.spmm_bsr_host(b.pinned(bsr.blockRowPtrs), b.pinned(bsr.blockColIdxs), b.pinned(bsr.blocks), bsr.M // bs, bs, bs,
                         bsr.K, b.pinned(B), C_h, variant=variant)
    assert ms > 0
    den = orc.absprod_csr(orc.csr_from_dense(orc.to_dense(bsr)), B)
    if variant == 1:
        assert orc.max_rel_err(C_h.numpy(), orc.spmm_bsr(bsr, B), den) <= TOL
    else:
        rnd = orc.bf16_round if variant == 2 else orc.fp16_round
        rb = orc.BSR(bsr.M, bsr.K, bs, bs, bsr.blockRowPtrs, bsr.blockColIdxs, rnd(bsr.blocks))
        assert orc.max_rel_err(C_h.numpy(), orc.spmm_bsr(rb, rnd(B)), den) <= TC_TOL


# ------------------------------------------------------------------ multi-GPU plans, every format
def _gpu_counts():
    import torch
    n = torch.cuda.device_count()
    return sorted({1, min(2, n), n})


def test_mgpu_coo_sell_bsr(b):
    import torch
    a = random_csr(4100, 1300, 0.05, seed=71, skew=True)
    N = 128
    B = np.random.default_rng(72).uniform(-1, 1, (1300, N)).astype(np.float32)
    ref, den = orc.spmm_csr(a, B, omp=True), orc.absprod_csr(a, B)
    coo, s = orc.csr_to_coo(a), orc.csr_to_sell(a)
    for n in _gpu_counts():
        for fmt, arrs in (("coo", (coo.rowIdxs, coo.colIdxs, coo.vals)), ("sell", (s.slicePtrs, s.colIdxs, s.vals))):
            plan = b.MgpuPlan(n, arrs[0], arrs[1], arrs[2], a.M, a.K, N, fmt=fmt)
            try:
                sp = plan.splits()
                assert sp[0] == 0 and sp[-1] == a.M and (np.diff(sp.astype(np.int64)) >= 0).all()
                if fmt == "sell":
                    assert all(int(x) % 32 == 0 for x in sp[:-1])
                assert int(plan.counts().sum()) == len(arrs[1])
                plan.set_B(B)
                for gather in (False, True):
                    assert plan.run(variant=0, gather=gather, iters=1) > 0
                    assert orc.max_rel_err(plan.get_C(), ref, den) <= TOL, (fmt, n, gather)
            finally:
                plan.close()
        # BSR: fp32 kernels and the tensor-core plan per panel
        bsr = orc.csr_to_bsr(a, 16, 16)
        Bp = np.zeros((bsr.K, N), np.float32)
        Bp[:a.K] = B
        plan = b.MgpuPlan(n, bsr.blockRowPtrs, bsr.blockColIdxs, bsr.blocks, bsr.M // 16, bsr.K, N, fmt="bsr", br=16, bc=16)
        try:
            assert all(int(x) % 16 == 0 for x in plan.splits())
            plan.set_B(Bp)
            dn = orc.absprod_csr(orc.csr_from_dense(orc.to_dense(bsr)), Bp)
            plan.run(variant=1, gather=False, iters=1)
            assert orc.max_rel_err(plan.get_C(), orc.spmm_bsr(bsr, Bp), dn) <= TOL
            plan.run(variant=2, gather=(n > 1), iters=2)
            rb = orc.BSR(bsr.M, bsr.K, 16, 16, bsr.blockRowPtrs, bsr.blockColIdxs, orc.bf16_round(bsr.blocks))
            assert orc.max_rel_err(plan.get_C(), orc.spmm_bsr(rb, orc.bf16_round(Bp)), dn) <= TC_TOL
        finally:
            plan.close()
    torch.cuda.set_device(0)


# ------------------------------------------------------------------ BASELINE configs at their stated sizes
def _sampled_rows_and_checksum(b, wl, rp, ci, va, M, K, Bd, C, nrows=6):
    import torch
    Bh = Bd.cpu().numpy()
    for r in sorted(set(int(x) for x in np.linspace(0, M - 1, num=nrows))):
        srp, sci, sva = wl.csr_sample_to_host(rp, ci, va, r, r + 1)
        a = orc.CSR(1, K, srp, sci, sva)
        err = orc.max_rel_err(C[r:r + 1].cpu().numpy(), orc.spmm_csr(a, Bh), orc.absprod_csr(a, Bh))
        assert err <= TOL, (r, err)
    w = torch.zeros(K, dtype=torch.float64, device="cuda")
    w.index_add_(0, ci.to(torch.int64), va.to(torch.float64))
    expect = w @ Bd.to(torch.float64)
    scale = w.abs() @ Bd.abs().to(torch.float64) + 1e-30
    assert ((C.to(torch.float64).sum(dim=0) - expect).abs() / scale).max().item() < 1e-6


def test_cfg2_medium_4096_all_formats(b, wl):
    """BASELINE configs[1]: 4096^2, 90 % sparse, N = 512: CSR vs COO vs ELL, every selector path, against the oracle."""
    import torch
    M = K = 4096
    N = 512
    rp, ci, va = wl.gen_csr_device(M, K, 0.10, seed=618)
    Bd = wl.gen_dense_device(K, N, seed=619)
    c0 = b.spmm_csr(rp, ci, va, M, K, Bd, variant=0)
    _sampled_rows_and_checksum(b, wl, rp, ci, va, M, K, Bd, c0)
    for v in (1, 2, 3, 5, 7):
        assert (b.spmm_csr(rp, ci, va, M, K, Bd, variant=v) == c0).all().item(), v
    rows = torch.repeat_interleave(torch.arange(M, device="cuda", dtype=torch.int32), (rp[1:] - rp[:-1]).to(torch.int64))
    for v in (0, 1, 2):
        assert (b.spmm_coo(rows, ci, va, M, K, Bd, variant=v) == c0).all().item(), v
    sp, sc, sv = b.csr_to_sell(rp, ci, va, M)
    for v in (0, 1, 2, 3, 4, 5):
        assert (b.spmm_sell(sp, sc, sv, M, K, Bd, variant=v) == c0).all().item(), v


@pytest.mark.parametrize("density", [0.50, 0.10, 0.01])
@pytest.mark.parametrize("N", [128, 512, 2048])
def test_cfg3_sparsity_grid_csr_ell(b, wl, density, N):
    """BASELINE configs[2]: 4000^2 (test/sparsity.sh shape), sparsity 50 / 90 / 99 %, N = 128 / 512 / 2048, CSR and ELL."""
    M = K = 4000
    rp, ci, va = wl.gen_csr_device(M, K, density, seed=700 + N)
    Bd = wl.gen_dense_device(K, N, seed=701)
    c0 = b.spmm_csr(rp, ci, va, M, K, Bd, variant=0)
    _sampled_rows_and_checksum(b, wl, rp, ci, va, M, K, Bd, c0)
    sp, sc, sv = b.csr_to_sell(rp, ci, va, M)
    ce = b.spmm_sell(sp, sc, sv, M, K, Bd, variant=0)
    assert (ce == c0).all().item()           # same terms, same order: bit-identical across the two formats


@pytest.mark.parametrize("bs,dtype", [(16, "bf16"), (32, "bf16"), (16, "fp16"), (32, "fp16")])
def test_cfg4_bsr_tcgen05_full_size(b, wl, bs, dtype):
    """BASELINE configs[3]: large_25605 BSR, blocks of 16 / 32 (M, K padded to 25616 / 25632), 10 % of the blocks stored,
    bf16 / fp16 blocks on tcgen05, fp32 accumulate: sampled block rows against the oracle on the rounded operands, and the
    fp64 checksum of all of C against (1^T A_rounded) B_rounded."""
    import torch
    M = K = 25605
    N = 512
    g = torch.Generator(device="cuda"); g.manual_seed(618)
    nbr, nbc = (M + bs - 1) // bs, (K + bs - 1) // bs
    mask = torch.rand((nbr, nbc), generator=g, device="cuda") < 0.10
    brp = torch.zeros(nbr + 1, dtype=torch.int64, device="cuda"); brp[1:] = torch.cumsum(mask.sum(dim=1, dtype=torch.int64), 0)
    bci = mask.nonzero(as_tuple=False)[:, 1].to(torch.int32)
    nb = int(bci.numel())
    blocks = torch.rand(nb * bs * bs, generator=g, device="cuda") * 2 - 1
    brp = brp.to(torch.int32)
    Kp = nbc * bs
    assert Kp in (25616, 25632)
    Bd = wl.gen_dense_device(Kp, N, seed=619)
    plan = b.BsrTcPlan(brp, bci, blocks, nbr, bs, Kp, N, dtype=dtype)
    plan.prepare_B(Bd)
    C = plan.run()
    plan.close()
    rnd = orc.bf16_round if dtype == "bf16" else orc.fp16_round
    Bh = rnd(Bd.cpu().numpy())
    for R in (0, nbr // 2, nbr - 1):
        i0, i1 = int(brp[R].item()), int(brp[R + 1].item())
        sub = orc.BSR(bs, Kp, bs, bs, np.array([0, i1 - i0], np.uint32), bci[i0:i1].cpu().numpy().view(np.uint32).copy(),
                      rnd(blocks[i0 * bs * bs:i1 * bs * bs].cpu().numpy()))
        ref = orc.spmm_bsr(sub, Bh)
        den = orc.absprod_csr(orc.csr_from_dense(np.abs(orc.to_dense(sub))), np.abs(Bh))
        err = orc.max_rel_err(C[R * bs:(R + 1) * bs].cpu().numpy(), ref, den)
        assert err <= TC_TOL, (R, err)
    # checksum: column sums of A (rounded) times B (rounded) in fp64
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float16
    blk_r = blocks.to(tdt).to(torch.float64).view(nb, bs, bs)
    w = torch.zeros(Kp, dtype=torch.float64, device="cuda")
    cols = (bci.to(torch.int64)[:, None] * bs + torch.arange(bs, device="cuda")[None, :]).reshape(-1)
    w.index_add_(0, cols, blk_r.sum(dim=1).reshape(-1))
    wa = torch.zeros(Kp, dtype=torch.float64, device="cuda")
    wa.index_add_(0, cols, blk_r.abs().sum(dim=1).reshape(-1))
    Br = Bd.to(tdt).to(torch.float64)
    expect, scale = w @ Br, wa @ Br.abs() + 1e-30
    assert ((C.to(torch.float64).sum(dim=0) - expect).abs() / scale).max().item() < 1e-6


def test_cusparse_blocked_ell_baseline_agrees_with_oracle(b):
    """the Blocked-ELL vendor baseline (the reference's ELL descriptor throws) -- and, like every
    baseline here, its result is checked."""
    import torch
    a = random_csr(1024, 768, 0.03, seed=81)
    bsr = orc.csr_to_bsr(a, 16, 16)
    N = 128
    B = np.random.default_rng(82).uniform(-1, 1, (bsr.K, N)).astype(np.float32)
    out = torch.zeros((bsr.M, N), device="cuda")
    avg, mn, w = b.cusparse_spmm_blockedell(b.dev_u32(bsr.blockRowPtrs), b.dev_u32(bsr.blockColIdxs), b.dev_f32(bsr.blocks),
                                            bsr.M // 16, bsr.K // 16, 16, b.dev_f32(B), out, warmup=1, iters=2)
    assert avg > 0 and mn > 0
    assert w == int(np.diff(bsr.blockRowPtrs.astype(np.int64)).max())
    den = orc.absprod_csr(orc.csr_from_dense(orc.to_dense(bsr)), B)
    # cuSPARSE runs fp32 Blocked-ELL on TF32 tensor cores (10-bit mantissa operands): observed 3.4e-4 component-wise
    assert orc.max_rel_err(out.cpu().numpy(), orc.spmm_bsr(bsr, B), den) <= 2e-3


def test_unsorted_rows_are_refused_under_the_debug_guard():
    """a legal CSR with unsorted rows must not silently go through the staged kernels.  With
    CUSPMM_CHECK_SORTED set (read once per process -> child process) the call fails with CUSPMM_ERR_INVALID; the
    order-agnostic variant 1 still computes the right product."""
    import os
    import subprocess
    import sys
    code = """
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from __graft_entry__ import load_package
from conftest import random_csr
from oracle import oracle as orc
b = load_package().binding
a = random_csr(2048, 1024, 0.1, seed=3)
p0 = int(a.rowPtrs[100]); a.colIdxs[p0], a.colIdxs[p0 + 1] = a.colIdxs[p0 + 1], a.colIdxs[p0]; a.vals[p0], a.vals[p0 + 1] = a.vals[p0 + 1], a.vals[p0]
B = np.random.default_rng(1).uniform(-1, 1, (1024, 512)).astype(np.float32)
rp, ci, va, Bd = b.dev_u32(a.rowPtrs), b.dev_u32(a.colIdxs), b.dev_f32(a.vals), b.dev_f32(B)
for v in (3, 5, 7, 0):
    try:
        b.spmm_csr(rp, ci, va, a.M, a.K, Bd, variant=v)
        print("variant", v, "ran"); sys.exit(1)
    except b.CuspmmError as e:
        assert "status 1" in str(e) and "ascending" in str(e), str(e)
got = b.spmm_csr(rp, ci, va, a.M, a.K, Bd, variant=1).cpu().numpy()
a.colIdxs[p0], a.colIdxs[p0 + 1] = a.colIdxs[p0 + 1], a.colIdxs[p0]; a.vals[p0], a.vals[p0 + 1] = a.vals[p0 + 1], a.vals[p0]
assert orc.max_rel_err(got, orc.spmm_csr(a, B), orc.absprod_csr(a, B)) <= 1e-5
print("guard ok")
""" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, CUSPMM_CHECK_SORTED="1"), capture_output=True, text=True,
                       timeout=600)
    assert p.returncode == 0 and "guard ok" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
