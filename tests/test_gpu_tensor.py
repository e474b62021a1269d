"""CSR variant 8 (tensor cores: A tiles made dense in shared memory, tcgen05.mma with a tf32 + bf16 three-product split,
spmm_csr_tc.cu) against the oracle, through the C ABI.

The kernel does not compute in fp32 FMAs, so nothing here is bit-exact: the bound is the north_star tolerance
max |C - Cref| / (|A||B|) <= 1e-5, asserted for
  * ragged shapes (M, K, N not multiples of the 256 x 16 x 256 tiles; N without 128-bit alignment; unaligned colIdxs / vals);
  * the cases the error analysis names as worst: rows with ONE entry (no averaging of the split error, bound 2^-17 = 7.6e-6)
    and long sums of same-sign terms (the tensor core truncates when it accumulates: bias ~2^-25.5 per MMA step, bounded by
    draining the accumulators every 64 chunks);
  * integer-valued operands, where every product is exact in the tf32 main product: bit-exact against the oracle;
  * non-finite values: Inf / NaN in B must not leak into rows that never reference them (device-side reroute to the plain
    fp32 kernel), Inf / NaN in A stay in their row;
  * the BASELINE configuration (25605^2, 90 % sparse, N = 512): sampled rows against the oracle, fp64 column checksum, and
    agreement with the fp32 staged kernel."""
import importlib

import numpy as np
import pytest

from conftest import random_csr
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

TOL = 1e-5
V = 8


@pytest.fixture(scope="module")
def b():
    import torch
    assert torch.cuda.is_available()
    from __graft_entry__ import load_package
    pkg = load_package()
    pkg.lib()
    pkg.binding.set_csr_tensor_mode(1)
    return pkg.binding


@pytest.fixture(scope="module")
def wl(b):
    return importlib.import_module("cuspmm_b200.workloads")


def dev_csr(b, a):
    return b.dev_u32(a.rowPtrs), b.dev_u32(a.colIdxs), b.dev_f32(a.vals)


def check(b, a, B, tol=TOL):
    ref, den = orc.spmm_csr(a, B, omp=True), orc.absprod_csr(a, B)
    rp, ci, va = dev_csr(b, a)
    got = b.spmm_csr(rp, ci, va, a.M, a.K, b.dev_f32(B), variant=V).cpu().numpy()
    assert np.isfinite(got).all()
    err = orc.max_rel_err(got, ref, den)
    assert err <= tol, err
    return got, ref


@pytest.mark.parametrize("M,K,N,d,skew", [
    (1, 1, 1, 1.0, False), (1, 1, 512, 1.0, False), (7, 15, 4, 0.5, False), (57, 129, 512, 0.10, False),
    (255, 16, 256, 0.3, False), (256, 17, 257, 0.3, False), (257, 31, 260, 0.3, False), (300, 500, 512, 0.5, False),
    (1000, 4096, 512, 0.01, False), (777, 900, 1536, 0.30, False), (2000, 1500, 512, 0.03, True), (300, 257, 1024, 0.0, False),
    (513, 2100, 130, 0.2, False), (64, 333, 21, 0.4, False), (3000, 1111, 7, 0.1, True), (5000, 3000, 512, 0.05, False),
    (600, 20000, 128, 0.02, False)])
def test_tensor_kernel_vs_oracle(b, M, K, N, d, skew):
    a = random_csr(M, K, d, seed=300 + M + N, skew=skew)
    B = np.random.default_rng(5).uniform(-1, 1, (K, N)).astype(np.float32)
    check(b, a, B)


def _random_cases(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        M = int(rng.integers(1, 6000))
        K = int(rng.integers(1, 6000))
        N = int(rng.choice([3, 64, 100, 256, 300, 512, 512, 768, 1024]))
        d = float(rng.choice([0.001, 0.02, 0.1, 0.3, 0.6, 0.9, 1.0]))
        if M * K * d > 6e6:
            d = 6e6 / (M * K)
        out.append((M, K, N, d, bool(rng.integers(0, 2)), 4000 + i))
    return out


@pytest.mark.parametrize("M,K,N,d,skew,seed", _random_cases(16, 77))
def test_random_shapes(b, M, K, N, d, skew, seed):
    """Drawn shapes: partial tiles in every dimension, several column tiles, tiles cut along K between CTA pairs, several accumulator
    drains per tile, full rows (the slow path of the entry ring), skewed row lengths."""
    a = random_csr(M, K, d, seed=seed, skew=skew)
    B = np.random.default_rng(seed + 1).uniform(-1, 1, (K, N)).astype(np.float32)
    check(b, a, B)


def test_empty_rows_and_empty_matrix(b):
    a = random_csr(700, 300, 0.05, seed=9)
    keep = np.ones(a.M, bool)
    keep[::3] = False                       # every third row empty
    lens = np.diff(a.rowPtrs.astype(np.int64)) * keep
    rp = np.zeros(a.M + 1, np.uint32)
    rp[1:] = np.cumsum(lens)
    sel = np.repeat(keep, np.diff(a.rowPtrs.astype(np.int64)))
    a2 = orc.CSR(a.M, a.K, rp, a.colIdxs[sel], a.vals[sel])
    B = np.random.default_rng(6).uniform(-1, 1, (a.K, 512)).astype(np.float32)
    got, _ = check(b, a2, B)
    assert not got[::3].any()
    z = orc.CSR(100, 50, np.zeros(101, np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.float32))
    got, _ = check(b, z, B[:50])
    assert not got.any()


def test_integer_operands_are_exact(b):
    """|a|, |b| <= 8 integers: tf32 holds them exactly, the remainders are zero, every partial sum is an integer < 2^24."""
    a = random_csr(900, 1300, 0.2, seed=11, vals="int")
    B = np.random.default_rng(12).integers(-8, 9, (1300, 384)).astype(np.float32)
    got, ref = check(b, a, B)
    np.testing.assert_array_equal(got, ref)


def test_unaligned_arrays_and_leading_dimensions(b):
    """colIdxs / vals that are not 16-byte aligned (the per-entry cp.async path), B and C with padding between rows."""
    import torch
    a = random_csr(400, 700, 0.1, seed=13)
    B = np.random.default_rng(14).uniform(-1, 1, (700, 200)).astype(np.float32)
    ref, den = orc.spmm_csr(a, B), orc.absprod_csr(a, B)
    rp = b.dev_u32(a.rowPtrs)
    ci = torch.zeros(a.nnz + 3, dtype=torch.int32, device="cuda")
    va = torch.zeros(a.nnz + 3, dtype=torch.float32, device="cuda")
    ci[3:] = b.dev_u32(a.colIdxs)
    va[3:] = b.dev_f32(a.vals)
    Bp = torch.zeros((700, 203), dtype=torch.float32, device="cuda")
    Bp[:, :200] = b.dev_f32(B)
    Cp = torch.full((400, 209), 7.0, dtype=torch.float32, device="cuda")
    got = b.spmm_csr(rp, ci[3:], va[3:], a.M, a.K, Bp[:, :200], variant=V, out=Cp[:, :200], nnz=a.nnz)
    assert orc.max_rel_err(got.cpu().numpy(), ref, den) <= TOL
    assert (Cp[:, 200:] == 7.0).all().item()            # nothing written past N


def test_single_entry_rows_split_error_bound(b):
    """No averaging: C[r, n] = a_r * B[k_r, n].  The three-product split is within 2^-17 of the exact product (each of the two
    correction products carries the bf16 rounding, 2^-8, of a factor that multiplies a tf32 remainder, 2^-11; observed: 4.4e-6)."""
    rng = np.random.default_rng(15)
    M, K, N = 4096, 999, 256
    cols = rng.integers(0, K, M).astype(np.uint32)
    vals = (rng.uniform(1, 2, M) * 2.0 ** rng.integers(-20, 20, M) * rng.choice([-1, 1], M)).astype(np.float32)
    a = orc.CSR(M, K, np.arange(M + 1, dtype=np.uint32), cols, vals)
    B = (rng.uniform(1, 2, (K, N)) * 2.0 ** rng.integers(-20, 20, (K, N)) * rng.choice([-1, 1], (K, N))).astype(np.float32)
    check(b, a, B, tol=2.0 ** -17)


@pytest.mark.parametrize("K", [4096, 40000])
def test_same_sign_long_sums(b, K):
    """All terms positive: the truncating accumulation of the tensor core is a bias, not noise.  Drained every 64 chunks."""
    a = random_csr(300, K, 0.25, seed=16)
    a = orc.CSR(a.M, a.K, a.rowPtrs, a.colIdxs, np.abs(a.vals) + np.float32(0.5))
    B = (np.random.default_rng(17).uniform(0.5, 1.5, (K, 256))).astype(np.float32)
    check(b, a, B)


def test_non_finite_B_does_not_leak(b):
    import torch
    a = random_csr(600, 400, 0.05, seed=18)
    B = np.random.default_rng(19).uniform(-1, 1, (400, 256)).astype(np.float32)
    B[7, 3] = np.inf
    B[200, :] = np.nan
    rp, ci, va = dev_csr(b, a)
    got = b.spmm_csr(rp, ci, va, a.M, a.K, b.dev_f32(B), variant=V).cpu().numpy()
    ref = b.spmm_csr(rp, ci, va, a.M, a.K, b.dev_f32(B), variant=1).cpu().numpy()
    dense = orc.to_dense(a)
    touched = (dense[:, 7] != 0) | (dense[:, 200] != 0)
    assert np.isfinite(got[~touched]).all() and (~touched).sum() > 100
    np.testing.assert_allclose(got[~touched], ref[~touched], rtol=1e-4, atol=1e-5)
    assert np.isnan(got[dense[:, 200] != 0]).all()
    np.testing.assert_array_equal(np.isfinite(got), np.isfinite(ref))
    # and the next call with a finite B is served by the tensor kernel again (the flag lives in the per-call workspace)
    B[7, 3] = 1.0
    B[200, :] = 0.5
    check(b, a, B)


def test_non_finite_A_stays_in_its_row(b):
    a = random_csr(500, 300, 0.1, seed=20)
    vals = a.vals.copy()
    r_inf, r_nan = 10, 333
    vals[a.rowPtrs[r_inf]] = np.inf
    vals[a.rowPtrs[r_nan]] = np.nan
    a2 = orc.CSR(a.M, a.K, a.rowPtrs, a.colIdxs, vals)
    B = np.random.default_rng(21).uniform(0.5, 1.0, (300, 128)).astype(np.float32)
    rp, ci, va = dev_csr(b, a2)
    got = b.spmm_csr(rp, ci, va, a.M, a.K, b.dev_f32(B), variant=V).cpu().numpy()
    assert np.isinf(got[r_inf]).all() and (got[r_inf] > 0).all()
    assert np.isnan(got[r_nan]).all()
    rest = np.ones(a.M, bool)
    rest[[r_inf, r_nan]] = False
    ref, den = orc.spmm_csr(a, B), orc.absprod_csr(a, B)
    assert orc.max_rel_err(got[rest], ref[rest], den[rest]) <= TOL


def test_huge_and_tiny_magnitudes(b):
    """Values near the ends of the fp32 range: the split must neither overflow nor lose the result."""
    rng = np.random.default_rng(22)
    a = random_csr(300, 200, 0.1, seed=23)
    big = orc.CSR(a.M, a.K, a.rowPtrs, a.colIdxs, (a.vals * np.float32(2.0 ** 100)).astype(np.float32))
    B = (rng.uniform(-1, 1, (200, 128)) * 2.0 ** 20).astype(np.float32)
    check(b, big, B)
    small = orc.CSR(a.M, a.K, a.rowPtrs, a.colIdxs, (a.vals * np.float32(2.0 ** -60)).astype(np.float32))
    check(b, small, (B * np.float32(2.0 ** -50)).astype(np.float32))


def test_baseline_config_full_size(b, wl):
    """BASELINE configs: 25605^2, 90 % sparse, N = 512 -- sampled rows vs the oracle, fp64 column checksum, and the fp32 kernel."""
    import torch
    M = K = 25605
    N = 512
    rp, ci, va = wl.gen_csr_device(M, K, 0.10, seed=618)
    Bd = wl.gen_dense_device(K, N, seed=619)
    C = torch.full((M, N), float("nan"), device="cuda")
    b.spmm_csr(rp, ci, va, M, K, Bd, variant=V, out=C)
    assert torch.isfinite(C).all().item()
    Bh = Bd.cpu().numpy()
    for r in sorted(set(int(x) for x in np.linspace(0, M - 1, num=8)) | {255, 256, 25599, 25600}):
        srp, sci, sva = wl.csr_sample_to_host(rp, ci, va, r, r + 1)
        a = orc.CSR(1, K, srp, sci, sva)
        err = orc.max_rel_err(C[r:r + 1].cpu().numpy(), orc.spmm_csr(a, Bh), orc.absprod_csr(a, Bh))
        assert err <= TOL, (r, err)
    w = torch.zeros(K, dtype=torch.float64, device="cuda")
    w.index_add_(0, ci.to(torch.int64), va.to(torch.float64))
    expect = w @ Bd.to(torch.float64)
    scale = w.abs() @ Bd.abs().to(torch.float64) + 1e-30
    assert ((C.to(torch.float64).sum(dim=0) - expect).abs() / scale).max().item() < 1e-6
    C5 = b.spmm_csr(rp, ci, va, M, K, Bd, variant=5)
    den = b.spmm_csr(rp, ci, va.abs(), M, K, Bd.abs(), variant=5)
    assert ((C - C5).abs() / den.clamp_min(1e-30)).max().item() <= 2e-6
    # a second run: same result within the arrival order of the few tiles that two CTAs share
    C2 = b.spmm_csr(rp, ci, va, M, K, Bd, variant=V)
    assert ((C - C2).abs() / den.clamp_min(1e-30)).max().item() <= 1e-6


# ------------------------------------------------------------------ the selector with the tensor kernel in it
def test_selector_rules(b):
    sel = b.csr_selected_variant
    assert sel(25605, 25605, 65571195, 512) == 8                 # BASELINE: 10 % dense
    assert sel(25605, 25605, 13111101, 512) != 8                 # 2 %: the fp32 kernels win
    assert sel(4096, 4096, 1678023, 512) == 8 and sel(4096, 4096, 838880, 512) != 8
    assert sel(3200, 25605, 8196778, 512) == 8                   # an 8-GPU row panel of the BASELINE matrix
    assert sel(300, 200, 6000, 512) != 8                         # small: fixed costs
    assert sel(25605, 25605, 65571195, 512, sell=True) == 8      # the kernel also reads the sliced-ELL layout
    assert sel(25605, 25605, 65571195, 510) in (1, 2, 4, 8)
    prev = b.set_csr_tensor_mode(0)
    try:
        assert prev == 1 and sel(25605, 25605, 65571195, 512) == 5
    finally:
        b.set_csr_tensor_mode(1)


def test_variant0_paths_with_tensor_selection(b, wl):
    """4096^2, 20 % dense, N = 512: variant 0 resolves to 8 for CSR, for COO (device COO -> CSR + selector) and inside the
    host-buffer and multi-GPU entries; everything within the tolerance of the oracle (sampled rows) and of the fp32 kernel."""
    import torch
    M = K = 4096
    N = 512
    rp, ci, va = wl.gen_csr_device(M, K, 0.20, seed=41)
    Bd = wl.gen_dense_device(K, N, seed=42)
    nnz = int(ci.numel())
    assert b.csr_selected_variant(M, K, nnz, N) == 8
    ref = b.spmm_csr(rp, ci, va, M, K, Bd, variant=1)
    den = b.spmm_csr(rp, ci, va.abs(), M, K, Bd.abs(), variant=1).clamp_min(1e-30)
    Bh = Bd.cpu().numpy()

    def close(C, what):
        assert ((C - ref).abs() / den).max().item() <= 2e-6, what
        assert not (C == ref).all().item(), what + ": bit-identical to the fp32 kernel -- the tensor kernel did not run"

    c0 = b.spmm_csr(rp, ci, va, M, K, Bd, variant=0)
    close(c0, "csr")
    for r in (0, 255, 256, 2047, 4095):
        srp, sci, sva = wl.csr_sample_to_host(rp, ci, va, r, r + 1)
        a = orc.CSR(1, K, srp, sci, sva)
        assert orc.max_rel_err(c0[r:r + 1].cpu().numpy(), orc.spmm_csr(a, Bh), orc.absprod_csr(a, Bh)) <= TOL
    rows = torch.repeat_interleave(torch.arange(M, device="cuda", dtype=torch.int32), (rp[1:] - rp[:-1]).to(torch.int64))
    close(b.spmm_coo(rows, ci, va, M, K, Bd, variant=0), "coo")
    C_h = torch.empty((M, N), dtype=torch.float32).pin_memory()
    b.spmm_csr_host(rp.cpu().pin_memory(), ci.cpu().pin_memory(), va.cpu().pin_memory(), M, K, Bd.cpu().pin_memory(), C_h)
    close(C_h.cuda(), "csr_host")
    n = min(2, torch.cuda.device_count())
    plan = b.MgpuPlan(n, rp.cpu().numpy().view(np.uint32), ci.cpu().numpy().view(np.uint32), va.cpu().numpy(), M, K, N)
    try:
        plan.set_B(Bh)
        plan.run(variant=0, gather=True, iters=1)
        close(torch.from_numpy(plan.get_C()).cuda(), "mgpu")
    finally:
        plan.close()
    torch.cuda.set_device(0)


# ------------------------------------------------------------------ the same kernel reading sliced ELL (ELL variant 6)
@pytest.mark.parametrize("M,K,N,d,skew", [(1, 1, 512, 1.0, False), (57, 129, 512, 0.10, False), (300, 500, 260, 0.5, False),
                                          (1000, 4096, 512, 0.01, False), (2000, 1500, 512, 0.03, True), (300, 257, 1024, 0.0, False),
                                          (3000, 2500, 21, 0.3, True), (5000, 3000, 512, 0.2, False)])
def test_tensor_kernel_on_sliced_ell(b, M, K, N, d, skew):
    """Rows of very different lengths inside a slice (skew), padding entries, slices past M, both builder layouts."""
    a = random_csr(M, K, d, seed=500 + M + N, skew=skew)
    B = np.random.default_rng(7).uniform(-1, 1, (K, N)).astype(np.float32)
    ref, den = orc.spmm_csr(a, B, omp=True), orc.absprod_csr(a, B)
    rp, ci, va = dev_csr(b, a)
    sp, sc, sv = b.csr_to_sell(rp, ci, va, M)
    got = b.spmm_sell(sp, sc, sv, M, K, b.dev_f32(B), variant=6).cpu().numpy()
    assert np.isfinite(got).all()
    assert orc.max_rel_err(got, ref, den) <= TOL


def test_sliced_ell_selector_and_non_finite_B(b, wl):
    """4096^2 at 20 %: ELL variant 0 resolves to the tensor kernel; a NaN row of B stays in the rows that reference it."""
    import torch
    M = K = 4096
    N = 512
    rp, ci, va = wl.gen_csr_device(M, K, 0.20, seed=43)
    Bd = wl.gen_dense_device(K, N, seed=44)
    sp, sc, sv = b.csr_to_sell(rp, ci, va, M)
    assert b.csr_selected_variant(M, K, int(sc.numel()), N, sell=True) == 8
    ref = b.spmm_csr(rp, ci, va, M, K, Bd, variant=1)
    den = b.spmm_csr(rp, ci, va.abs(), M, K, Bd.abs(), variant=1).clamp_min(1e-30)
    c0 = b.spmm_sell(sp, sc, sv, M, K, Bd, variant=0)
    assert ((c0 - ref).abs() / den).max().item() <= 2e-6
    assert not (c0 == ref).all().item()                          # the tensor kernel ran, not an fp32 one
    Bn = Bd.clone()
    Bn[100, :] = float("nan")
    cn = b.spmm_sell(sp, sc, sv, M, K, Bn, variant=6)
    rn = b.spmm_csr(rp, ci, va, M, K, Bn, variant=1)
    assert (torch.isnan(cn) == torch.isnan(rn)).all().item()
    ok = ~torch.isnan(rn)
    assert ((cn[ok] - rn[ok]).abs() / den[ok]).max().item() <= 2e-6


def test_mgpu_plans_with_the_tensor_kernel(b):
    """Multi-GPU plans (n = 1, and every count the box has) with the tensor kernel forced: CSR variant 8, COO 2 -> selector, sliced ELL 6,
    with and without the gather into GPU 0 (panels on other GPUs are computed locally and sent with one peer copy)."""
    import torch
    a = random_csr(4100, 1300, 0.25, seed=81, skew=False)
    N = 256
    B = np.random.default_rng(82).uniform(-1, 1, (1300, N)).astype(np.float32)
    ref, den = orc.spmm_csr(a, B, omp=True), orc.absprod_csr(a, B)
    coo, s = orc.csr_to_coo(a), orc.csr_to_sell(a)
    counts = sorted({1, min(2, torch.cuda.device_count()), torch.cuda.device_count()})
    for n in counts:
        for fmt, arrs, v in (("csr", (a.rowPtrs, a.colIdxs, a.vals), 8), ("sell", (s.slicePtrs, s.colIdxs, s.vals), 6)):
            plan = b.MgpuPlan(n, arrs[0], arrs[1], arrs[2], a.M, a.K, N, fmt=fmt)
            try:
                plan.set_B(B)
                for gather in (False, True):
                    assert plan.run(variant=v, gather=gather, iters=2) > 0
                    assert orc.max_rel_err(plan.get_C(), ref, den) <= TOL, (fmt, n, gather)
            finally:
                plan.close()
    torch.cuda.set_device(0)
