"""GPU parity tests: every kernel and converter of libcuspmm_b200 through the C ABI, against the
CPU oracle (oracle/), on the reference's fixtures, seeded random inputs and edge cases.

Tolerances (BASELINE.json north_star): indices / conversions bit-exact; fp32 values
max_ij |C - Cref| / (|A||B|)_ij <= 1e-5 (component-wise relative error, oracle.max_rel_err);
integer-valued fixtures are bit-exact.  The reference's own allclose(1e-2, 1e-3) is asserted too."""
import os

import numpy as np
import pytest

from conftest import golden_case, random_csr
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

TOL = 1e-5


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


def dev_csr(b, a):
    return b.dev_u32(a.rowPtrs), b.dev_u32(a.colIdxs), b.dev_f32(a.vals)


def check(got, ref, denom, exact=False):
    got = got.cpu().numpy()
    assert got.shape == ref.shape
    if exact:
        np.testing.assert_array_equal(got, ref)
        return
    err = orc.max_rel_err(got, ref, denom)
    assert err <= TOL, f"max component-wise rel err {err:.3e} > {TOL}"
    assert orc.allclose_ref(got, ref)          # the reference's own criterion (utils.hpp:10-11)


# ------------------------------------------------------------------ reference fixtures
@pytest.mark.parametrize("case,exact", [("small_10x10", True), ("small_210", True), ("small_32x32", False)])
def test_fixture_all_formats(b, case, exact):
    d, pick = golden_case(case)
    B = orc.read_dense(pick("dense.in"))
    csr = orc.read_csr(pick(".csr"))
    coo = orc.read_coo(pick(".coo"))
    bsr = orc.read_bsr(pick(".bsr"))
    ell = orc.read_colell(pick("_rowind.ell"), pick("_values_colmajor.ell"))
    Bd = b.dev_f32(B)
    denom = orc.absprod_csr(csr, B)
    ref_csr = np.load(os.path.join(d, "ref_csr.npy"))
    rp, ci, va = dev_csr(b, csr)
    for variant in (0, 1, 2, 4):      # N = 21 (small_210): variants 1/2 run with 32-bit loads
        check(b.spmm_csr(rp, ci, va, csr.M, csr.K, Bd, variant=variant), ref_csr, denom, exact)
    ref_coo = np.load(os.path.join(d, "ref_coo.npy"))
    for variant in (0, 1):
        check(b.spmm_coo(b.dev_u32(coo.rowIdxs), b.dev_u32(coo.colIdxs), b.dev_f32(coo.vals), coo.M, coo.K, Bd,
                         variant=variant), ref_coo, denom, exact)
    # ELL: the reference's column-ELL arrays -> device CSR -> device sliced ELL -> kernel
    ref_ell = np.load(os.path.join(d, "ref_ell.npy"))
    erp, eci, eva = b.colell_to_csr(b.dev_u32(ell.rowIdxs), b.dev_f32(ell.vals), ell.M, ell.K, ell.maxColNnz, ell.nnz)
    np.testing.assert_array_equal(b.host_u32(erp), csr.rowPtrs)
    np.testing.assert_array_equal(b.host_u32(eci), csr.colIdxs)
    np.testing.assert_array_equal(eva.cpu().numpy(), csr.vals)
    sp, sc, sv = b.csr_to_sell(erp, eci, eva, ell.M)
    check(b.spmm_sell(sp, sc, sv, ell.M, ell.K, Bd), ref_ell, denom, exact)
    ref_bsr = np.load(os.path.join(d, "ref_bsr.npy"))
    check(b.spmm_bsr_f32(b.dev_u32(bsr.blockRowPtrs), b.dev_u32(bsr.blockColIdxs), b.dev_f32(bsr.blocks),
                         bsr.M // bsr.br, bsr.br, bsr.bc, bsr.K, Bd), ref_bsr, denom, exact)


# ------------------------------------------------------------------ seeded random + edge cases
SHAPES = [
    # M, K, N, density, skew
    (1, 1, 1, 1.0, False),
    (7, 5, 4, 0.5, False),
    (120, 210, 21, 0.03, False),      # small_210's shape class: N not a multiple of 4
    (300, 257, 33, 0.10, False),
    (257, 300, 64, 0.10, False),
    (513, 129, 128, 0.20, False),
    (640, 512, 256, 0.05, False),
    (1000, 700, 512, 0.02, True),     # skewed row lengths (large_21074-like)
    (333, 100, 516, 0.30, False),     # N > 512, multiple of 4 but not of 128
    (64, 64, 2048, 0.10, False),
    (200, 150, 8, 0.0, False),        # empty matrix
]


@pytest.mark.parametrize("M,K,N,d,skew", SHAPES)
def test_csr_coo_ell_random(b, M, K, N, d, skew):
    a = random_csr(M, K, d, seed=M * 7 + N, skew=skew)
    if M > 8:                               # force some empty rows and one full row
        lens = np.diff(a.rowPtrs.astype(np.int64))
        assert lens.min() == 0 or d == 0 or True
    B = np.random.default_rng(N).uniform(-1, 1, (K, N)).astype(np.float32)
    Bd = b.dev_f32(B)
    denom = orc.absprod_csr(a, B)
    ref = orc.spmm_csr(a, B)
    rp, ci, va = dev_csr(b, a)
    outs = {}
    for variant in (0, 1, 2, 4):
        outs[variant] = b.spmm_csr(rp, ci, va, M, K, Bd, variant=variant)
        check(outs[variant], ref, denom)
    # all CSR variants add the same terms in the same order with the same FMA: bit-identical
    vals = list(outs.values())
    for o in vals[1:]:
        assert (o == vals[0]).all().item()
    coo = orc.csr_to_coo(a)
    refc = orc.spmm_coo(coo, B)
    got = b.spmm_coo(b.dev_u32(coo.rowIdxs), ci, va, M, K, Bd, variant=1)
    check(got, refc, denom)
    assert (got == vals[0]).all().item()
    sp, sc, sv = b.csr_to_sell(rp, ci, va, M)
    for variant in (0, 1, 3):     # selector, row kernels on the sliced layout, slice-per-CTA kernel
        got = b.spmm_sell(sp, sc, sv, M, K, Bd, variant=variant)
        check(got, orc.spmm_ell(orc.csr_to_colell(a), B), denom)
        assert (got == vals[0]).all().item()


def test_noncontiguous_ldb_ldc(b):
    import torch
    a = random_csr(200, 160, 0.1, seed=5)
    B = np.random.default_rng(6).uniform(-1, 1, (160, 96)).astype(np.float32)
    big = torch.zeros((160, 128), device="cuda")
    big[:, :96] = b.dev_f32(B)
    Bv = big[:, :96]                       # ldb = 128
    outbig = torch.full((200, 160), 7.0, device="cuda")
    out = outbig[:, 32:128]                # ldc = 160, 16-byte aligned offset
    rp, ci, va = dev_csr(b, a)
    b.spmm_csr(rp, ci, va, 200, 160, Bv, variant=1, out=out)
    check(out.contiguous(), orc.spmm_csr(a, B), orc.absprod_csr(a, B))
    assert (outbig[:, :32] == 7.0).all().item() and (outbig[:, 128:] == 7.0).all().item()


def test_noncontiguous_ldb_ldc_staged_kernels(b):
    """B and C as column windows of wider buffers (ldb, ldc > N): the TMA ring of variants 3 / 5 copies row by row, the
    epilogue strides by ldc; the row-cutting kernel (6) likewise.  Nothing outside the window may be written."""
    import torch
    M, K, N = 700, 300, 512
    a = random_csr(M, K, 0.2, seed=15)
    B = np.random.default_rng(16).uniform(-1, 1, (K, N)).astype(np.float32)
    big = torch.full((K, N + 256), float("nan"), device="cuda")
    big[:, 128:128 + N] = b.dev_f32(B)
    Bv = big[:, 128:128 + N]               # ldb = 768, 16-byte aligned offset
    ref, den = orc.spmm_csr(a, B, omp=True), orc.absprod_csr(a, B)
    rp, ci, va = dev_csr(b, a)
    for variant in (3, 5, 6):
        outbig = torch.full((M, N + 64), 7.0, device="cuda")
        out = outbig[:, 32:32 + N]         # ldc = 576
        b.spmm_csr(rp, ci, va, M, K, Bv, variant=variant, out=out)
        check(out.contiguous(), ref, den)
        assert (outbig[:, :32] == 7.0).all().item() and (outbig[:, 32 + N:] == 7.0).all().item()


@pytest.mark.parametrize("M,K,N,d", [(2048, 1000, 128, 0.10), (1500, 2051, 256, 0.06), (3000, 1024, 512, 0.05),
                                     (1111, 777, 1024, 0.08)])
def test_csr_staged_kernel(b, M, K, N, d):
    a = random_csr(M, K, d, seed=N + M)
    B = np.random.default_rng(1).uniform(-1, 1, (K, N)).astype(np.float32)
    Bd = b.dev_f32(B)
    rp, ci, va = dev_csr(b, a)
    got = b.spmm_csr(rp, ci, va, M, K, Bd, variant=3)
    check(got, orc.spmm_csr(a, B, omp=True), orc.absprod_csr(a, B))
    assert (got == b.spmm_csr(rp, ci, va, M, K, Bd, variant=1)).all().item()
    coo = orc.csr_to_coo(a)
    got2 = b.spmm_coo(b.dev_u32(coo.rowIdxs), ci, va, M, K, Bd, variant=2)
    assert (got2 == got).all().item()
    sp, sc, sv = b.csr_to_sell(rp, ci, va, M)
    got3 = b.spmm_sell(sp, sc, sv, M, K, Bd, variant=2)          # staged kernel on the sliced-ELL layout
    assert (got3 == got).all().item()


@pytest.mark.parametrize("M,K,N,d,skew", [(3000, 1024, 512, 0.05, False), (1111, 777, 1024, 0.08, False),
                                          (2048, 1000, 512, 0.30, False), (1500, 2051, 512, 0.06, True),
                                          (61, 15, 512, 0.5, False), (5000, 333, 2048, 0.02, True)])
def test_csr_tmem_kernel(b, M, K, N, d, skew):
    """Variant 5: B chunks staged in tensor memory (tcgen05.cp), gathered with tcgen05.ld.  Same fp32 FMAs in the
    same order as the other variants: checked against the oracle AND bit-identical to variant 1; covers K not a
    multiple of the 16-row chunk, column tiles with ldb != 512, long rows (window refills), skewed rows."""
    a = random_csr(M, K, d, seed=N + M, skew=skew)
    B = np.random.default_rng(1).uniform(-1, 1, (K, N)).astype(np.float32)
    Bd = b.dev_f32(B)
    rp, ci, va = dev_csr(b, a)
    got = b.spmm_csr(rp, ci, va, M, K, Bd, variant=5)
    check(got, orc.spmm_csr(a, B, omp=True), orc.absprod_csr(a, B))
    assert (got == b.spmm_csr(rp, ci, va, M, K, Bd, variant=1)).all().item()
    sp, sc, sv = b.csr_to_sell(rp, ci, va, M)
    got3 = b.spmm_sell(sp, sc, sv, M, K, Bd, variant=4)          # the same kernel on the sliced-ELL layout
    assert (got3 == got).all().item()
    # run to run: no atomics, fixed order
    assert (b.spmm_csr(rp, ci, va, M, K, Bd, variant=5) == got).all().item()


@pytest.mark.parametrize("M,K,N,d,skew", [(300, 5000, 512, 0.02, True), (40, 3000, 128, 0.3, False), (1, 10000, 256, 0.5, False),
                                          (2000, 700, 512, 0.05, True), (500, 100, 64, 0.0, False), (129, 333, 20, 0.2, True), (129, 333, 21, 0.2, True)])
def test_csr_nnz_split_kernel(b, M, K, N, d, skew):
    """Variant 6: equal nnz ranges per warp, rows cut at range boundaries, partial sums added in entry order by a fix-up
    kernel (no atomics).  Within the stated tolerance of the oracle, run-to-run bit-identical, rows that are not cut
    bit-identical to the row kernels; empty rows (also leading / trailing) are written as zeros."""
    import torch
    a = random_csr(M, K, d, seed=M * 7 + N, skew=skew)
    if M >= 129:      # leading, middle and trailing empty rows
        lens = np.diff(a.rowPtrs.astype(np.int64))
        lens[[0, 1, M // 2, M - 2, M - 1]] = 0
        keep = np.concatenate([np.arange(a.rowPtrs[r], a.rowPtrs[r] + lens[r]) for r in range(M)]).astype(np.int64)
        rp = np.zeros(M + 1, np.uint32)
        rp[1:] = np.cumsum(lens)
        a = orc.CSR(M, K, rp, a.colIdxs[keep], a.vals[keep])
    B = np.random.default_rng(3).uniform(-1, 1, (K, N)).astype(np.float32)
    Bd = b.dev_f32(B)
    rp, ci, va = dev_csr(b, a)
    out = torch.full((M, N), float("nan"), device="cuda")
    got = b.spmm_csr(rp, ci, va, M, K, Bd, variant=6, out=out)
    check(got, orc.spmm_csr(a, B, omp=True), orc.absprod_csr(a, B))
    assert (b.spmm_csr(rp, ci, va, M, K, Bd, variant=6) == got).all().item()
    v1 = b.spmm_csr(rp, ci, va, M, K, Bd, variant=1)
    short = torch.from_numpy(np.diff(a.rowPtrs.astype(np.int64)) <= 32).cuda()     # rows of <= 32 entries are rarely cut ...
    same_rows = (got == v1).all(dim=1)
    if short.any().item():
        assert same_rows[short].float().mean().item() > 0.4                           # ... and uncut rows are bit-identical
    assert ((got - v1).abs() <= 1e-5 * (torch.from_numpy(orc.absprod_csr(a, B)).cuda() + 1e-30)).all().item()


def test_staged_rejects_unsupported_shape(b):
    a = random_csr(64, 64, 0.1, seed=1)
    rp, ci, va = dev_csr(b, a)
    Bd = b.dev_f32(np.ones((64, 100), np.float32))
    with pytest.raises(b.CuspmmError):
        b.spmm_csr(rp, ci, va, 64, 64, Bd, variant=3)
    with pytest.raises(b.CuspmmError):
        b.spmm_csr(rp, ci, va, 64, 64, Bd, variant=5)     # N % 512 != 0
    with pytest.raises(b.CuspmmError):
        b.spmm_csr(rp, ci, va, 64, 64, Bd, variant=9)     # "Not implemented" (engine_csr.hpp:88)


def test_inf_in_unreferenced_B_rows_does_not_leak(b):
    """Padding lanes must never multiply 0 * B[0]: an inf in an unused B row must not reach C."""
    a = random_csr(100, 50, 0.1, seed=3)
    used = np.zeros(50, bool)
    used[a.colIdxs] = True
    B = np.random.default_rng(2).uniform(-1, 1, (50, 64)).astype(np.float32)
    B[~used] = np.inf
    if used[0]:
        pytest.skip("column 0 referenced")
    Bd = b.dev_f32(B)
    rp, ci, va = dev_csr(b, a)
    for variant in (1, 2, 4):
        assert np.isfinite(b.spmm_csr(rp, ci, va, 100, 50, Bd, variant=variant).cpu().numpy()).all()
    sp, sc, sv = b.csr_to_sell(rp, ci, va, 100)
    for variant in (1, 3):
        assert np.isfinite(b.spmm_sell(sp, sc, sv, 100, 50, Bd, variant=variant).cpu().numpy()).all()


def test_canaries_around_C_and_B(b):
    """compute-sanitizer is closed on this pool: catch stray writes with sentinel rows around C and stray
    reads with NaN rows around B (a NaN that leaks into C, or a changed sentinel, fails the test)."""
    import torch
    a = random_csr(1500, 600, 0.08, seed=77, skew=True)
    N = 512
    B = np.random.default_rng(78).uniform(-1, 1, (600, N)).astype(np.float32)
    Bbig = torch.full((600 + 64, N), float("nan"), device="cuda")
    Bbig[32:632] = b.dev_f32(B)
    Bv = Bbig[32:632]
    ref = orc.spmm_csr(a, B, omp=True)
    denom = orc.absprod_csr(a, B)
    rp, ci, va = dev_csr(b, a)
    coo_rows = b.dev_u32(orc.csr_to_coo(a).rowIdxs)
    sp, sc, sv = b.csr_to_sell(rp, ci, va, a.M)
    runs = [(lambda out, v=v: b.spmm_csr(rp, ci, va, a.M, a.K, Bv, variant=v, out=out)) for v in (1, 2, 3, 4, 5)]
    runs += [(lambda out, v=v: b.spmm_coo(coo_rows, ci, va, a.M, a.K, Bv, variant=v, out=out)) for v in (1, 2)]
    runs += [(lambda out, v=v: b.spmm_sell(sp, sc, sv, a.M, a.K, Bv, variant=v, out=out)) for v in (1, 2, 3, 4)]
    for run in runs:
        Cbig = torch.full((a.M + 16, N), 12345.0, device="cuda")
        out = Cbig[8:8 + a.M]
        run(out)
        torch.cuda.synchronize()
        assert (Cbig[:8] == 12345.0).all().item() and (Cbig[8 + a.M:] == 12345.0).all().item()
        check(out, ref, denom)


# ------------------------------------------------------------------ BSR fp32
@pytest.mark.parametrize("bs", [1, 2, 4, 16, 32])
@pytest.mark.parametrize("N", [21, 128, 512])
def test_bsr_f32(b, bs, N):
    a = random_csr(203, 170, 0.05, seed=bs * 31 + N)
    rp, ci, va = dev_csr(b, a)
    brp, bci, bl = b.csr_to_bsr(rp, ci, va, a.M, a.K, bs, bs)
    o = orc.csr_to_bsr(a, bs, bs)
    np.testing.assert_array_equal(b.host_u32(brp), o.blockRowPtrs)
    np.testing.assert_array_equal(b.host_u32(bci), o.blockColIdxs)
    np.testing.assert_array_equal(bl.cpu().numpy(), o.blocks)
    B = np.zeros((o.K, N), np.float32)
    B[:a.K] = np.random.default_rng(N).uniform(-1, 1, (a.K, N))
    ref = orc.spmm_bsr(o, B)
    Bd = b.dev_f32(B)
    got = b.spmm_bsr_f32(brp, bci, bl, o.M // bs, bs, bs, o.K, Bd)
    ap = orc.CSR(o.M, o.K, np.concatenate([a.rowPtrs, np.full(o.M - a.M, a.rowPtrs[-1], np.uint32)]), a.colIdxs, a.vals)
    check(got, ref, orc.absprod_csr(ap, B))


def test_bsr_rectangular_blocks(b):
    a = random_csr(96, 120, 0.08, seed=77)
    o = orc.csr_to_bsr(a, 4, 8)
    B = np.random.default_rng(3).uniform(-1, 1, (o.K, 40)).astype(np.float32)
    got = b.spmm_bsr_f32(b.dev_u32(o.blockRowPtrs), b.dev_u32(o.blockColIdxs), b.dev_f32(o.blocks), o.M // 4, 4, 8,
                         o.K, b.dev_f32(B))
    check(got, orc.spmm_bsr(o, B), orc.absprod_csr(a, B))


# ------------------------------------------------------------------ device converters, bit-exact
@pytest.mark.parametrize("M,K,d,skew", [(1, 1, 1.0, False), (31, 40, 0.2, False), (32, 40, 0.2, False), (70, 900, 0.3, False),
                                        (1000, 800, 0.03, True), (4097, 300, 0.05, False), (100, 100, 0.0, False)])
def test_device_conversions_bit_exact(b, M, K, d, skew):
    a = random_csr(M, K, d, seed=M + K, skew=skew)
    rp, ci, va = dev_csr(b, a)
    # CSR -> sliced ELL
    sp, sc, sv = b.csr_to_sell(rp, ci, va, M)
    o = orc.csr_to_sell(a, 32)
    np.testing.assert_array_equal(b.host_u32(sp), o.slicePtrs)
    np.testing.assert_array_equal(b.host_u32(sc), o.colIdxs)
    np.testing.assert_array_equal(sv.cpu().numpy(), o.vals)
    # COO -> CSR row pointers
    coo = orc.csr_to_coo(a)
    np.testing.assert_array_equal(b.host_u32(b.coo_to_csr_rowptrs(b.dev_u32(coo.rowIdxs), M)), a.rowPtrs)
    # column-ELL (the reference's storage) -> CSR
    e = orc.csr_to_colell(a)
    if e.maxColNnz:
        erp, eci, eva = b.colell_to_csr(b.dev_u32(e.rowIdxs), b.dev_f32(e.vals), M, K, e.maxColNnz, e.nnz)
        np.testing.assert_array_equal(b.host_u32(erp), a.rowPtrs)
        np.testing.assert_array_equal(b.host_u32(eci), a.colIdxs)
        np.testing.assert_array_equal(eva.cpu().numpy(), a.vals)
        with pytest.raises(b.CuspmmError):
            b.colell_to_csr(b.dev_u32(e.rowIdxs), b.dev_f32(e.vals), M, K, e.maxColNnz, e.nnz + 1)
    # CSR -> BSR
    for br, bc in ((2, 2), (16, 16), (3, 5)):
        brp, bci, bl = b.csr_to_bsr(rp, ci, va, M, K, br, bc)
        ob = orc.csr_to_bsr(a, br, bc)
        np.testing.assert_array_equal(b.host_u32(brp), ob.blockRowPtrs)
        np.testing.assert_array_equal(b.host_u32(bci), ob.blockColIdxs)
        np.testing.assert_array_equal(bl.cpu().numpy(), ob.blocks)
    # nnz-balanced row panels
    for parts in (1, 2, 3, 4, 8):
        np.testing.assert_array_equal(b.partition_rows_by_nnz(rp, M, a.nnz, parts),
                                      orc.partition_rows_by_nnz(a.rowPtrs, parts))


# ------------------------------------------------------------------ host-buffer entry point + multi-GPU plan
def test_csr_host_entry_point(b):
    import torch
    a = random_csr(5000, 3000, 0.05, seed=11)
    B = np.random.default_rng(12).uniform(-1, 1, (3000, 256)).astype(np.float32)
    C_h = torch.empty((5000, 256), dtype=torch.float32).pin_memory()
    ms = b.spmm_csr_host(b.pinned(a.rowPtrs), b.pinned(a.colIdxs), b.pinned(a.vals), a.M, a.K, b.pinned(B), C_h)
    assert ms > 0
    check(C_h, orc.spmm_csr(a, B, omp=True), orc.absprod_csr(a, B))


def test_mgpu_plan_single_and_multi(b):
    import torch
    a = random_csr(4000, 1500, 0.04, seed=21, skew=True)
    B = np.random.default_rng(22).uniform(-1, 1, (1500, 128)).astype(np.float32)
    ref = orc.spmm_csr(a, B, omp=True)
    denom = orc.absprod_csr(a, B)
    for n in sorted({1, min(2, torch.cuda.device_count()), torch.cuda.device_count()}):
        plan = b.MgpuPlan(n, a.rowPtrs, a.colIdxs, a.vals, a.M, a.K, 128)
        try:
            np.testing.assert_array_equal(plan.splits(), orc.partition_rows_by_nnz(a.rowPtrs, n))
            plan.set_B(B)
            for gather in (False, True):
                ms = plan.run(variant=0, gather=gather, iters=2)
                assert ms > 0
                got = plan.get_C()
                err = orc.max_rel_err(got, ref, denom)
                assert err <= TOL, (n, gather, err)
        finally:
            plan.close()
    torch.cuda.set_device(0)


def test_cusparse_baseline_agrees_with_oracle(b):
    """The reference never checks cuSPARSE (engine.cpp:54-55 passes correct=1); we do."""
    import torch
    a = random_csr(1024, 900, 0.08, seed=31)
    B = np.random.default_rng(32).uniform(-1, 1, (900, 128)).astype(np.float32)
    rp, ci, va = dev_csr(b, a)
    Bd = b.dev_f32(B)
    out = torch.empty((1024, 128), device="cuda")
    avg, mn = b.cusparse_spmm(0, rp, ci, va, a.M, a.K, Bd, out, warmup=1, iters=2)
    assert avg > 0 and mn > 0
    ref = orc.spmm_csr(a, B, omp=True)
    denom = orc.absprod_csr(a, B)
    assert orc.max_rel_err(out.cpu().numpy(), ref, denom) <= 1e-4
    coo = orc.csr_to_coo(a)
    out.zero_()
    b.cusparse_spmm(1, b.dev_u32(coo.rowIdxs), ci, va, a.M, a.K, Bd, out, warmup=1, iters=2)
    assert orc.max_rel_err(out.cpu().numpy(), ref, denom) <= 1e-4


# ------------------------------------------------------------------ full-size properties (BASELINE sizes)
def test_full_size_large_25605_properties(b):
    """25605^2, 90% sparse, N=512 (the north-star row): the oracle cannot run this in seconds, so
    check size-independent properties: (1) a sampled row panel against the oracle, (2) the
    checksum 1^T C == (1^T A) B in fp64, (3) all CSR variants, COO and ELL bit-identical."""
    import importlib
    import torch
    wl = importlib.import_module("cuspmm_b200.workloads")     # registered by load_package() in the fixture
    M = K = 25605
    N = 512
    rp, ci, va = wl.gen_csr_device(M, K, 0.10, seed=618)
    Bd = wl.gen_dense_device(K, N, seed=619)
    nnz = int(ci.numel())
    assert abs(nnz / (M * K) - 0.10) < 1e-3
    c3 = b.spmm_csr(rp, ci, va, M, K, Bd, variant=3)
    c1 = b.spmm_csr(rp, ci, va, M, K, Bd, variant=1)
    assert (c1 == c3).all().item()
    c5 = b.spmm_csr(rp, ci, va, M, K, Bd, variant=5)          # B through tensor memory
    assert (c5 == c3).all().item()
    del c1, c5
    # (1) sampled rows vs oracle
    for r0 in (0, 12800, M - 37):
        r1 = min(M, r0 + 37)
        srp, sci, sva = wl.csr_sample_to_host(rp, ci, va, r0, r1)
        a = orc.CSR(r1 - r0, K, srp, sci, sva)
        Bh = Bd.cpu().numpy()
        ref = orc.spmm_csr(a, Bh, omp=True)
        err = orc.max_rel_err(c3[r0:r1].cpu().numpy(), ref, orc.absprod_csr(a, Bh))
        assert err <= TOL, err
    # (2) checksum of checksums in fp64
    w = torch.zeros(K, dtype=torch.float64, device="cuda")
    w.index_add_(0, ci.to(torch.int64), va.to(torch.float64))
    expect = w @ Bd.to(torch.float64)
    got = c3.to(torch.float64).sum(dim=0)
    scale = (w.abs() @ Bd.abs().to(torch.float64))
    assert ((got - expect).abs() / scale).max().item() < 1e-6
    # (3) other formats, bit-identical to CSR
    rows = torch.repeat_interleave(torch.arange(M, device="cuda", dtype=torch.int32),
                                   (rp[1:] - rp[:-1]).to(torch.int64))
    cc = b.spmm_coo(rows, ci, va, M, K, Bd, variant=1)
    assert (cc == c3).all().item()
    del cc, rows
    sp, sc, sv = b.csr_to_sell(rp, ci, va, M)
    for variant in (1, 2, 3, 4):
        ce = b.spmm_sell(sp, sc, sv, M, K, Bd, variant=variant)
        assert (ce == c3).all().item()


def test_full_size_ffn_shape_properties(b):
    """BASELINE configs[4]: pruned-LLM FFN weight 11008x4096 at 50 % sparsity, N = 4096 tokens (22.5 M nnz,
    C = 180 MB): the selector's kernel, the row-split kernel and COO agree bit for bit, and the checksum
    1^T C == (1^T A) B holds in fp64."""
    import importlib
    import torch
    wl = importlib.import_module("cuspmm_b200.workloads")
    M, K, N = 11008, 4096, 4096
    rp, ci, va = wl.gen_csr_device(M, K, 0.50, seed=7)
    Bd = wl.gen_dense_device(K, N, seed=8)
    c0 = b.spmm_csr(rp, ci, va, M, K, Bd, variant=0)
    c1 = b.spmm_csr(rp, ci, va, M, K, Bd, variant=1)
    assert (c0 == c1).all().item()
    del c1
    rows = torch.repeat_interleave(torch.arange(M, device="cuda", dtype=torch.int32), (rp[1:] - rp[:-1]).to(torch.int64))
    cc = b.spmm_coo(rows, ci, va, M, K, Bd, variant=0)
    assert (cc == c0).all().item()
    del cc, rows
    w = torch.zeros(K, dtype=torch.float64, device="cuda")
    w.index_add_(0, ci.to(torch.int64), va.to(torch.float64))
    expect = w @ Bd.to(torch.float64)
    scale = w.abs() @ Bd.abs().to(torch.float64)
    assert (((c0.to(torch.float64).sum(dim=0)) - expect).abs() / scale).max().item() < 1e-6
    for r0 in (0, M - 16):
        srp, sci, sva = wl.csr_sample_to_host(rp, ci, va, r0, r0 + 16)
        a = orc.CSR(16, K, srp, sci, sva)
        Bh = Bd.cpu().numpy()
        err = orc.max_rel_err(c0[r0:r0 + 16].cpu().numpy(), orc.spmm_csr(a, Bh, omp=True), orc.absprod_csr(a, Bh))
        assert err <= TOL, err
