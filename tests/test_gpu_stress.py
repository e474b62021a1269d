"""Randomised GPU parity stress (seeded): random shapes / densities / skews through EVERY CSR, COO and ELL variant, the
host-buffer entry points and the converters, against the oracle.  Complements the fixed cases of test_gpu_parity.py /
test_gpu_round2.py: the shapes here are drawn, not chosen."""
import numpy as np
import pytest

from conftest import random_csr
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


def _cases(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        M = int(rng.integers(1, 3000))
        K = int(rng.integers(1, 2500))
        N = int(rng.choice([1, 7, 21, 64, 128, 200, 256, 512, 512, 512, 1024, 1536]))
        d = float(rng.choice([0.0, 0.002, 0.01, 0.05, 0.1, 0.3, 0.7]))
        out.append((M, K, N, d, bool(rng.integers(0, 2)), 1000 + i))
    return out


@pytest.mark.parametrize("M,K,N,d,skew,seed", _cases(24, 2026))
def test_random_shapes_all_variants(b, M, K, N, d, skew, seed):
    import torch
    a = random_csr(M, K, d, seed=seed, skew=skew)
    B = np.random.default_rng(seed + 1).uniform(-1, 1, (K, N)).astype(np.float32)
    ref, den = orc.spmm_csr(a, B, omp=True), orc.absprod_csr(a, B)
    rp, ci, va, Bd = b.dev_u32(a.rowPtrs), b.dev_u32(a.colIdxs), b.dev_f32(a.vals), b.dev_f32(B)
    base = None
    for v in (0, 1, 2, 3, 4, 5, 6, 7, 8):
        try:
            got = b.spmm_csr(rp, ci, va, M, K, Bd, variant=v)
        except b.CuspmmError as e:        # a variant may decline a shape it cannot tile; nothing else is acceptable
            assert "status 3" in str(e) and v in (3, 5, 7), (v, str(e))
            continue
        assert orc.max_rel_err(got.cpu().numpy(), ref, den) <= TOL, v
        if v not in (6, 8):               # variant 6 cuts rows, variant 8 runs on the tensor cores: same tolerance, different rounding
            base = got if base is None else base
            assert (got == base).all().item(), f"variant {v} is not bit-identical to the first variant that ran"
    coo = orc.csr_to_coo(a)
    for v in (0, 1, 2):
        got = b.spmm_coo(b.dev_u32(coo.rowIdxs), ci, va, M, K, Bd, variant=v)
        assert (got == base).all().item(), ("coo", v)
    sp, sc, sv = b.csr_to_sell(rp, ci, va, M)
    ores = orc.csr_to_sell(a)
    np.testing.assert_array_equal(b.host_u32(sp), ores.slicePtrs)
    np.testing.assert_array_equal(b.host_u32(sc), ores.colIdxs)
    for v in (0, 1, 2, 3, 4, 5):
        try:
            got = b.spmm_sell(sp, sc, sv, M, K, Bd, variant=v)
        except b.CuspmmError as e:
            assert "status 3" in str(e) and v in (2, 4, 5), (v, str(e))
            continue
        assert (got == base).all().item(), ("ell", v)
    # ELL 6 = the tensor-core kernel on the sliced layout: the tolerance, not bit-identity
    got = b.spmm_sell(sp, sc, sv, M, K, Bd, variant=6)
    assert orc.max_rel_err(got.cpu().numpy(), ref, den) <= TOL, "ell 6"
    # host-buffer entry points (pipelined H2D / kernel / D2H)
    C_h = torch.empty((M, N), dtype=torch.float32).pin_memory()
    b.spmm_csr_host(b.pinned(a.rowPtrs), b.pinned(a.colIdxs), b.pinned(a.vals), M, K, b.pinned(B), C_h)
    assert orc.max_rel_err(C_h.numpy(), ref, den) <= TOL
    if a.nnz:
        C_h.zero_()
        b.spmm_coo_host(b.pinned(coo.rowIdxs), b.pinned(coo.colIdxs), b.pinned(coo.vals), M, K, b.pinned(B), C_h)
        assert orc.max_rel_err(C_h.numpy(), ref, den) <= TOL
        C_h.zero_()
        b.spmm_sell_host(b.pinned(ores.slicePtrs), b.pinned(ores.colIdxs), b.pinned(ores.vals), M, K, b.pinned(B), C_h)
        assert orc.max_rel_err(C_h.numpy(), ref, den) <= TOL


@pytest.mark.parametrize("M,K,bs,d,seed", [(777, 500, 4, 0.02, 1), (1024, 1024, 16, 0.01, 2), (2500, 1300, 32, 0.003, 3),
                                           (333, 2000, 8, 0.05, 4), (64, 64, 16, 0.5, 5), (4096, 300, 2, 0.01, 6)])
def test_random_bsr_conversion_and_kernels(b, M, K, bs, d, seed):
    """Device CSR -> BSR (sort-free bitmap path) is bit-exact against the numpy restatement of scipy's tobsr, and the fp32
    BSR kernel reproduces spmmBSRCpu's order on the converted matrix."""
    a = random_csr(M, K, d, seed=seed, skew=bool(seed % 2))
    o = orc.csr_to_bsr(a, bs, bs)
    rp, ci, va = b.dev_u32(a.rowPtrs), b.dev_u32(a.colIdxs), b.dev_f32(a.vals)
    brp, bci, bbl = b.csr_to_bsr(rp, ci, va, M, K, bs, bs)
    np.testing.assert_array_equal(b.host_u32(brp), o.blockRowPtrs)
    np.testing.assert_array_equal(b.host_u32(bci), o.blockColIdxs)
    np.testing.assert_array_equal(bbl.cpu().numpy(), o.blocks)
    N = 96
    Bp = np.random.default_rng(seed).uniform(-1, 1, (o.K, N)).astype(np.float32)
    got = b.spmm_bsr_f32(brp, bci, bbl, o.M // bs, bs, bs, o.K, b.dev_f32(Bp)).cpu().numpy()
    den = orc.absprod_csr(orc.csr_from_dense(orc.to_dense(o)), Bp)
    assert orc.max_rel_err(got, orc.spmm_bsr(o, Bp), den) <= TOL
