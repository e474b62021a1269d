"""CPU-only: pin the oracle (oracle/spmm_oracle.c, oracle/oracle.py) against
 (a) the fixtures the reference commits (data/small_10x10, data/small_32x32),
 (b) files its own converter (utils/python_utils/convert_mtx.py) produced, and
 (c) outputs of the reference's own compiled CPU SpMM (oracle/_ref), stored by
     tests/golden/make_golden.py as ref_*.npy and, when the .so is present, run live."""
import os

import numpy as np
import pytest

from conftest import golden_case, random_csr
from oracle import oracle as orc

CASES = ["small_10x10", "small_32x32", "small_210"]


def load(case):
    d, pick = golden_case(case)
    B = orc.read_dense(pick("dense.in"))
    return d, pick, B


@pytest.mark.parametrize("case", ["small_10x10", "small_32x32"])
def test_converter_output_equals_committed_files(case):
    """What the reference ships == what its converter writes here (text identical)."""
    d, pick = golden_case(case)
    for name in os.listdir(os.path.join(d, "committed")):
        if name.endswith((".csr", ".coo", "dense.in")):
            a = open(os.path.join(d, "committed", name)).read()
            b = open(os.path.join(d, name)).read()
            assert a == b, name


@pytest.mark.parametrize("case", CASES)
def test_oracle_equals_reference_binary_outputs(case):
    d, pick, B = load(case)
    csr = orc.read_csr(pick(".csr"))
    coo = orc.read_coo(pick(".coo"))
    bsr = orc.read_bsr(pick(".bsr"))
    ell = orc.read_colell(pick("_rowind.ell"), pick("_values_colmajor.ell"))
    np.testing.assert_array_equal(orc.spmm_csr(csr, B), np.load(os.path.join(d, "ref_csr.npy")))
    np.testing.assert_array_equal(orc.spmm_coo(coo, B), np.load(os.path.join(d, "ref_coo.npy")))
    np.testing.assert_array_equal(orc.spmm_bsr(bsr, B), np.load(os.path.join(d, "ref_bsr.npy")))
    np.testing.assert_array_equal(orc.spmm_ell(ell, B), np.load(os.path.join(d, "ref_ell.npy")))


@pytest.mark.parametrize("case", ["small_10x10", "small_32x32"])
def test_oracle_matches_result_expect_and_coo_out(case):
    d, pick, B = load(case)
    csr = orc.read_csr(pick(".csr"))
    got = orc.spmm_csr(csr, B)
    expect = np.loadtxt(os.path.join(d, "committed", "result.expect"), dtype=np.float64)
    assert expect.shape == got.shape
    # result.expect is scipy A@B printed with %.10f (validate.py:31-96)
    np.testing.assert_allclose(got, expect, rtol=1e-6, atol=1e-6)
    for name in ("coo.out", "coo_cuda.out"):
        with open(os.path.join(d, "committed", name)) as f:
            rows = [ln.split() for ln in f.read().strip().split("\n")]
        if len(rows[0]) == 2 and len(rows) == got.shape[0] + 1:   # save2File header
            rows = rows[1:]
        out = np.array(rows, dtype=np.float64)
        # 6 significant digits (operator<< default precision, dense.cu:194-232)
        np.testing.assert_allclose(orc.spmm_coo(orc.read_coo(pick(".coo")), B), out, rtol=2e-5, atol=1e-4)


@pytest.mark.parametrize("case", CASES)
def test_all_formats_describe_the_same_matrix(case):
    d, pick, B = load(case)
    D = orc.to_dense(orc.read_csr(pick(".csr")))
    np.testing.assert_array_equal(orc.to_dense(orc.read_coo(pick(".coo"))), D)
    np.testing.assert_array_equal(orc.to_dense(orc.read_bsr(pick(".bsr"))), D)
    np.testing.assert_array_equal(
        orc.to_dense(orc.read_colell(pick("_rowind.ell"), pick("_values_colmajor.ell"))), D)


@pytest.mark.parametrize("case", CASES)
def test_conversion_oracles_reproduce_converter_files(case):
    """csr_to_coo / csr_to_bsr(1,1) / csr_to_colell == the arrays convert_mtx.py wrote."""
    d, pick, B = load(case)
    csr = orc.read_csr(pick(".csr"))
    coo = orc.read_coo(pick(".coo"))
    mine = orc.csr_to_coo(csr)
    for f in ("rowIdxs", "colIdxs", "vals"):
        np.testing.assert_array_equal(getattr(mine, f), getattr(coo, f))
    back = orc.coo_to_csr(coo)
    for f in ("rowPtrs", "colIdxs", "vals"):
        np.testing.assert_array_equal(getattr(back, f), getattr(csr, f))
    bsr = orc.read_bsr(pick(".bsr"))
    assert (bsr.br, bsr.bc) == (1, 1)          # convert_mtx.py:22 forces size = 1
    mine = orc.csr_to_bsr(csr, 1, 1)
    # convert_mtx.py:112 calls tobsr() with scipy's *estimated* block size first, so the
    # later tobsr((1,1)) keeps the explicit zeros of those larger blocks as 1x1 blocks
    # (small_10x10: 100 blocks for 90 non-zeros).  Dropping the zero blocks must give
    # exactly the blocks a direct conversion stores.
    keep = bsr.blocks != 0
    np.testing.assert_array_equal(mine.blocks, bsr.blocks[keep])
    np.testing.assert_array_equal(mine.blockColIdxs, bsr.blockColIdxs[keep])
    rows = np.repeat(np.arange(bsr.M), np.diff(bsr.blockRowPtrs.astype(np.int64)))[keep]
    np.testing.assert_array_equal(np.bincount(rows, minlength=bsr.M), np.diff(mine.blockRowPtrs.astype(np.int64)))
    ell = orc.read_colell(pick("_rowind.ell"), pick("_values_colmajor.ell"))
    mine = orc.csr_to_colell(csr)
    if mine.maxColNnz == ell.maxColNnz:
        np.testing.assert_array_equal(mine.rowIdxs, ell.rowIdxs)
        np.testing.assert_array_equal(mine.vals, ell.vals)
    rt = orc.colell_to_csr(ell)
    for f in ("rowPtrs", "colIdxs", "vals"):
        np.testing.assert_array_equal(getattr(rt, f), getattr(csr, f))


def test_writers_round_trip_through_readers(tmp_path):
    a = random_csr(37, 53, 0.15, seed=3, vals="wide")
    B = np.random.default_rng(4).uniform(-100, 100, size=(53, 9)).astype(np.float32)
    orc.write_csr(tmp_path / "m.csr", a)
    orc.write_coo(tmp_path / "m.coo", orc.csr_to_coo(a))
    orc.write_bsr(tmp_path / "m.bsr", orc.csr_to_bsr(a, 1, 1))
    ell = orc.csr_to_colell(a)
    orc.write_colell(tmp_path / "m_rowind.ell", tmp_path / "m_values_colmajor.ell", ell)
    orc.write_rowell(tmp_path / "m_colind.ell", tmp_path / "m_values.ell", a)
    orc.write_dense(tmp_path / "dense.in", B)
    b = orc.read_csr(str(tmp_path / "m.csr"))
    for f in ("rowPtrs", "colIdxs", "vals"):
        np.testing.assert_array_equal(getattr(b, f), getattr(a, f))
    np.testing.assert_array_equal(orc.read_dense(str(tmp_path / "dense.in")), B)
    e2 = orc.read_colell(str(tmp_path / "m_rowind.ell"), str(tmp_path / "m_values_colmajor.ell"))
    np.testing.assert_array_equal(e2.rowIdxs, ell.rowIdxs)
    np.testing.assert_array_equal(e2.vals, ell.vals)
    np.testing.assert_array_equal(orc.to_dense(orc.read_bsr(str(tmp_path / "m.bsr"))), orc.to_dense(a))


@pytest.mark.parametrize("M,K,N,d,seed", [(1, 1, 1, 1.0, 0), (64, 96, 33, 0.2, 1), (200, 150, 64, 0.05, 2),
                                          (33, 70, 5, 0.5, 3), (128, 128, 128, 0.0, 4)])
def test_oracle_bitwise_equals_reference_code_on_random_inputs(M, K, N, d, seed):
    if orc.ref_lib() is None:          # checked at run time: the session fixture may just have built it
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    a = random_csr(M, K, d, seed, vals="wide")
    B = np.random.default_rng(seed + 100).uniform(-100, 100, size=(K, N)).astype(np.float32)
    np.testing.assert_array_equal(orc.spmm_csr(a, B), orc.spmm_csr(a, B, use_ref=True))
    np.testing.assert_array_equal(orc.spmm_csr(a, B, omp=True), orc.spmm_csr(a, B, use_ref=True))
    coo = orc.csr_to_coo(a)
    np.testing.assert_array_equal(orc.spmm_coo(coo, B), orc.spmm_coo(coo, B, use_ref=True))
    ell = orc.csr_to_colell(a)
    if ell.maxColNnz:
        np.testing.assert_array_equal(orc.spmm_ell(ell, B), orc.spmm_ell(ell, B, use_ref=True))
    for bs in (1, 4):
        bsr = orc.csr_to_bsr(a, bs, bs)
        Bp = np.zeros((bsr.K, N), np.float32)
        Bp[:K] = B
        np.testing.assert_array_equal(orc.spmm_bsr(bsr, Bp), orc.spmm_bsr(bsr, Bp, use_ref=True))
        np.testing.assert_array_equal(orc.spmm_bsr(bsr, Bp, omp=True), orc.spmm_bsr(bsr, Bp))


def test_conversion_oracles_agree_with_scipy():
    sp = pytest.importorskip("scipy.sparse")
    a = random_csr(70, 90, 0.1, seed=9)
    S = sp.csr_matrix((a.vals, a.colIdxs.astype(np.int32), a.rowPtrs.astype(np.int32)), shape=(a.M, a.K))
    for bs in (2, 5, 10):
        b = orc.csr_to_bsr(a, bs, bs)
        T = S.tobsr((bs, bs))
        T.sort_indices()      # scipy leaves block columns in first-touch order; ours are sorted
        np.testing.assert_array_equal(b.blockRowPtrs, T.indptr)
        np.testing.assert_array_equal(b.blockColIdxs, T.indices)
        np.testing.assert_array_equal(b.blocks, T.data.ravel())
    csc = S.tocsc()
    e = orc.csr_to_colell(a)
    assert e.maxColNnz == np.diff(csc.indptr).max()
    s = orc.csr_to_sell(a, 32)
    np.testing.assert_array_equal(orc.to_dense(s), S.toarray())


def test_partition_rows_by_nnz_properties():
    a = random_csr(500, 300, 0.05, seed=11, skew=True)
    for parts in (1, 2, 3, 4, 8):
        s = orc.partition_rows_by_nnz(a.rowPtrs, parts)
        assert s[0] == 0 and s[-1] == a.M and np.all(np.diff(s.astype(np.int64)) >= 0)
        per = np.diff(a.rowPtrs[s].astype(np.int64))
        assert per.sum() == a.nnz
        longest = np.diff(a.rowPtrs.astype(np.int64)).max()
        assert per.max() <= a.nnz / parts + longest + 1
