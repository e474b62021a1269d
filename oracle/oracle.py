"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Python face of the CPU oracle:

* ctypes bindings to ``oracle/liboracle.so`` (``spmm_oracle.c``: restatement of the
  reference's ``spmm{CSR,COO,ELL,BSR}Cpu`` and ``toDense``) and, when present, to
  ``oracle/_ref/libref_spmm.so`` (the reference's own ``src/spmm/*/spmm_*.cpp``
  compiled from /root/reference by ``oracle/Makefile``);
* readers/writers for the reference's on-disk text formats (grammar: SURVEY.md
  appendix; readers restate ``src/formats/*.cu`` constructors, writers restate
  ``utils/python_utils/convert_mtx.py``);
* numpy restatements of the format conversions (``convert_mtx.py`` uses scipy
  ``tocsr/tocoo/tobsr``; the arrays below are defined to be exactly those), used to
  check the DEVICE converters bit for bit;
* the parity metrics.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module.  Nothing under ``cuda-optimization-for-spmm_b200/`` does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
U = C.c_uint32


def build(ref: bool = True) -> None:
    """Compile liboracle.so (always) and _ref/libref_spmm.so (when /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref and os.path.isdir("/root/reference/src/spmm"):
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = C.CDLL(path)
        L.oracle_num_threads.restype = C.c_int
        L.oracle_spmm_csr.argtypes = [U, U, U, u32p, u32p, f32p, f32p, f32p]
        for name in ("oracle_spmm_csr_rows", "oracle_spmm_csr_rows_omp"):
            getattr(L, name).argtypes = [U, U, U, u32p, u32p, f32p, f32p, f32p]
        L.oracle_spmm_coo.argtypes = [U, U, U, U, u32p, u32p, f32p, f32p, f32p]
        L.oracle_spmm_ell.argtypes = [U, U, U, U, u32p, f32p, f32p, f32p]
        for name in ("oracle_spmm_bsr", "oracle_spmm_bsr_omp"):
            getattr(L, name).argtypes = [U, U, U, U, u32p, u32p, f32p, f32p, f32p]
        L.oracle_csr_to_dense.argtypes = [U, U, u32p, u32p, f32p, f32p]
        L.oracle_coo_to_dense.argtypes = [U, U, U, u32p, u32p, f32p, f32p]
        L.oracle_ell_to_dense.argtypes = [U, U, U, u32p, f32p, f32p]
        L.oracle_bsr_to_dense.argtypes = [U, U, U, U, u32p, u32p, f32p, f32p]
        L.oracle_allclose.argtypes = [f32p, f32p, C.c_size_t, C.c_float, C.c_float]
        L.oracle_allclose.restype = C.c_int
        L.oracle_absprod_csr_rows.argtypes = [U, U, U, u32p, u32p, f32p, f32p, f64p]
        L.oracle_parse_f32.argtypes = [C.c_char_p, f32p, C.c_size_t]
        L.oracle_parse_f32.restype = C.c_size_t
        L.oracle_parse_u32.argtypes = [C.c_char_p, u32p, C.c_size_t]
        L.oracle_parse_u32.restype = C.c_size_t
        _LIB = L
    return _LIB


def ref_lib():
    """The reference's own compiled CPU SpMM, or None when it was never built."""
    global _REF
    if _REF is None:
        path = os.path.join(HERE, "_ref", "libref_spmm.so")
        if not os.path.exists(path):
            return None
        R = C.CDLL(path)
        R.ref_spmm_csr.argtypes = [U, U, U, U, u32p, u32p, f32p, f32p, f32p]
        R.ref_spmm_coo.argtypes = [U, U, U, U, u32p, u32p, f32p, f32p, f32p]
        R.ref_spmm_ell.argtypes = [U, U, U, U, U, u32p, f32p, f32p, f32p]
        R.ref_spmm_bsr.argtypes = [U, U, U, U, U, U, u32p, u32p, f32p, f32p, f32p]
        _REF = R
    return _REF


# --------------------------------------------------------------------- containers
@dataclass
class CSR:
    M: int
    K: int
    rowPtrs: np.ndarray
    colIdxs: np.ndarray
    vals: np.ndarray

    @property
    def nnz(self):
        return int(self.colIdxs.shape[0])


@dataclass
class COO:
    M: int
    K: int
    rowIdxs: np.ndarray
    colIdxs: np.ndarray
    vals: np.ndarray

    @property
    def nnz(self):
        return int(self.colIdxs.shape[0])


@dataclass
class ColELL:
    """The reference's column-ELL (include/formats/sparse_ell.hpp:12-37)."""
    M: int
    K: int
    nnz: int
    maxColNnz: int
    rowIdxs: np.ndarray  # [K * maxColNnz] uint32, 0xFFFFFFFF = padding
    vals: np.ndarray     # [K * maxColNnz] float32


@dataclass
class BSR:
    M: int
    K: int
    br: int
    bc: int
    blockRowPtrs: np.ndarray
    blockColIdxs: np.ndarray
    blocks: np.ndarray  # [numBlocks * br * bc] float32, row-major inside a block

    @property
    def numBlocks(self):
        return int(self.blockColIdxs.shape[0])


@dataclass
class SELL:
    """Sliced ELL as the device converter emits it ('Sliced ELL')."""
    M: int
    K: int
    sliceH: int
    slicePtrs: np.ndarray  # [numSlices + 1] uint32, in slots (element offsets / 1)
    colIdxs: np.ndarray    # [slicePtrs[-1]] uint32, 0xFFFFFFFF = padding
    vals: np.ndarray       # [slicePtrs[-1]] float32, 0 = padding


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


# --------------------------------------------------------------------- SpMM oracles
def spmm_csr(a: CSR, B: np.ndarray, rows=None, omp=False, use_ref=False) -> np.ndarray:
    B = _c(B, np.float32)
    N = B.shape[1]
    Cm = np.zeros((a.M, N), dtype=np.float32)
    if use_ref:
        ref_lib().ref_spmm_csr(a.M, a.K, N, a.nnz, a.rowPtrs, a.colIdxs, a.vals, B, Cm)
        return Cm
    r0, r1 = (0, a.M) if rows is None else rows
    fn = lib().oracle_spmm_csr_rows_omp if omp else lib().oracle_spmm_csr_rows
    fn(r0, r1, N, a.rowPtrs, a.colIdxs, a.vals, B, Cm)
    return Cm if rows is None else Cm[r0:r1]


def spmm_coo(a: COO, B: np.ndarray, use_ref=False) -> np.ndarray:
    B = _c(B, np.float32)
    N = B.shape[1]
    Cm = np.zeros((a.M, N), dtype=np.float32)
    if use_ref:
        ref_lib().ref_spmm_coo(a.M, a.K, N, a.nnz, a.rowIdxs, a.colIdxs, a.vals, B, Cm)
    else:
        lib().oracle_spmm_coo(a.M, a.K, N, a.nnz, a.rowIdxs, a.colIdxs, a.vals, B, Cm)
    return Cm


def spmm_ell(a: ColELL, B: np.ndarray, use_ref=False) -> np.ndarray:
    B = _c(B, np.float32)
    N = B.shape[1]
    Cm = np.zeros((a.M, N), dtype=np.float32)
    if use_ref:
        ref_lib().ref_spmm_ell(a.M, a.K, N, a.nnz, a.maxColNnz, a.rowIdxs, a.vals, B, Cm)
    else:
        lib().oracle_spmm_ell(a.M, a.K, N, a.maxColNnz, a.rowIdxs, a.vals, B, Cm)
    return Cm


def spmm_bsr(a: BSR, B: np.ndarray, omp=False, use_ref=False) -> np.ndarray:
    B = _c(B, np.float32)
    N = B.shape[1]
    Cm = np.zeros((a.M, N), dtype=np.float32)
    if use_ref:
        ref_lib().ref_spmm_bsr(a.M, a.K, N, a.br, a.bc, a.numBlocks, a.blockRowPtrs,
                               a.blockColIdxs, a.blocks, B, Cm)
    else:
        fn = lib().oracle_spmm_bsr_omp if omp else lib().oracle_spmm_bsr
        fn(a.M // a.br, N, a.br, a.bc, a.blockRowPtrs, a.blockColIdxs, a.blocks, B, Cm)
    return Cm


def to_dense(a) -> np.ndarray:
    D = np.zeros((a.M, a.K), dtype=np.float32)
    L = lib()
    if isinstance(a, CSR):
        L.oracle_csr_to_dense(a.M, a.K, a.rowPtrs, a.colIdxs, a.vals, D)
    elif isinstance(a, COO):
        L.oracle_coo_to_dense(a.M, a.K, a.nnz, a.rowIdxs, a.colIdxs, a.vals, D)
    elif isinstance(a, ColELL):
        L.oracle_ell_to_dense(a.M, a.K, a.maxColNnz, a.rowIdxs, a.vals, D)
    elif isinstance(a, BSR):
        L.oracle_bsr_to_dense(a.M, a.K, a.br, a.bc, a.blockRowPtrs, a.blockColIdxs, a.blocks, D)
    elif isinstance(a, SELL):
        H = a.sliceH
        for s in range(len(a.slicePtrs) - 1):
            lo, hi = int(a.slicePtrs[s]), int(a.slicePtrs[s + 1])
            w = (hi - lo) // H
            cols = a.colIdxs[lo:hi].reshape(w, H)
            vals = a.vals[lo:hi].reshape(w, H)
            for j in range(w):
                for i in range(H):
                    r = s * H + i
                    if r < a.M and cols[j, i] != 0xFFFFFFFF:
                        D[r, cols[j, i]] = vals[j, i]
    else:
        raise TypeError(type(a))
    return D


# --------------------------------------------------------------------- parity metrics
def allclose_ref(c: np.ndarray, ref: np.ndarray, rtol=1e-2, atol=1e-3) -> bool:
    """include/utils.hpp:10-11 + torch::allclose as the reference's wrappers use it."""
    c = _c(c, np.float32).ravel()
    ref = _c(ref, np.float32).ravel()
    return bool(lib().oracle_allclose(c, ref, c.size, rtol, atol))


def absprod_csr(a: CSR, B: np.ndarray, rows=None) -> np.ndarray:
    B = _c(B, np.float32)
    r0, r1 = (0, a.M) if rows is None else rows
    S = np.zeros((r1 - r0, B.shape[1]), dtype=np.float64)
    lib().oracle_absprod_csr_rows(r0, r1, B.shape[1], a.rowPtrs, a.colIdxs, a.vals, B, S)
    return S


def max_rel_err(c: np.ndarray, ref: np.ndarray, denom: np.ndarray) -> float:
    """North-star parity metric: max_ij |C - Cref|_ij / (|A|.|B|)_ij  (component-wise
    relative error; the denominator is the magnitude the sum passes through, so the
    figure is meaningful for elements that cancel to ~0).  Elements whose denominator
    is 0 must match exactly."""
    d = np.abs(c.astype(np.float64) - ref.astype(np.float64))
    z = denom == 0
    if np.any(d[z] != 0):
        return float("inf")
    with np.errstate(divide="ignore", invalid="ignore"):
        q = np.where(z, 0.0, d / denom)
    return float(q.max()) if q.size else 0.0


# --------------------------------------------------------------------- file readers
def _tokens_u32(text: str, n: int) -> np.ndarray:
    out = np.zeros(max(n, 1), dtype=np.uint32)
    got = lib().oracle_parse_u32(text.encode(), out, n)
    # the reference leaves unread slots at their memset-0 value (allocateSpace)
    return out[:n] if got >= 0 else out[:n]


def _tokens_f32(text: str, n: int) -> np.ndarray:
    out = np.zeros(max(n, 1), dtype=np.float32)
    lib().oracle_parse_f32(text.encode(), out, n)
    return out[:n]


def read_dense(path: str) -> np.ndarray:
    """src/formats/dense.cu:9-36: header 'rows cols [..]', then one text row per line."""
    with open(path) as f:
        head = f.readline().split()
        rows, cols = int(head[0]), int(head[1])
        out = np.zeros((rows, cols), dtype=np.float32)
        for i in range(rows):
            out[i] = _tokens_f32(f.readline(), cols)
    return out


def read_csr(path: str) -> CSR:
    """src/formats/sparse_csr.cu:12-51."""
    with open(path) as f:
        M, K, nnz = (int(t) for t in f.readline().split()[:3])
        rp = _tokens_u32(f.readline(), M + 1)
        ci = _tokens_u32(f.readline(), nnz)
        va = _tokens_f32(f.readline(), nnz)
    return CSR(M, K, rp, ci, va)


def read_coo(path: str) -> COO:
    """src/formats/sparse_coo.cu:13-38."""
    with open(path) as f:
        M, K, nnz = (int(t) for t in f.readline().split()[:3])
        r = np.zeros(nnz, np.uint32)
        c = np.zeros(nnz, np.uint32)
        v = np.zeros(nnz, np.float32)
        for i in range(nnz):
            t = f.readline().split()
            r[i] = _tokens_u32(t[0], 1)[0]
            c[i] = _tokens_u32(t[1], 1)[0]
            v[i] = _tokens_f32(t[2], 1)[0]
    return COO(M, K, r, c, v)


def read_colell(rowind_path: str, values_path: str) -> ColELL:
    """src/formats/sparse_ell.cu:13-55 (the *_rowind.ell / *_values_colmajor.ell pair)."""
    with open(rowind_path) as f:
        M, K, nnz, w = (int(t) for t in f.readline().split()[:4])
        ri = _tokens_u32(f.read(), K * w)
    with open(values_path) as f:
        va = _tokens_f32(f.read(), K * w)
    return ColELL(M, K, nnz, w, ri, va)


def read_bsr(path: str) -> BSR:
    """src/formats/sparse_bsr.cu:18-61."""
    with open(path) as f:
        M, K, _nnz, br, bc, nb = (int(t) for t in f.readline().split()[:6])
        nbr = M // br
        rp = _tokens_u32(f.readline(), nbr + 1)
        ci = _tokens_u32(f.readline(), nb)
        va = _tokens_f32(f.read(), nb * br * bc)
    return BSR(M, K, br, bc, rp, ci, va)


# --------------------------------------------------------------------- file writers
def _num(v) -> str:
    """convert_mtx.py writes values with Python str() of the numpy scalar."""
    return str(v)


def write_dense(path: str, D: np.ndarray) -> None:
    """convert_mtx.py:85-92 ('rows cols nnz' header, one row per line)."""
    with open(path, "w") as f:
        f.write(f"{D.shape[0]} {D.shape[1]} {int(np.count_nonzero(D))}\n")
        for row in D:
            f.write(" ".join(_num(x) for x in row) + "\n")


def write_csr(path: str, a: CSR) -> None:
    """convert_mtx.py:127-143."""
    with open(path, "w") as f:
        f.write(f"{a.M} {a.K} {a.nnz}\n")
        f.write(" ".join(map(str, a.rowPtrs)) + "\n")
        f.write(" ".join(map(str, a.colIdxs)) + "\n")
        f.write(" ".join(_num(x) for x in a.vals) + "\n")


def write_coo(path: str, a: COO) -> None:
    """convert_mtx.py:173-188."""
    with open(path, "w") as f:
        f.write(f"{a.M} {a.K} {a.nnz}\n")
        for r, c, v in zip(a.rowIdxs, a.colIdxs, a.vals):
            f.write(f"{r} {c} {_num(v)}\n")


def write_colell(rowind_path: str, values_path: str, a: ColELL) -> None:
    """convert_mtx.py:245-286."""
    ri = a.rowIdxs.astype(np.int64).reshape(a.K, a.maxColNnz)
    ri[ri == 0xFFFFFFFF] = -1
    va = a.vals.reshape(a.K, a.maxColNnz)
    with open(rowind_path, "w") as f:
        f.write(f"{a.M} {a.K} {a.nnz} {a.maxColNnz}\n")
        for row in ri:
            f.write(" ".join(map(str, row)) + "\n")
    with open(values_path, "w") as f:
        for row in va:
            f.write(" ".join(_num(x) for x in row) + "\n")


def write_rowell(colind_path: str, values_path: str, a: CSR) -> None:
    """convert_mtx.py:198-239: the row-ELL pair the CLI insists on but never loads."""
    lens = np.diff(a.rowPtrs.astype(np.int64))
    w = int(lens.max()) if a.M else 0
    with open(colind_path, "w") as fc, open(values_path, "w") as fv:
        fc.write(f"{a.M} {a.K} {a.nnz} {w}\n")
        for r in range(a.M):
            lo, hi = int(a.rowPtrs[r]), int(a.rowPtrs[r + 1])
            cols = [str(c) for c in a.colIdxs[lo:hi]] + ["-1"] * (w - (hi - lo))
            vals = [_num(v) for v in a.vals[lo:hi]] + ["0"] * (w - (hi - lo))
            fc.write(" ".join(cols) + "\n")
            fv.write(" ".join(vals) + "\n")


def write_bsr(path: str, a: BSR) -> None:
    """convert_mtx.py:46-59."""
    with open(path, "w") as f:
        f.write(f"{a.M} {a.K} {a.blocks.size} {a.br} {a.bc} {a.numBlocks}\n")
        f.write(" ".join(map(str, a.blockRowPtrs)) + "\n")
        f.write(" ".join(map(str, a.blockColIdxs)) + "\n")
        for blk in a.blocks.reshape(a.numBlocks, a.br * a.bc):
            f.write(" ".join(_num(x) for x in blk) + "\n")


# --------------------------------------------------------------------- conversions
def csr_from_dense(D: np.ndarray) -> CSR:
    D = np.asarray(D, dtype=np.float32)
    M, K = D.shape
    mask = D != 0
    rp = np.zeros(M + 1, dtype=np.uint32)
    rp[1:] = np.cumsum(mask.sum(axis=1))
    r, c = np.nonzero(mask)
    return CSR(M, K, rp, c.astype(np.uint32), D[r, c].astype(np.float32))


def csr_to_coo(a: CSR) -> COO:
    """scipy tocoo() + np.lexsort((cols, rows)) (convert_mtx.py:181-185) == CSR order
    when the CSR has sorted column indices."""
    rows = np.repeat(np.arange(a.M, dtype=np.uint32), np.diff(a.rowPtrs.astype(np.int64)))
    return COO(a.M, a.K, rows, a.colIdxs.copy(), a.vals.copy())


def coo_to_csr(a: COO) -> CSR:
    counts = np.bincount(a.rowIdxs, minlength=a.M).astype(np.uint32)
    rp = np.zeros(a.M + 1, dtype=np.uint32)
    rp[1:] = np.cumsum(counts)
    return CSR(a.M, a.K, rp, a.colIdxs.copy(), a.vals.copy())


def csr_to_colell(a: CSR) -> ColELL:
    """Column-ELL as the reference stores it.  Width = max nnz per COLUMN (the
    reference script takes getnnz(axis=1) of the CSC matrix, i.e. the max per ROW,
    convert_mtx.py:252, which under-allocates when a column is longer; the storage
    class itself only needs 'maxColNnz', sparse_ell.cu:32)."""
    M, K = a.M, a.K
    rows = np.repeat(np.arange(M, dtype=np.int64), np.diff(a.rowPtrs.astype(np.int64)))
    cols = a.colIdxs.astype(np.int64)
    order = np.lexsort((rows, cols))          # by column, then row  (== tocsc())
    rows_s, cols_s, vals_s = rows[order], cols[order], a.vals[order]
    counts = np.bincount(cols_s, minlength=K)
    w = int(counts.max()) if a.nnz else 0
    starts = np.zeros(K + 1, dtype=np.int64)
    starts[1:] = np.cumsum(counts)
    pos = np.arange(a.nnz, dtype=np.int64) - starts[cols_s]
    ri = np.full(K * max(w, 0), 0xFFFFFFFF, dtype=np.uint32)
    va = np.zeros(K * max(w, 0), dtype=np.float32)
    ri[cols_s * w + pos] = rows_s.astype(np.uint32)
    va[cols_s * w + pos] = vals_s
    return ColELL(M, K, a.nnz, w, ri, va)


def colell_to_csr(a: ColELL) -> CSR:
    ri = a.rowIdxs.reshape(a.K, a.maxColNnz)
    va = a.vals.reshape(a.K, a.maxColNnz)
    col, slot = np.nonzero(ri != 0xFFFFFFFF)
    rows = ri[col, slot].astype(np.int64)
    order = np.lexsort((col, rows))           # stable: by row, then column
    rows_s = rows[order]
    rp = np.zeros(a.M + 1, dtype=np.uint32)
    rp[1:] = np.cumsum(np.bincount(rows_s, minlength=a.M))
    return CSR(a.M, a.K, rp, col[order].astype(np.uint32), va[col, slot][order].astype(np.float32))


def csr_to_sell(a: CSR, H: int = 32) -> SELL:
    """Sliced ELL: slices of H consecutive rows; slice s is W_s = max row length in
    the slice wide and stored slot-major ('column-major'): entry j of row s*H+i sits
    at slicePtrs[s] + j*H + i.  Padding: col 0xFFFFFFFF, value 0.  Rows past M in the
    last slice are all padding."""
    M = a.M
    ns = (M + H - 1) // H
    lens = np.zeros(ns * H, dtype=np.int64)
    lens[:M] = np.diff(a.rowPtrs.astype(np.int64))
    W = lens.reshape(ns, H).max(axis=1) if ns else np.zeros(0, np.int64)
    sp = np.zeros(ns + 1, dtype=np.uint32)
    sp[1:] = np.cumsum(W * H)
    ci = np.full(int(sp[-1]), 0xFFFFFFFF, dtype=np.uint32)
    va = np.zeros(int(sp[-1]), dtype=np.float32)
    rows = np.repeat(np.arange(M, dtype=np.int64), lens[:M])
    j = np.arange(a.nnz, dtype=np.int64) - a.rowPtrs.astype(np.int64)[rows]
    dst = sp.astype(np.int64)[rows // H] + j * H + rows % H
    ci[dst] = a.colIdxs
    va[dst] = a.vals
    return SELL(M, a.K, H, sp, ci, va)


def csr_to_bsr(a: CSR, br: int, bc: int) -> BSR:
    """scipy ``tobsr((br, bc))`` semantics (what convert_mtx.py:24 calls): a block is
    stored iff it holds at least one stored entry; blocks are ordered by block row,
    then ascending block column (scipy leaves them in first-touch order until
    sort_indices(); any order is a valid input to spmmBSRCpu); values are row-major inside the block, zeros filled in.
    M and K are zero-padded up to multiples of br / bc first (scipy would refuse)."""
    Mp = (a.M + br - 1) // br * br
    Kp = (a.K + bc - 1) // bc * bc
    nbr, nbc = Mp // br, Kp // bc
    rows = np.repeat(np.arange(a.M, dtype=np.int64), np.diff(a.rowPtrs.astype(np.int64)))
    cols = a.colIdxs.astype(np.int64)
    key = (rows // br) * nbc + cols // bc
    ukeys, inv = np.unique(key, return_inverse=True)
    nb = ukeys.shape[0]
    rp = np.zeros(nbr + 1, dtype=np.uint32)
    rp[1:] = np.cumsum(np.bincount(ukeys // nbc, minlength=nbr))
    ci = (ukeys % nbc).astype(np.uint32)
    blocks = np.zeros(nb * br * bc, dtype=np.float32)
    blocks[inv * (br * bc) + (rows % br) * bc + cols % bc] = a.vals
    return BSR(Mp, Kp, br, bc, rp, ci, blocks)


def partition_rows_by_nnz(rowPtrs: np.ndarray, parts: int) -> np.ndarray:
    """Row-panel split points: s_0 = 0, s_parts = M and
    s_g = the first row r with rowPtrs[r] >= g * nnz / parts  (integer division of
    g*nnz by parts; numpy searchsorted 'left' == std::lower_bound)."""
    M = rowPtrs.shape[0] - 1
    nnz = int(rowPtrs[-1])
    out = np.zeros(parts + 1, dtype=np.uint32)
    for g in range(1, parts):
        t = (g * nnz) // parts
        out[g] = min(M, int(np.searchsorted(rowPtrs, t, side="left")))
    out[parts] = M
    return np.maximum.accumulate(out)


def bf16_round(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 (round to nearest even) -> fp32, as __float2bfloat16_rn does."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).reshape(np.shape(x))


def fp16_round(x: np.ndarray) -> np.ndarray:
    return np.asarray(x, dtype=np.float32).astype(np.float16).astype(np.float32)
