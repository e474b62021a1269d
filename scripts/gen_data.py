#!/usr/bin/env python
"""Write a data directory in the reference's on-disk formats from a seeded random matrix.

Counterpart of the reference's utils/python_utils/gen_sparse.py:63-84 (scipy.sparse.random +
uniform values, A sp_{d}_{M}x{K}, B K x N) with the same file grammar as convert_mtx.py
(SURVEY.md appendix), but seeded and emitting every format the CLI reads:
  matrix.csr  matrix.coo  matrix.bsr  matrix_colind.ell  matrix_values.ell
  matrix_rowind.ell  matrix_values_colmajor.ell  dense.in
Usage:  python scripts/gen_data.py OUT_DIR --rows 2048 --cols 2048 --density 0.1 --N 1024 [--seed 618] [--bsr-block 1]
"""
import argparse
import os

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--rows", type=int, default=2048)
    ap.add_argument("--cols", type=int, default=2048)
    ap.add_argument("--density", type=float, default=0.1)
    ap.add_argument("--N", type=int, default=1024)
    ap.add_argument("--seed", type=int, default=618)
    ap.add_argument("--range", type=float, nargs=2, default=(-100.0, 100.0))     # gen_sparse.py:77,81
    ap.add_argument("--bsr-block", type=int, default=1)                           # convert_mtx.py:22 forces 1
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    M, K, N = a.rows, a.cols, a.N
    mask = rng.random((M, K)) < a.density
    rows, cols = np.nonzero(mask)                      # row-major => (row, col) sorted
    vals = rng.uniform(a.range[0], a.range[1], size=rows.size).astype(np.float32)
    B = rng.uniform(a.range[0], a.range[1], size=(K, N)).astype(np.float32)
    rp = np.zeros(M + 1, dtype=np.int64)
    rp[1:] = np.cumsum(np.bincount(rows, minlength=M))
    nnz = rows.size
    os.makedirs(a.out, exist_ok=True)
    j = lambda xs: " ".join(str(x) for x in xs)
    with open(os.path.join(a.out, "matrix.csr"), "w") as f:
        f.write(f"{M} {K} {nnz}\n{j(rp)}\n{j(cols)}\n{j(vals)}\n")
    with open(os.path.join(a.out, "matrix.coo"), "w") as f:
        f.write(f"{M} {K} {nnz}\n")
        f.writelines(f"{r} {c} {v}\n" for r, c, v in zip(rows, cols, vals))
    # row-ELL pair (required by the CLI, never loaded: main.cu:118-126) and column-ELL pair (loaded)
    lens = np.diff(rp)
    w = int(lens.max()) if M else 0
    with open(os.path.join(a.out, "matrix_colind.ell"), "w") as fc, open(os.path.join(a.out, "matrix_values.ell"), "w") as fv:
        fc.write(f"{M} {K} {nnz} {w}\n")
        for r in range(M):
            lo, hi = rp[r], rp[r + 1]
            fc.write(j(list(cols[lo:hi]) + [-1] * (w - (hi - lo))) + "\n")
            fv.write(j(list(vals[lo:hi]) + [0] * (w - (hi - lo))) + "\n")
    order = np.lexsort((rows, cols))
    cr, cc, cv = rows[order], cols[order], vals[order]
    cp = np.zeros(K + 1, dtype=np.int64)
    cp[1:] = np.cumsum(np.bincount(cc, minlength=K))
    wc = int(np.diff(cp).max()) if K else 0
    with open(os.path.join(a.out, "matrix_rowind.ell"), "w") as fr, open(os.path.join(a.out, "matrix_values_colmajor.ell"), "w") as fv:
        fr.write(f"{M} {K} {nnz} {wc}\n")
        for c in range(K):
            lo, hi = cp[c], cp[c + 1]
            fr.write(j(list(cr[lo:hi]) + [-1] * (wc - (hi - lo))) + "\n")
            fv.write(j(list(cv[lo:hi]) + [0] * (wc - (hi - lo))) + "\n")
    bs = a.bsr_block
    assert M % bs == 0 and K % bs == 0, "--bsr-block must divide rows and cols"
    key = (rows // bs) * (K // bs) + cols // bs
    uk, inv = np.unique(key, return_inverse=True)
    brp = np.zeros(M // bs + 1, dtype=np.int64)
    brp[1:] = np.cumsum(np.bincount(uk // (K // bs), minlength=M // bs))
    blocks = np.zeros((uk.size, bs * bs), dtype=np.float32)
    blocks[inv, (rows % bs) * bs + cols % bs] = vals
    with open(os.path.join(a.out, "matrix.bsr"), "w") as f:
        f.write(f"{M} {K} {blocks.size} {bs} {bs} {uk.size}\n{j(brp)}\n{j(uk % (K // bs))}\n")
        f.writelines(j(b) + "\n" for b in blocks)
    with open(os.path.join(a.out, "dense.in"), "w") as f:
        f.write(f"{K} {N} {int(np.count_nonzero(B))}\n")
        f.writelines(j(row) + "\n" for row in B)
    print(f"wrote {a.out}: A {M}x{K} nnz {nnz}, B {K}x{N}")


if __name__ == "__main__":
    main()
