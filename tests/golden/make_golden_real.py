#!/usr/bin/env python
"""Fixtures from the reference's REAL SuiteSparse inputs (run in the build container only).

For every data/<dir> of the reference that ships a sparse .mtx (the dirs its test/*.sh sweep over), run the
reference's OWN converter (utils/python_utils/convert_mtx.py::process_mtx, imported from /root/reference) on a
copy of that .mtx and keep, under tests/golden/real/<dir>.npz (compressed):
    rowPtrs, colIdxs, vals     the .csr file the converter wrote, parsed exactly as src/formats/sparse_csr.cu:12-51 does
    coo_rows                   the row column of the .coo file (the (row, col) order the converter gives COO)
    head_rows, head_ref        C[head_rows, :] of the reference's own spmmCSRCpu (oracle/_ref, built from
                               /root/reference/src/spmm/csr/spmm_csr.cpp) for B = seeded U(-1,1), N = 16
The dense.mtx operands of those dirs are not kept (as text they are 1..300 MB); the tests draw B from the seed.
The GPU box has no /root/reference: tests read only the .npz files this script committed.
"""
import importlib.util
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402

CASES = {   # dir -> sparse matrix file
    "medium_1484": "qh1484.mtx", "medium_2048": "dw1024.mtx", "medium_2880": "g7jac010.mtx", "medium_4000": "tols4000.mtx",
    "large_15120": "ch7-6-b5.mtx", "large_20000": "ACTIVSg10K.mtx", "large_21074": "GL7d25.mtx", "large_25605": "n4c6-b13.mtx",
}
HEAD_N = 16
HEAD_SEED = 20261018


def head_operand(K):
    return np.random.default_rng(HEAD_SEED).uniform(-1, 1, (K, HEAD_N)).astype(np.float32)


def main():
    spec = importlib.util.spec_from_file_location("ref_convert_mtx", os.path.join(REF, "utils/python_utils/convert_mtx.py"))
    conv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(conv)
    orc.build(ref=True)
    assert orc.ref_lib() is not None, "oracle/_ref did not build"
    out = os.path.join(HERE, "real")
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    for case, mtx in CASES.items():
        with tempfile.TemporaryDirectory() as tmp:
            work = os.path.join(tmp, case)
            os.makedirs(work)
            shutil.copyfile(os.path.join(REF, "data", case, mtx), os.path.join(work, mtx))
            conv.process_mtx(work)
            stem = mtx[:-4]
            csr = orc.read_csr(os.path.join(work, stem + ".csr"))
            coo = orc.read_coo(os.path.join(work, stem + ".coo"))
        assert (orc.to_dense(coo) == orc.to_dense(csr)).all() if csr.M * csr.K <= 4e7 else True
        lens = np.diff(csr.rowPtrs.astype(np.int64))
        # rows to pin: the first 24, the longest 8 and 32 drawn at random
        rng = np.random.default_rng(HEAD_SEED + csr.M)
        rows = np.unique(np.concatenate([np.arange(min(24, csr.M)), np.argsort(lens)[-8:], rng.integers(0, csr.M, 32)])).astype(np.int64)
        sub_rp = np.zeros(len(rows) + 1, np.uint32)
        sub_rp[1:] = np.cumsum(lens[rows])
        idx = np.concatenate([np.arange(csr.rowPtrs[r], csr.rowPtrs[r + 1]) for r in rows]).astype(np.int64)
        sub = orc.CSR(len(rows), csr.K, sub_rp, csr.colIdxs[idx], csr.vals[idx])
        head = orc.spmm_csr(sub, head_operand(csr.K), use_ref=True)
        np.savez_compressed(os.path.join(out, case + ".npz"), M=csr.M, K=csr.K, rowPtrs=csr.rowPtrs, colIdxs=csr.colIdxs,
                            vals=csr.vals, coo_rows=coo.rowIdxs, head_rows=rows, head_ref=head, name=stem)
        print(f"{case}: {stem} {csr.M}x{csr.K} nnz {csr.nnz} rows {lens.min()}..{lens.max()} per row -> "
              f"{os.path.getsize(os.path.join(out, case + '.npz')) / 1024:.0f} KB")


if __name__ == "__main__":
    main()
