#!/usr/bin/env python
"""Regenerate tests/golden/* from the reference tree (run in the build container only).

What it does, for data/small_10x10, data/small_32x32 and data/small_210:
  1. copies the directory out of the read-only /root/reference/data,
  2. runs the reference's OWN converter on the copy
     (utils/python_utils/convert_mtx.py::process_mtx, imported from /root/reference),
     producing .csr/.coo/.bsr/_colind.ell/_values.ell/_rowind.ell/_values_colmajor.ell/dense.in,
  3. keeps the reference's committed artefacts as they are
     (sparse.csr / Hamrle1.csr / *.coo / dense.in / result.expect / coo.out / coo_cuda.out)
     under ``committed/`` so tests can tell "what the reference ships" from
     "what its converter produced here",
  4. runs the reference's own compiled CPU SpMM (oracle/_ref/libref_spmm.so, built from
     /root/reference/src/spmm/*/spmm_*.cpp by oracle/Makefile) on the converted files and
     stores the four outputs as ``ref_{csr,coo,ell,bsr}.npy``.

The GPU box has no /root/reference; tests read only what this script committed.
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

CASES = ["small_10x10", "small_32x32", "small_210"]
COMMITTED = ["sparse.csr", "sparse.coo", "sparse.csc", "Hamrle1.csr", "Hamrle1.coo", "Hamrle1.csc",
             "dense.in", "result.expect", "coo.out", "coo_cuda.out"]


def main():
    spec = importlib.util.spec_from_file_location(
        "ref_convert_mtx", os.path.join(REF, "utils/python_utils/convert_mtx.py"))
    conv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(conv)
    orc.build(ref=True)
    assert orc.ref_lib() is not None, "oracle/_ref did not build"

    for case in CASES:
        src = os.path.join(REF, "data", case)
        out = os.path.join(HERE, case)
        shutil.rmtree(out, ignore_errors=True)
        os.makedirs(os.path.join(out, "committed"))
        for name in COMMITTED:
            p = os.path.join(src, name)
            if os.path.exists(p):
                shutil.copyfile(p, os.path.join(out, "committed", name))
        with tempfile.TemporaryDirectory() as tmp:
            work = os.path.join(tmp, case)
            os.makedirs(work)
            for name in os.listdir(src):
                if name.endswith(".mtx"):
                    shutil.copyfile(os.path.join(src, name), os.path.join(work, name))
            conv.process_mtx(work)
            for name in sorted(os.listdir(work)):
                # converter outputs, plus the .mtx inputs themselves (a few KB) so that the native
                # converter (host/cuspmm_convert) can be checked byte for byte without /root/reference
                shutil.copyfile(os.path.join(work, name), os.path.join(out, name))
        files = os.listdir(out)
        pick = lambda suf: os.path.join(out, [f for f in files if f.endswith(suf)][0])  # noqa: E731
        B = orc.read_dense(pick("dense.in"))
        csr = orc.read_csr(pick(".csr"))
        coo = orc.read_coo(pick(".coo"))
        bsr = orc.read_bsr(pick(".bsr"))
        np.save(os.path.join(out, "ref_csr.npy"), orc.spmm_csr(csr, B, use_ref=True))
        np.save(os.path.join(out, "ref_coo.npy"), orc.spmm_coo(coo, B, use_ref=True))
        np.save(os.path.join(out, "ref_bsr.npy"), orc.spmm_bsr(bsr, B, use_ref=True))
        rowind = [f for f in files if f.endswith("_rowind.ell")]
        if rowind:
            ell = orc.read_colell(pick("_rowind.ell"), pick("_values_colmajor.ell"))
            np.save(os.path.join(out, "ref_ell.npy"), orc.spmm_ell(ell, B, use_ref=True))
        print(case, "->", sorted(os.listdir(out)))


if __name__ == "__main__":
    main()
