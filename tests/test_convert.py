"""CPU-only: the native converter host/cuspmm_convert reproduces the reference's offline converter
(utils/python_utils/convert_mtx.py) byte for byte on the fixtures (tests/golden/* were written by that
script, see tests/golden/make_golden.py), and its --bsr-block option stores real blocks."""
import filecmp
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle import oracle as orc

CONVERT = os.path.join(ROOT, "cuda-optimization-for-spmm_b200", "host", "cuspmm_convert")
SUFFIXES = (".csr", ".coo", ".bsr", "_colind.ell", "_values.ell", "_rowind.ell", "_values_colmajor.ell", "dense.in")


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(CONVERT):
        subprocess.run(["make", "-s", "-C", os.path.dirname(CONVERT)], check=True)
    assert os.path.exists(CONVERT)


@pytest.mark.parametrize("case", ["small_10x10", "small_32x32", "small_210"])
def test_converter_is_byte_identical_to_convert_mtx_py(case, tmp_path):
    src = os.path.join(GOLDEN, case)
    work = tmp_path / case
    work.mkdir()
    for name in os.listdir(src):
        if name.endswith(".mtx"):
            shutil.copy(os.path.join(src, name), work)
    subprocess.run([CONVERT, str(work)], check=True, capture_output=True)
    produced = sorted(n for n in os.listdir(work) if n.endswith(SUFFIXES))
    expected = sorted(n for n in os.listdir(src) if n.endswith(SUFFIXES))
    assert produced == expected
    for name in expected:
        assert filecmp.cmp(os.path.join(src, name), work / name, shallow=False), f"{case}/{name} differs"


def test_converter_python_float_formatting(tmp_path):
    """repr(float) corner cases: exponents, trailing .0, shortest round trip."""
    vals = [1.0, -0.5, 1e16, 1.5e16, 123456789012345680.0, 1e-4, 1e-5, 0.1 + 0.2, 2.5e-7, 1e22, 5e-324, 100.0, -215.0, 1 / 3]
    (tmp_path / "m.mtx").write_text("%%%%MatrixMarket matrix coordinate real general\n1 %d %d\n" % (len(vals), len(vals)) +
                                    "".join(f"1 {i + 1} {v!r}\n" for i, v in enumerate(vals)))
    subprocess.run([CONVERT, str(tmp_path)], check=True, capture_output=True)
    line = (tmp_path / "m.csr").read_text().split("\n")[3].split()
    assert line == [repr(v) for v in vals]


def test_converter_symmetric_and_real_blocks(tmp_path):
    (tmp_path / "s.mtx").write_text("%%MatrixMarket matrix coordinate integer symmetric\n4 4 4\n1 1 5\n2 1 7\n4 2 -3\n4 4 9\n")
    (tmp_path / "dense.mtx").write_text("%%MatrixMarket matrix coordinate pattern general\n4 2 3\n1 1\n3 2\n4 1\n")
    subprocess.run([CONVERT, str(tmp_path), "--bsr-block", "2"], check=True, capture_output=True)
    a = orc.read_csr(str(tmp_path / "s.csr"))
    D = orc.to_dense(a)
    expect = np.array([[5, 7, 0, 0], [7, 0, 0, -3], [0, 0, 0, 0], [0, -3, 0, 9]], np.float32)
    np.testing.assert_array_equal(D, expect)                       # off-diagonal entries mirrored
    b = orc.read_bsr(str(tmp_path / "s.bsr"))
    assert (b.br, b.bc) == (2, 2)
    np.testing.assert_array_equal(orc.to_dense(b), expect)
    mine = orc.csr_to_bsr(a, 2, 2)
    np.testing.assert_array_equal(b.blockColIdxs, mine.blockColIdxs)
    np.testing.assert_array_equal(b.blocks, mine.blocks)
    np.testing.assert_array_equal(orc.read_dense(str(tmp_path / "dense.in")), np.array([[1, 0], [0, 0], [0, 1], [1, 0]], np.float32))
    ell = orc.read_colell(str(tmp_path / "s_rowind.ell"), str(tmp_path / "s_values_colmajor.ell"))
    np.testing.assert_array_equal(orc.to_dense(ell), expect)
