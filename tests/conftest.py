import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # `-m gpu` selects them; without a device they are skipped rather than failed.
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_once():
    """A fresh checkout has no binaries (they are git-ignored): build them once per test session.
    (Test convenience only -- the product itself never builds or falls back at run time.)"""
    pkg = os.path.join(ROOT, "cuda-optimization-for-spmm_b200")
    need = [os.path.join(pkg, "libcuspmm_b200.so"), os.path.join(pkg, "host", "cuspmm"),
            os.path.join(pkg, "host", "cuspmm_convert"), os.path.join(ROOT, "oracle", "liboracle.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__ as g
        g.build()
    yield


def golden_case(name):
    d = os.path.join(GOLDEN, name)
    files = os.listdir(d)

    def pick(suf):
        m = [f for f in files if f.endswith(suf)]
        return os.path.join(d, m[0]) if m else None
    return d, pick


def random_csr(M, K, density, seed, vals="uniform", skew=False):
    """Seeded synthetic CSR with sorted column indices (numpy Generator PCG64)."""
    from oracle import oracle as orc
    rng = np.random.default_rng(seed)
    if skew:
        # a few very long rows, many short ones
        lens = np.minimum(K, (rng.pareto(1.2, size=M) * density * K).astype(np.int64))
    else:
        lens = rng.binomial(K, density, size=M).astype(np.int64)
    rp = np.zeros(M + 1, dtype=np.uint32)
    rp[1:] = np.cumsum(lens)
    cols = np.concatenate([np.sort(rng.choice(K, size=int(n), replace=False)) for n in lens]
                          or [np.zeros(0, np.int64)]).astype(np.uint32)
    nnz = int(rp[-1])
    if vals == "uniform":
        v = rng.uniform(-1, 1, size=nnz).astype(np.float32)
    elif vals == "int":
        v = rng.integers(-8, 9, size=nnz).astype(np.float32)
        v[v == 0] = 1
    else:
        v = rng.uniform(-100, 100, size=nnz).astype(np.float32)
    return orc.CSR(M, K, rp, cols, v)
