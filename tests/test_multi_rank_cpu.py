"""N > 1 host logic on CPU: world size 2 over gloo (127.0.0.1).  Each rank owns the nnz-balanced row
panel the sharding module assigns it, computes it (with the CPU oracle standing in for the GPU kernel --
this is a test), and the gathered C and the reduced metrics must equal the single-process result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, random_csr


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from __graft_entry__ import load_package
        load_package()
        import importlib
        sh = importlib.import_module("cuspmm_b200.sharding")
        from oracle import oracle as orc
        from conftest import random_csr as rc
        a = rc(301, 200, 0.06, seed=42, skew=True)
        B = np.random.default_rng(43).uniform(-1, 1, (200, 24)).astype(np.float32)
        splits = sh.splits_by_nnz(a.rowPtrs, world)
        np.testing.assert_array_equal(splits, orc.partition_rows_by_nnz(a.rowPtrs, world))
        r0, r1 = sh.rank_panel(splits, rank)
        C_local = torch.from_numpy(orc.spmm_csr(a, B, rows=(r0, r1)).copy())
        C = sh.gather_panels(C_local, splits)
        local_nnz = int(a.rowPtrs[r1]) - int(a.rowPtrs[r0])
        gflops, ms = sh.job_throughput(local_ms_total=10.0 * (rank + 1), steps=5, local_flops_per_step=2.0 * local_nnz * 24)
        # end-to-end path: every rank uploads its 1/world row slice of B, an all-gather completes B (bench.py --gpus N)
        K = B.shape[0]
        ks = sh.b_slice_rows(K, world)
        k0, k1 = sh.local_b_slice(None, K, rank, world)
        mine = torch.zeros((ks, B.shape[1]), dtype=torch.float32)
        mine[:k1 - k0] = torch.from_numpy(B[k0:k1])
        full = sh.allgather_B(torch.empty((ks * world, B.shape[1]), dtype=torch.float32), mine)
        np.testing.assert_array_equal(full[:K].numpy(), B)
        assert not full[K:].any()
        np.save(os.path.join(out_dir, f"C_{rank}.npy"), C.numpy())
        np.save(os.path.join(out_dir, f"m_{rank}.npy"), np.array([gflops, ms, local_nnz]))
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from oracle import oracle as orc
    a = random_csr(301, 200, 0.06, seed=42, skew=True)
    B = np.random.default_rng(43).uniform(-1, 1, (200, 24)).astype(np.float32)
    ref = orc.spmm_csr(a, B)
    m = [np.load(tmp_path / f"m_{r}.npy") for r in range(world)]
    for r in range(world):
        np.testing.assert_array_equal(np.load(tmp_path / f"C_{r}.npy"), ref)     # every rank holds the full, exact C
    assert m[0][2] + m[1][2] == a.nnz
    # max over ranks of the time (rank 1: 20 ms / 5 steps), sum over ranks of the work
    assert m[0][1] == m[1][1] == pytest.approx(4.0)
    assert m[0][0] == m[1][0] == pytest.approx(2.0 * a.nnz * 24 / 4.0e-3 / 1e9)
    # nnz balance: no panel exceeds its share by more than the longest row
    longest = np.diff(a.rowPtrs.astype(np.int64)).max()
    assert max(m[0][2], m[1][2]) <= a.nnz / 2 + longest + 1


def test_splits_match_device_rule_properties():
    from __graft_entry__ import load_package
    load_package()
    import importlib
    sh = importlib.import_module("cuspmm_b200.sharding")
    from oracle import oracle as orc
    for seed in range(5):
        a = random_csr(257, 90, 0.1, seed=seed, skew=bool(seed % 2))
        for parts in (1, 2, 4, 8):
            s = sh.splits_by_nnz(a.rowPtrs, parts)
            np.testing.assert_array_equal(s, orc.partition_rows_by_nnz(a.rowPtrs, parts))
            assert s[0] == 0 and s[-1] == a.M and (np.diff(s) >= 0).all()
