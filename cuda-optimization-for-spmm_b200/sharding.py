"""Host-side logic of the one-process-per-GPU (torchrun) front end: which row panel a rank owns,
how per-rank timings / work are combined, and how unequal C panels are gathered.

Rows of C are independent (reference: src/spmm/csr/spmm_csr.cpp:15-27), so there is NO collective in
the device-timed data path; the collectives are the timing reductions, the optional gather of C and -- in the end-to-end
(host operands) path -- the all-gather that completes B from the 1/world row slices every rank uploads over its own PCIe link.
Backend-agnostic (nccl on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def splits_by_nnz(rowPtrs: np.ndarray, parts: int) -> np.ndarray:
    """Same rule as the device partitioner cuspmm_partition_rows_by_nnz: s_g = first row r with
    rowPtrs[r] >= g*nnz/parts (integer division)."""
    M = rowPtrs.shape[0] - 1
    nnz = int(rowPtrs[-1])
    out = np.zeros(parts + 1, dtype=np.int64)
    for g in range(1, parts):
        out[g] = min(M, int(np.searchsorted(rowPtrs, (g * nnz) // parts, side="left")))
    out[parts] = M
    return np.maximum.accumulate(out)


def rank_panel(splits, rank: int):
    return int(splits[rank]), int(splits[rank + 1])


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def reduce_max(x: float, device="cpu") -> float:
    t = torch.tensor([x], dtype=torch.float64, device=device)
    if world() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(x: float, device="cpu") -> float:
    t = torch.tensor([x], dtype=torch.float64, device=device)
    if world() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def job_throughput(local_ms_total: float, steps: int, local_flops_per_step: float, device="cpu"):
    """Whole-job GFLOP/s: all ranks' work divided by the slowest rank's time."""
    ms = reduce_max(local_ms_total, device) / steps
    flops = reduce_sum(local_flops_per_step, device)
    return flops / (ms * 1e-3) / 1e9, ms


def gather_panels(C_local: torch.Tensor, splits) -> torch.Tensor:
    """All-gather of row panels of unequal height: pad to the tallest panel, all_gather, trim."""
    W = world()
    if W == 1:
        return C_local
    heights = [int(splits[g + 1] - splits[g]) for g in range(W)]
    hmax = max(heights)
    pad = torch.zeros((hmax, C_local.shape[1]), dtype=C_local.dtype, device=C_local.device)
    pad[:C_local.shape[0]] = C_local
    bufs = [torch.empty_like(pad) for _ in range(W)]
    dist.all_gather(bufs, pad)
    return torch.cat([bufs[g][:heights[g]] for g in range(W)], dim=0)


def b_slice_rows(K: int, world_size: int) -> int:
    """Rows of B every rank uploads in the end-to-end path: equal slices (all_gather needs them equal), the last one padded."""
    return (K + world_size - 1) // world_size


def local_b_slice(B_rows_of_rank, K: int, rank: int, world_size: int):
    """(k0, k1): the rows of B that rank owns; rows past K are padding."""
    ks = b_slice_rows(K, world_size)
    return rank * ks, min(K, (rank + 1) * ks)


def allgather_B(B_full: torch.Tensor, B_slice: torch.Tensor) -> torch.Tensor:
    """B_full[(world * ks) x N] <- the ranks' slices [ks x N] in rank order (NCCL: over NVLink; gloo in the CPU tests).
    Rows [0, K) of the result are B."""
    if world() == 1:
        B_full[:B_slice.shape[0]].copy_(B_slice)
        return B_full
    dist.all_gather_into_tensor(B_full, B_slice)
    return B_full
