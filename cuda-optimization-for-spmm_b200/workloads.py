"""Seeded synthetic workloads of the BASELINE.json shapes, generated ON THE DEVICE (no
multi-GB text round trips).  Test / bench plumbing only.

Pattern: Bernoulli(density) per element (row nnz ~ Binomial), values U(-1, 1) fp32, column
indices ascending inside a row -- the same distribution family as the reference's
gen_sparse.py:63-84 (scipy.sparse.random + uniform values), seeded (default seed 618).
"""
from __future__ import annotations

import numpy as np
import torch

NAMED = {
    # name: (M, K, density, N)
    "small_synth": (512, 384, 0.10, 64),
    "medium_4096": (4096, 4096, 0.10, 512),          # BASELINE configs[1]
    "medium_4000_s99": (4000, 4000, 0.01, 512),      # configs[2] points
    "medium_4000_s90": (4000, 4000, 0.10, 512),
    "medium_4000_s50": (4000, 4000, 0.50, 512),
    "large_25605": (25605, 25605, 0.10, 512),        # north-star target row (configs[3] CSR/COO/ELL)
    "large_25605_s70": (25605, 25605, 0.30, 512),    # the same shape across the sparsity sweep of configs[2]
    "large_25605_s50": (25605, 25605, 0.50, 512),
    "large_20000": (20000, 20000, 0.10, 512),        # configs[4] row-sharded
    "ffn_11008x4096_s90": (11008, 4096, 0.10, 4096),
    "ffn_11008x4096_s50": (11008, 4096, 0.50, 4096),
}


def gen_csr_device(M, K, density, seed=618, device="cuda", rows_per_chunk=None):
    """-> (rowPtrs int32[M+1], colIdxs int32[nnz], vals float32[nnz]) on `device` (uint32 bits)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    if rows_per_chunk is None:
        rows_per_chunk = max(1, min(M, (1 << 26) // max(K, 1)))
    counts, cols = [], []
    for r0 in range(0, M, rows_per_chunk):
        r1 = min(M, r0 + rows_per_chunk)
        mask = torch.rand((r1 - r0, K), generator=g, device=device) < density
        counts.append(mask.sum(dim=1, dtype=torch.int64))
        cols.append(mask.nonzero(as_tuple=False)[:, 1].to(torch.int32))   # row-major => sorted
        del mask
    counts = torch.cat(counts) if counts else torch.zeros(0, dtype=torch.int64, device=device)
    colIdxs = torch.cat(cols) if cols else torch.zeros(0, dtype=torch.int32, device=device)
    rowPtrs = torch.zeros(M + 1, dtype=torch.int64, device=device)
    rowPtrs[1:] = torch.cumsum(counts, 0)
    nnz = int(rowPtrs[-1].item())
    assert nnz < 2 ** 31
    vals = torch.rand(nnz, generator=g, device=device, dtype=torch.float32) * 2.0 - 1.0
    return rowPtrs.to(torch.int32), colIdxs, vals


def gen_dense_device(K, N, seed=619, device="cuda"):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return torch.rand((K, N), generator=g, device=device, dtype=torch.float32) * 2.0 - 1.0


def csr_bytes(M, K, N, nnz):
    """Algorithmic (compulsory) bytes of one CSR SpMM: SURVEY.md section 8d."""
    return 8 * nnz + 4 * (M + 1) + 4 * K * N + 4 * M * N


def coo_bytes(M, K, N, nnz):
    return 12 * nnz + 4 * K * N + 4 * M * N


def sell_bytes(M, K, N, slots, slices):
    return 8 * slots + 4 * (slices + 1) + 4 * K * N + 4 * M * N


def csr_sample_to_host(rowPtrs, colIdxs, vals, r0, r1):
    """Rows [r0, r1) of a device CSR as a host oracle CSR-like tuple (rebased)."""
    rp = rowPtrs[r0:r1 + 1].to(torch.int64)
    i0, i1 = int(rp[0].item()), int(rp[-1].item())
    return ((rp - i0).cpu().numpy().astype(np.uint32), colIdxs[i0:i1].cpu().numpy().view(np.uint32).copy(),
            vals[i0:i1].cpu().numpy().copy())
