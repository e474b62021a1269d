"""GPU parity of the tensor-core BSR path (tcgen05 / TMEM / TMA bulk copies).

fp16/bf16 block variants are stated separately from the fp32 bar (BASELINE.json north_star):
  (a) against the oracle run on the ROUNDED operands (blocks and B rounded to bf16/fp16 exactly as the
      device conversion rounds them: products are then exact in fp32, only the fp32 accumulation order
      differs) the component-wise relative error must be <= 2e-5;
  (b) against the un-rounded fp32 oracle (the reference's spmmBSRCpu) it must be within the operand
      rounding bound 2*u + u^2 (u = 2^-9 bf16, 2^-12 fp16) plus (a)."""
import numpy as np
import pytest

from conftest import random_csr
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b():
    import torch
    assert torch.cuda.is_available()
    from __graft_entry__ import load_package
    pkg = load_package()
    pkg.lib()
    return pkg.binding


CASES = [
    # M, K, N, bs, density of the underlying CSR
    (64, 64, 128, 16, 0.05),
    (203, 170, 128, 16, 0.02),      # M, K not multiples of the block
    (512, 384, 512, 16, 0.01),
    (512, 384, 512, 32, 0.01),
    (300, 520, 200, 32, 0.02),      # N not a multiple of 128
    (1024, 1024, 1024, 16, 0.004),  # two 512-column tiles per block row
    (256, 256, 640, 32, 0.01),      # 512 + 128 columns
    (96, 96, 21, 16, 0.05),         # tiny N
    (160, 160, 128, 16, 0.0),       # no blocks at all: C must be zero
    # enough block rows for the PANEL kernel (union walk over the block columns of P block rows; P from the grid fit)
    (2048, 1024, 512, 16, 0.05),    # 128 block rows x 4 column tiles -> panels of 4
    (3000, 700, 200, 32, 0.30),     # 94 block rows, N not a multiple of 128, most block columns hit by both rows of a panel
    (11264, 256, 128, 16, 1.0),     # 704 block rows -> panels of 5, EVERY block column hit by all 5: more hits than a stage holds
    (5000, 640, 384, 32, 0.002),    # sparse block rows, some empty, last panel ragged
]


@pytest.mark.parametrize("dtype", ["bf16", "fp16"])
@pytest.mark.parametrize("M,K,N,bs,d", CASES)
def test_bsr_tensor_core(b, M, K, N, bs, d, dtype):
    a = random_csr(M, K, d, seed=M + N + bs)
    o = orc.csr_to_bsr(a, bs, bs)
    # dense blocks (as a pruned block-sparse weight would be): fill every stored block with values
    rng = np.random.default_rng(bs + N)
    o.blocks[:] = rng.uniform(-1, 1, o.blocks.shape).astype(np.float32)
    B = rng.uniform(-1, 1, (K, N)).astype(np.float32)
    Bp = np.zeros((o.K, N), np.float32)
    Bp[:K] = B
    rnd = orc.bf16_round if dtype == "bf16" else orc.fp16_round
    u = 2.0 ** -9 if dtype == "bf16" else 2.0 ** -12

    plan = b.BsrTcPlan(b.dev_u32(o.blockRowPtrs), b.dev_u32(o.blockColIdxs), b.dev_f32(o.blocks), o.M // bs, bs, K, N,
                       dtype=dtype)
    try:
        plan.prepare_B(b.dev_f32(B))
        got = plan.run().cpu().numpy()
        got2 = plan.run().cpu().numpy()
    finally:
        plan.close()
    assert got.shape == (o.M, N)
    np.testing.assert_array_equal(got, got2)                       # deterministic

    o_r = orc.BSR(o.M, o.K, bs, bs, o.blockRowPtrs, o.blockColIdxs, rnd(o.blocks))
    ref_rounded = orc.spmm_bsr(o_r, rnd(Bp), omp=True)
    dense = orc.csr_from_dense(np.abs(orc.to_dense(o_r)))
    denom = orc.absprod_csr(dense, np.abs(rnd(Bp)))
    err_a = orc.max_rel_err(got, ref_rounded, denom)
    assert err_a <= 2e-5, f"vs rounded-operand oracle: {err_a:.3e}"

    ref = orc.spmm_bsr(o, Bp, omp=True)                            # the reference's fp32 spmmBSRCpu
    dense32 = orc.csr_from_dense(np.abs(orc.to_dense(o)))
    denom32 = orc.absprod_csr(dense32, np.abs(Bp))
    err_b = orc.max_rel_err(got, ref, denom32)
    assert err_b <= 2 * u + u * u + 2e-5, f"vs fp32 oracle: {err_b:.3e}"
    if d == 0.0:
        assert not got.any()


def test_bsr_tc_rejects_other_block_sizes(b):
    a = random_csr(64, 64, 0.1, seed=1)
    o = orc.csr_to_bsr(a, 8, 8)
    with pytest.raises(b.CuspmmError):
        b.BsrTcPlan(b.dev_u32(o.blockRowPtrs), b.dev_u32(o.blockColIdxs), b.dev_f32(o.blocks), o.M // 8, 8, 64, 64)


def test_bsr_tc_integer_data_is_exact(b):
    """Small integers are exact in bf16 and fp32 accumulation of integers is exact: bit-for-bit."""
    a = random_csr(128, 128, 0.03, seed=5, vals="int")
    o = orc.csr_to_bsr(a, 16, 16)
    rng = np.random.default_rng(1)
    o.blocks[:] = rng.integers(-4, 5, o.blocks.shape).astype(np.float32)
    B = rng.integers(-4, 5, (128, 256)).astype(np.float32)
    plan = b.BsrTcPlan(b.dev_u32(o.blockRowPtrs), b.dev_u32(o.blockColIdxs), b.dev_f32(o.blocks), 8, 16, 128, 256)
    try:
        plan.prepare_B(b.dev_f32(B))
        got = plan.run().cpu().numpy()
    finally:
        plan.close()
    np.testing.assert_array_equal(got, orc.spmm_bsr(o, B))


def test_panel_kernel_through_the_hook():
    """The panel kernel (union walk over the block columns of P block rows) is not the default; CUSPMM_BSR_PANEL = 1 selects it
    (read once per process), so the parity cases above are re-run in a child process with the hook set."""
    import os
    import subprocess
    import sys
    if os.environ.get("CUSPMM_BSR_PANEL"):
        pytest.skip("already inside the child run")
    env = dict(os.environ, CUSPMM_BSR_PANEL="1")
    p = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-q", "-x", "-k",
                        "test_bsr_tensor_core or integer_data"], env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
