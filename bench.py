#!/usr/bin/env python
"""bench.py -- SpMM GFLOP/s (2*nnz*N) + HBM GB/s (% roofline) vs cuSPARSE, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload NAME] [--format csr|coo|ell|bsr16|bsr32] [--variant V] [--no-extras]

A "step" is one C = A*B over the workload.  Default workload: large_25605 (25605^2, 90 % sparse, N = 512, CSR) -- the
configuration BASELINE.json's target is quoted on; it fits one GPU.

N > 1 (torchrun, one rank per GPU): STRONG scaling of that ONE matrix.  Every rank generates the same matrix from the same
seed, the row panels come from the device partitioner (cuspmm_partition_rows_by_nnz: nnz-balanced contiguous rows), a rank
keeps only its own panel of A and its rows of C; B is replicated.  No collective in the device-timed data path (rows of C are
independent: src/spmm/csr/spmm_csr.cpp:15-27).  `value` = the whole matrix's flops / the slowest rank's device time.

`e2e` = the same multiply from HOST buffers, copies inside the timed region, wall clock, max over ranks:
    N = 1: cuspmm_spmm_csr_host (pinned CSR + B in, C out, pipelined row panels);
    N > 1: every rank uploads ITS panel of A over its own PCIe link and ITS 1/N row slice of B, B is completed by an NCCL
           all-gather over NVLink, cuspmm_spmm_csr_host_devB multiplies (kernels wait for the all-gather) and returns the
           rank's rows of C to the host.
Every N also checks sampled rows of every rank's panel against the CPU oracle (`parity`).

Timing: CUDA events on the launching stream around each step, W >= 3 warm-up steps, barrier + synchronize on both sides, max over
ranks.  A + B + C = 630 MB > the 126 MB L2, so no flush is needed at N = 1 ("inputs_larger_than_l2"); whenever a rank's
operands are under 252 MB a 256 MB buffer is read between steps, outside the per-step events.

N = 1 only: `cpu_baseline` (the reference's own spmmCSRCpu from oracle/_ref on all host threads, one row block per thread, plus the
single-thread figure the reference would show), same-run cuSPARSE, and -- unless --no-extras -- the sub-records `formats` (COO, ELL,
BSR 16/32 on the headline shape), `configs` (the other BASELINE configs) and `convert` (device converters in GB/s).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "spmm_gflops"
UNIT = "GFLOP/s"
L2_BYTES = 126e6


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    out = {"hbm": 6650.0, "hbm_src": "fallback (B200_PROFILING.md)", "bf16": 1682.6, "bf16_src": "fallback"}
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            out["hbm"], out["hbm_src"] = float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
            out["bf16"], out["bf16_src"] = float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
        except Exception:
            pass
    return out


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe).
    nvidia-smi needs tens of ms before its first sample, so it is started before the warm-up and the
    samples are filtered by their timestamps to [mark_start(), mark_end()]."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def collect(rows):
            sm, mx, pw, reasons = [], [], [], set()
            for _, r in rows:
                try:
                    sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                    for name, v in zip(names, r[5:9]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
                except Exception:
                    pass
            return sm, mx, pw, reasons
        inside = [x for x in self.rows if self.t0 is not None and self.t0 - 0.02 <= x[0] <= (self.t1 or 1e18) + 0.03]
        sm, mx, pw, reasons = collect(inside if inside else self.rows)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm),
                "window": "timed region" if inside else "whole run (no sample fell inside the timed region)",
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------ CPU legs
def cpu_spmm_rows(orc, a_host, B_host, rows, threads):
    """The reference's own spmmCSRCpu (oracle/_ref; src/spmm/csr/spmm_csr.cpp:15-27 is single-threaded) on rows [0, rows)
    of a_host.  threads > 1: the rows are cut into `threads` nnz-balanced blocks and the SAME reference function runs on each
    block in its own host thread (ctypes releases the GIL), i.e. the reference's code on all host cores.  Without oracle/_ref
    the oracle port (OpenMP over rows) stands in.  -> (seconds, kind, threads used, nnz processed)."""
    import numpy as np
    rp = a_host.rowPtrs
    nnz = int(rp[rows])
    if orc.ref_lib() is None:
        sub = orc.CSR(rows, a_host.K, rp[:rows + 1], a_host.colIdxs[:nnz], a_host.vals[:nnz])
        t0 = time.perf_counter(); orc.spmm_csr(sub, B_host, omp=threads > 1); dt = time.perf_counter() - t0
        return dt, "port", (orc.lib().oracle_num_threads() if threads > 1 else 1), nnz
    threads = max(1, min(threads, rows))
    cuts = [int(np.searchsorted(rp[:rows + 1], (g * nnz) // threads, side="left")) for g in range(threads)] + [rows]
    cuts = sorted(set(cuts))
    blocks = []
    for r0, r1 in zip(cuts[:-1], cuts[1:]):
        i0, i1 = int(rp[r0]), int(rp[r1])
        blocks.append(orc.CSR(r1 - r0, a_host.K, (rp[r0:r1 + 1] - rp[r0]).astype(np.uint32), a_host.colIdxs[i0:i1], a_host.vals[i0:i1]))
    if len(blocks) == 1:
        t0 = time.perf_counter(); orc.spmm_csr(blocks[0], B_host, use_ref=True); dt = time.perf_counter() - t0
        return dt, "reference", 1, nnz
    ths = [threading.Thread(target=orc.spmm_csr, args=(blk, B_host), kwargs={"use_ref": True}) for blk in blocks]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return time.perf_counter() - t0, "reference", len(blocks), nnz


def calibrated_rows(run, target_s, start_rows, max_rows):
    """Rows such that run(rows) takes about target_s seconds (one probe run)."""
    dt = run(start_rows)[0]
    est = int(start_rows * target_s / max(dt, 1e-6))
    return max(start_rows, min(est, max_rows))


def numpy_csr_rows(np, K, density, rows, seed):
    rng = np.random.default_rng(seed)
    lens = rng.binomial(K, density, size=rows)
    rp = np.zeros(rows + 1, np.uint32); rp[1:] = np.cumsum(lens)
    ci = np.concatenate([np.sort(rng.choice(K, size=int(n), replace=False)) for n in lens]).astype(np.uint32)
    va = rng.uniform(-1, 1, size=int(rp[-1])).astype(np.float32)
    return rp, ci, va


def reference_arm(args, orc, np):
    """--impl reference: the reference's own CPU SpMM (oracle/_ref, else the oracle port) with all the host threads, each step a
    bounded row sample of the same workload.  Inputs are generated with numpy (same distribution, same density; no GPU)."""
    from importlib import util
    spec = util.spec_from_file_location("wl_named", os.path.join(ROOT, "cuda-optimization-for-spmm_b200", "workloads.py"))
    text = open(spec.origin).read()       # only the NAMED table is needed; avoid importing torch.cuda
    ns = {}
    exec(text[text.index("NAMED = {"):text.index("def gen_csr_device")], ns)
    M, K, density, N = ns["NAMED"][args.workload]
    threads = os.cpu_count() or 1
    rows_cap = min(M, 256 * threads)
    rp, ci, va = numpy_csr_rows(np, K, density, rows_cap, 618)
    B = np.random.default_rng(619).uniform(-1, 1, size=(K, N)).astype(np.float32)
    a = orc.CSR(rows_cap, K, rp, ci, va)
    run = lambda r: cpu_spmm_rows(orc, a, B, r, threads)
    total_steps = max(1, args.steps + args.warmup)
    per_step_s = max(0.5, min(5.0, 120.0 / total_steps))
    rows_s = calibrated_rows(run, per_step_s, min(rows_cap, 2 * threads), rows_cap)
    for _ in range(args.warmup):
        run(rows_s)
    t = 0.0
    kind, cores, snnz = "port", 1, 0
    for _ in range(args.steps):
        dt, kind, cores, snnz = run(rows_s)
        t += dt
    ms = t / args.steps * 1e3
    val = 2.0 * snnz * N / (ms * 1e-3) / 1e9
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"{args.workload}: {M}x{K} A, density {density}, B {K}x{N}, CSR; each step = rows "
                                  f"[0,{rows_s}) ({snnz} nnz) through the reference's spmmCSRCpu on {cores} host threads"},
           "cpu_baseline": {"value": val, "unit": UNIT, "kind": kind, "cores": cores,
                            "sample": f"{rows_s} rows ({snnz} nnz) per step, one nnz-balanced row block per thread",
                            "host_cpus": os.cpu_count()},
           "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)
    return 0


# ------------------------------------------------------------------------------------------ GPU helpers
class Ctx:
    pass


def timed_steps(torch, step, steps, warmup, flush=None):
    """-> per-step milliseconds (CUDA events on the current stream; L2 evicted before each step when `flush` is given)."""
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for e0, e1 in evs:
        if flush is not None:
            flush.sum()
        e0.record(); step(); e1.record()
    torch.cuda.synchronize()
    return [e0.elapsed_time(e1) for e0, e1 in evs]


def sampled_parity(cx, rp, ci, va, M, K, Bd, Cd, nrows=8, rounded=None):
    """max component-wise relative error |C - Cref| / (|A||B|) of `nrows` sampled rows of a device CSR product against the
    CPU oracle (oracle/spmm_oracle.c restating spmmCSRCpu)."""
    np, orc, wl = cx.np, cx.orc, cx.wl
    if M == 0:
        return 0.0
    rows = sorted(set(int(x) for x in np.linspace(0, M - 1, num=min(nrows, M))))
    B_host = Bd.cpu().numpy()
    worst = 0.0
    for r in rows:
        srp, sci, sva = wl.csr_sample_to_host(rp, ci, va, r, r + 1)
        a = orc.CSR(1, K, srp, sci, sva)
        ref = orc.spmm_csr(a, B_host)
        den = orc.absprod_csr(a, B_host)
        got = Cd[r:r + 1].cpu().numpy()
        worst = max(worst, float(orc.max_rel_err(got, ref, den)))
    return worst


def roofline_record(pk, alg_bytes, ms, flops, nnz, N, sm_mhz=None, traffic=None):
    achieved = alg_bytes / (ms * 1e-3) / 1e9
    clk = sm_mhz or 1965.0
    t_hbm = alg_bytes / (pk["hbm"] * 1e9) * 1e3
    t_fp32 = flops / (148 * 128 * 2 * clk * 1e6) * 1e3
    t_l1 = (4.0 * nnz * N) / (148 * 128 * clk * 1e6) * 1e3
    return {"bound": "hbm", "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s", "frac": achieved / pk["hbm"],
            "traffic": traffic, "peak_source": pk["hbm_src"], "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": ms,
            "bounds_ms": {"hbm": t_hbm, "fp32_fma": t_fp32, "smem_operand_bw": t_l1},
            "binding": max((("hbm", t_hbm), ("fp32_fma", t_fp32), ("smem_operand_bw", t_l1)), key=lambda x: x[1])[0]}


def tensor_roofline_record(pk, M, K, N, alg_bytes, ms, nnz, sm_mhz=None, traffic=None):
    """CSR variant 8 (spmm_csr_tc.cu): the kernel multiplies DENSE tiles, so its roofline is the tensor pipe.  Executed work per
    launch: Mpad x Kpad x Npad MACs in tf32 (half the bf16 rate) + the same MACs twice over in bf16 (the pair product has K
    doubled); in bf16-equivalent flops that is 8 x Mpad x Kpad x Npad.  peak = the measured dense bf16 rate."""
    Mp, Kp, Np = -(-M // 512) * 512, -(-K // 16) * 16, -(-N // 256) * 256
    macs = float(Mp) * Kp * Np
    achieved = 8.0 * macs / (ms * 1e-3) / 1e12
    clk = sm_mhz or 1965.0
    t_tensor = 8.0 * macs / (pk["bf16"] * 1e12) * 1e3
    t_hbm = alg_bytes / (pk["hbm"] * 1e9) * 1e3
    useful = 2.0 * nnz * N / (ms * 1e-3) / 1e12
    return {"bound": "tensor", "achieved": achieved, "peak": pk["bf16"], "unit": "TFLOP/s", "frac": achieved / pk["bf16"],
            "traffic": traffic, "peak_source": pk["bf16_src"], "kernel_ms": ms,
            "executed_macs_per_launch": macs, "padded_shape": [Mp, Kp, Np],
            "unit_note": "bf16-equivalent executed TFLOP/s: 2 flop x (tf32 MACs x 2 + bf16 MACs), zeros and padding included",
            "useful_tflops": useful, "useful_frac_of_fp32_fma_peak": useful / (148 * 128 * 2 * clk * 1e6 / 1e12),
            "hbm_view": {"algorithmic_bytes_per_launch": alg_bytes, "achieved_GBs": alg_bytes / (ms * 1e-3) / 1e9,
                         "frac_of_hbm_peak": alg_bytes / (ms * 1e-3) / 1e9 / pk["hbm"], "peak_source": pk["hbm_src"]},
            "bounds_ms": {"tensor": t_tensor, "hbm": t_hbm}, "binding": "tensor"}


def lookup_traffic(workload, fmt, kernel):
    """DRAM bytes of one launch from the committed ncu --set full capture -- only when that capture was taken on the SAME
    kernel the selector launches now (the file records the kernel per entry); otherwise null rather than a stale number."""
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tj.get(f"{workload}/{fmt}")
        if ent and ent.get("kernel") == kernel:
            return ent["bytes"]
    except Exception:
        pass
    return None


def bsr_case(cx, M, K, density, N, bs, seed, steps, dtype="bf16"):
    """BASELINE configs[3]: block-sparse A (density = fraction of bs x bs blocks stored), bf16 blocks on tcgen05."""
    torch, b, wl, pk = cx.torch, cx.b, cx.wl, cx.pk
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    nbr, nbc = (M + bs - 1) // bs, (K + bs - 1) // bs
    mask = torch.rand((nbr, nbc), generator=g, device="cuda") < density
    brp = torch.zeros(nbr + 1, dtype=torch.int64, device="cuda"); brp[1:] = torch.cumsum(mask.sum(dim=1, dtype=torch.int64), 0)
    bci = mask.nonzero(as_tuple=False)[:, 1].to(torch.int32)
    nb = int(bci.numel())
    blocks = torch.rand(nb * bs * bs, generator=g, device="cuda") * 2 - 1
    brp = brp.to(torch.int32)
    Mp, Kp = nbr * bs, nbc * bs
    Bd = wl.gen_dense_device(Kp, N, seed=619)
    Cd = torch.empty((Mp, N), dtype=torch.float32, device="cuda")
    nnz = nb * bs * bs
    flops = 2.0 * nnz * N
    t0 = time.perf_counter()
    plan = b.BsrTcPlan(brp, bci, blocks, nbr, bs, Kp, N, dtype=dtype)
    torch.cuda.synchronize()
    plan_ms = (time.perf_counter() - t0) * 1e3
    prep = timed_steps(torch, lambda: plan.prepare_B(Bd), 3, 1)
    alg_bytes = 2 * nnz + 4 * nb + 4 * (nbr + 1) + 2 * Kp * N + 4 * Mp * N
    flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda") if alg_bytes < 2 * L2_BYTES else None
    ts = timed_steps(torch, lambda: plan.run(out=Cd), steps, 3, flush)
    ms = statistics.median(ts)
    tf = flops / (ms * 1e-3) / 1e12
    rec = {"format": f"bsr{bs}", "kernel": "bsr_tc (tcgen05)", "dtype": f"{dtype} blocks and B, f32 accumulate", "M": Mp, "K": Kp, "N": N,
           "blocks": nb, "block_density": density, "ms": ms, "ms_min": min(ts), "gflops_executed": flops / (ms * 1e-3) / 1e9,
           "prepare_B_ms": statistics.median(prep), "plan_create_ms_wall": plan_ms,
           "roofline": {"bound": "tensor", "achieved": tf, "peak": pk["bf16"], "unit": "TFLOP/s", "frac": tf / pk["bf16"],
                        "peak_source": pk["bf16_src"], "algorithmic_bytes_per_launch": alg_bytes,
                        "hbm_GBs": alg_bytes / (ms * 1e-3) / 1e9, "hbm_frac": alg_bytes / (ms * 1e-3) / 1e9 / pk["hbm"]}}
    # parity: sampled block rows against the oracle on the ROUNDED operands (the stated tensor-core tolerance, 2e-5)
    try:
        np, orc = cx.np, cx.orc
        rnd = orc.bf16_round if dtype == "bf16" else orc.fp16_round
        Bh = rnd(Bd.cpu().numpy())
        worst = 0.0
        for R in sorted(set(int(x) for x in np.linspace(0, nbr - 1, num=4))):
            i0, i1 = int(brp[R].item()), int(brp[R + 1].item())
            sub = orc.BSR(bs, Kp, bs, bs, np.array([0, i1 - i0], np.uint32), bci[i0:i1].cpu().numpy().view(np.uint32).copy(),
                          rnd(blocks[i0 * bs * bs:i1 * bs * bs].cpu().numpy()))
            ref = orc.spmm_bsr(sub, Bh)
            den = orc.absprod_csr(orc.csr_from_dense(np.abs(orc.to_dense(sub))), np.abs(Bh))
            worst = max(worst, float(orc.max_rel_err(Cd[R * bs:(R + 1) * bs].cpu().numpy(), ref, den)))
        rec["parity"] = {"max_rel_err_vs_oracle_on_rounded_operands": worst, "tolerance": 2e-5, "ok": worst <= 2e-5,
                         "sample": "4 block rows"}
    except Exception as ex:
        rec["parity"] = {"error": str(ex)[:160]}
    try:
        tmp = torch.empty_like(Cd)
        avg, mn = b.cusparse_spmm_bsr(brp, bci, blocks, nbr, nbc, bs, Bd, tmp, warmup=1, iters=3)
        rec["cusparse"] = {"alg": "BSR fp32 ALG_DEFAULT", "ms_avg": avg, "speedup_vs_cusparse": avg / ms}
        try:
            avg, mn, w = b.cusparse_spmm_blockedell(brp, bci, blocks, nbr, nbc, bs, Bd, tmp, warmup=1, iters=3)
            rec["cusparse_blocked_ell"] = {"alg": "BLOCKED_ELL_ALG1 fp32", "ms_avg": avg, "speedup_vs_cusparse": avg / ms,
                                           "ell_width_blocks": w, "padding_factor": w * nbr / max(nb, 1)}
        except Exception as ex:
            rec["cusparse_blocked_ell"] = {"error": str(ex)[:160]}
        del tmp
    except Exception as ex:
        rec["cusparse"] = {"error": str(ex)[:160]}
    # end to end from host buffers: upload fp32 blocks + B, cast / re-tile on the device, multiply, download C
    try:
        h = [t.cpu().pin_memory() for t in (brp, bci, blocks, Bd)]
        C_h = torch.empty((Mp, N), dtype=torch.float32).pin_memory()
        variant = 2 if dtype == "bf16" else 3
        b.spmm_bsr_host(h[0], h[1], h[2], nbr, bs, bs, Kp, h[3], C_h, variant=variant)
        t0 = time.perf_counter()
        for _ in range(2):
            b.spmm_bsr_host(h[0], h[1], h[2], nbr, bs, bs, Kp, h[3], C_h, variant=variant)
        ems = (time.perf_counter() - t0) * 1e3 / 2
        rec["e2e"] = {"ms_per_step": ems, "value": flops / (ems * 1e-3) / 1e9, "unit": UNIT, "api": "cuspmm_spmm_bsr_host",
                      "h2d_bytes_per_step": 4 * (nbr + 1) + 4 * nb + 4 * nnz + 4 * Kp * N, "d2h_bytes_per_step": 4 * Mp * N,
                      "same_result": bool((C_h.cuda() == Cd).all().item())}
    except Exception as ex:
        rec["e2e"] = {"error": str(ex)[:200]}
    plan.close()
    return rec


def sparse_case(cx, name, M, K, density, N, fmts, steps, seed=618, pre=None):
    """One BASELINE shape, formats `fmts` of (csr, coo, ell): ms (median), GFLOP/s, kernel picked, x cuSPARSE, roofline with the
    in-run algorithmic bytes, sampled-row parity vs the oracle."""
    torch, b, wl, pk = cx.torch, cx.b, cx.wl, cx.pk
    rp, ci, va = pre if pre is not None else wl.gen_csr_device(M, K, density, seed=seed)
    Bd = wl.gen_dense_device(K, N, seed=619)
    Cd = torch.empty((M, N), dtype=torch.float32, device="cuda")
    nnz = int(ci.numel())
    flops = 2.0 * nnz * N
    out = {}
    small = wl.csr_bytes(M, K, N, nnz) < 2 * L2_BYTES
    flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda") if small else None
    cus = {}
    for fmt in fmts:
        rows = None
        if fmt == "csr":
            alg = wl.csr_bytes(M, K, N, nnz)
            step = lambda: b.spmm_csr(rp, ci, va, M, K, Bd, variant=0, out=Cd)
            kern = b.CSR_KERNEL_NAMES.get(b.csr_selected_variant(M, K, nnz, N))
        elif fmt == "coo":
            rows = torch.repeat_interleave(torch.arange(M, device="cuda", dtype=torch.int32), (rp[1:] - rp[:-1]).to(torch.int64))
            alg = wl.coo_bytes(M, K, N, nnz)
            step = lambda: b.spmm_coo(rows, ci, va, M, K, Bd, variant=0, out=Cd)
            kern = "coo->" + str(b.CSR_KERNEL_NAMES.get(b.csr_selected_variant(M, K, nnz, N)))
        else:
            sp, sc, sv = b.csr_to_sell(rp, ci, va, M)
            alg = wl.sell_bytes(M, K, N, int(sc.numel()), int(sp.numel()) - 1)
            step = lambda: b.spmm_sell(sp, sc, sv, M, K, Bd, variant=0, out=Cd)
            kern = "sell->" + str(b.CSR_KERNEL_NAMES.get(b.csr_selected_variant(M, K, int(sc.numel()), N, sell=True)))
        ts = timed_steps(torch, step, steps, 3, flush)
        ms = statistics.median(ts)
        rec = {"ms": ms, "ms_min": min(ts), "gflops": flops / (ms * 1e-3) / 1e9, "kernel": kern, "nnz": nnz,
               "roofline": roofline_record(pk, alg, ms, flops, nnz, N),
               "l2": "l2_flushed_between_steps" if small else "inputs_larger_than_l2"}
        err = sampled_parity(cx, rp, ci, va, M, K, Bd, Cd)
        rec["parity"] = {"max_rel_err_vs_oracle": err, "tolerance": 1e-5, "ok": err <= 1e-5, "sample": "8 rows"}
        which = 1 if fmt == "coo" else 0
        if which not in cus:
            try:
                tmp = torch.empty_like(Cd)
                rws = rows if which == 1 else rp
                if flush is None:
                    avg, mn = b.cusparse_spmm(which, rws, ci, va, M, K, Bd, tmp, warmup=1, iters=3)
                else:
                    b.cusparse_spmm(which, rws, ci, va, M, K, Bd, tmp, warmup=1, iters=1)
                    tt = []
                    for _ in range(5):
                        flush.sum()
                        tt.append(b.cusparse_spmm(which, rws, ci, va, M, K, Bd, tmp, warmup=0, iters=1)[0])
                    avg, mn = statistics.median(tt), min(tt)
                cus[which] = avg
                del tmp
            except Exception as ex:
                cus[which] = None
                rec["cusparse_error"] = str(ex)[:160]
        if cus.get(which):
            rec["cusparse"] = {"alg": "COO_ALG4" if which else "CSR_ALG2", "ms": cus[which], "speedup_vs_cusparse": cus[which] / ms}
            if fmt == "ell":
                rec["cusparse"]["note"] = "cuSPARSE has no sliced-ELL SpMM: compared with its CSR_ALG2 on the same matrix"
        if pre is not None and fmt in ("coo", "ell"):        # headline shape: the format's own end-to-end call (host operands)
            try:
                C_h = torch.empty((M, N), dtype=torch.float32).pin_memory()
                B_h = Bd.cpu().pin_memory()
                if fmt == "coo":
                    h = [t.cpu().pin_memory() for t in (rows, ci, va)]
                    call = lambda: b.spmm_coo_host(h[0], h[1], h[2], M, K, B_h, C_h)
                    h2d = 12 * nnz + 4 * K * N
                else:
                    h = [t.cpu().pin_memory() for t in (sp, sc, sv)]
                    call = lambda: b.spmm_sell_host(h[0], h[1], h[2], M, K, B_h, C_h)
                    h2d = 8 * int(sc.numel()) + 4 * int(sp.numel()) + 4 * K * N
                call()
                t0 = time.perf_counter()
                for _ in range(2):
                    call()
                ems = (time.perf_counter() - t0) * 1e3 / 2
                rec["e2e"] = {"ms_per_step": ems, "value": flops / (ems * 1e-3) / 1e9, "unit": UNIT,
                              "api": "cuspmm_spmm_coo_host" if fmt == "coo" else "cuspmm_spmm_sell_host",
                              "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * M * N,
                              "same_result": bool((C_h.cuda() == Cd).all().item()),
                              "max_abs_diff_vs_device_path": float((C_h.cuda() - Cd).abs().max().item())}
                del h, C_h, B_h
            except Exception as ex:
                rec["e2e"] = {"error": str(ex)[:200]}
        out[fmt] = rec
    return out


def convert_bench(cx, rp, ci, va, M, K):
    """The device converters as the HBM-bound kernels they are: GB/s of compulsory bytes against the measured HBM peak."""
    torch, b, pk = cx.torch, cx.b, cx.pk
    nnz = int(ci.numel())
    out = {}

    def rec(name, fn, bytes_):
        try:
            ts = timed_steps(torch, fn, 3, 1)
            ms = statistics.median(ts)
            out[name] = {"ms": ms, "algorithmic_bytes": bytes_, "GBs": bytes_ / (ms * 1e-3) / 1e9,
                         "frac_of_hbm_peak": bytes_ / (ms * 1e-3) / 1e9 / pk["hbm"]}
        except Exception as ex:
            out[name] = {"error": str(ex)[:160]}
    sp, sc, sv = b.csr_to_sell(rp, ci, va, M)
    slots = int(sc.numel())
    rec("csr_to_sell32", lambda: b.csr_to_sell(rp, ci, va, M), 4 * (M + 1) * 2 + 8 * nnz + 8 * slots)
    del sp, sc, sv
    rows = torch.repeat_interleave(torch.arange(M, device="cuda", dtype=torch.int32), (rp[1:] - rp[:-1]).to(torch.int64))
    rec("coo_to_csr_rowptrs", lambda: b.coo_to_csr_rowptrs(rows, M), 4 * (M + 1))
    rec("partition_rows_by_nnz_x8", lambda: b.partition_rows_by_nnz(rp, M, nnz, 8), 4 * (M + 1))
    rec("csr_check_sorted", lambda: b.csr_check_sorted(rp, ci, M, K), 4 * (M + 1) + 4 * nnz)
    del rows
    out["note"] = ("bytes = compulsory reads + writes of the conversion; coo_to_csr_rowptrs and the partitioner search instead of "
                   "streaming (M+1 binary / 32-ary searches), so their GB/s is not a bandwidth figure; CSR->BSR is measured on the "
                   "block-sparse configs by scripts/convert_bench.py (unstructured 10 % fills every 16x16 block: 2.6 GB of fp32 blocks)")
    return out


# ------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="large_25605")
    ap.add_argument("--format", default="csr", choices=["csr", "coo", "ell", "bsr16", "bsr32"])
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=8.0)
    ap.add_argument("--no-cusparse", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the formats / configs / convert sub-records (N = 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    import numpy as np
    from oracle import oracle as orc   # CPU-baseline / parity-check legs only

    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(args, orc, np)

    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    pkg = load_package()
    b = pkg.binding
    import importlib
    wl = importlib.import_module("cuspmm_b200.workloads")
    sh = importlib.import_module("cuspmm_b200.sharding")

    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a GPU: there is no CPU fallback"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = b.lib()
    cx = Ctx()
    cx.torch, cx.b, cx.wl, cx.np, cx.orc, cx.pk = torch, b, wl, np, orc, measured_peaks()
    pk = cx.pk

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    M, K, density, N = wl.NAMED[args.workload]
    fmt = args.format

    if fmt.startswith("bsr"):        # stand-alone tensor-core BSR line (one GPU; the default line carries it under "formats")
        assert world == 1, "--format bsr16/bsr32 is a single-GPU line"
        sampler = ClockSampler(local); sampler.start(); sampler.mark_start()
        rec = bsr_case(cx, M, K, density, N, int(fmt[3:]), 618, args.steps)
        sampler.mark_end()
        out = {"metric": METRIC, "value": rec["gflops_executed"], "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": 3,
               "ms_per_step": rec["ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": rec["dtype"],
               "data": "synthetic", "config": {"workload": f"{args.workload} block-sparse {fmt}", "format": fmt},
               "gpu_launches": args.steps, "roofline": rec["roofline"], "e2e": rec.get("e2e"), "cusparse": rec.get("cusparse"),
               "parity": rec.get("parity"), "prepare_B_ms": rec["prepare_B_ms"], "clocks": sampler.stop()}
        print(json.dumps(out), flush=True)
        return 0

    # ---- the ONE matrix (same seed on every rank) and this rank's panel of it
    rp, ci, va = wl.gen_csr_device(M, K, density, seed=618)
    nnz_total = int(ci.numel())
    Bd = wl.gen_dense_device(K, N, seed=619)          # replicated operand
    if world > 1:
        splits = b.partition_rows_by_nnz(rp, M, nnz_total, world)       # device partitioner, identical on every rank
        r0, r1 = int(splits[rank]), int(splits[rank + 1])
        i0, i1 = int(rp[r0].item()), int(rp[r1].item())
        rp_l = (rp[r0:r1 + 1] - rp[r0]).contiguous()
        ci_l, va_l = ci[i0:i1].clone(), va[i0:i1].clone()
        del rp, ci, va
        torch.cuda.empty_cache()
    else:
        splits = np.array([0, M], dtype=np.uint32)
        r0, r1, rp_l, ci_l, va_l = 0, M, rp, ci, va
    Ml = r1 - r0
    nnz = int(ci_l.numel())
    Cd = torch.empty((Ml, N), dtype=torch.float32, device="cuda")
    flops_l = 2.0 * nnz * N
    flops_total = 2.0 * nnz_total * N

    if fmt == "csr":
        alg_l = wl.csr_bytes(Ml, K, N, nnz)
        alg_total = wl.csr_bytes(M, K, N, nnz_total)
        step = lambda: b.spmm_csr(rp_l, ci_l, va_l, Ml, K, Bd, variant=args.variant, out=Cd)
        kvar = args.variant or b.csr_selected_variant(Ml, K, nnz, N)
        kernel_name = b.CSR_KERNEL_NAMES.get(kvar)
    elif fmt == "coo":
        rows_l = torch.repeat_interleave(torch.arange(Ml, device="cuda", dtype=torch.int32), (rp_l[1:] - rp_l[:-1]).to(torch.int64))
        alg_l, alg_total = wl.coo_bytes(Ml, K, N, nnz), wl.coo_bytes(M, K, N, nnz_total)
        step = lambda: b.spmm_coo(rows_l, ci_l, va_l, Ml, K, Bd, variant=args.variant, out=Cd)
        kernel_name = "coo variant %d" % args.variant
    else:
        sp, sc, sv = b.csr_to_sell(rp_l, ci_l, va_l, Ml)
        alg_l = wl.sell_bytes(Ml, K, N, int(sc.numel()), int(sp.numel()) - 1)
        alg_total = alg_l if world == 1 else None
        step = lambda: b.spmm_sell(sp, sc, sv, Ml, K, Bd, variant=args.variant, out=Cd)
        kernel_name = "sell variant %d" % args.variant

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()

    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    # timing rule: inputs larger than L2, or L2 evicted between timed steps (outside the per-step events)
    l2_flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda") if alg_l < 2 * L2_BYTES else None
    L.cuspmm_reset_launch_count()
    barrier()
    sampler.mark_start()
    t_wall0 = time.perf_counter()
    for e0, e1 in evs:
        if l2_flush is not None:
            l2_flush.sum()
        e0.record()
        step()
        e1.record()
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    sampler.mark_end()
    launches = int(sh.reduce_sum(float(L.cuspmm_launch_count()), device="cuda"))
    clocks = sampler.stop() if rank == 0 else None
    per_step = [e0.elapsed_time(e1) for e0, e1 in evs]
    # whole-job throughput: the whole matrix's flops / the slowest rank's device time
    ms_local = sum(per_step) / args.steps
    ms_per_step = sh.reduce_max(ms_local, device="cuda")
    value = flops_total / (ms_per_step * 1e-3) / 1e9
    rank_ms = [0.0] * world
    rank_nnz = [0] * world
    if world > 1:
        tt = torch.zeros(world, dtype=torch.float64, device="cuda"); tt[rank] = ms_local
        nn = torch.zeros(world, dtype=torch.float64, device="cuda"); nn[rank] = nnz
        dist.all_reduce(tt); dist.all_reduce(nn)
        rank_ms, rank_nnz = [float(x) for x in tt.tolist()], [int(x) for x in nn.tolist()]
    else:
        rank_ms, rank_nnz = [ms_local], [nnz]

    # ---- parity: sampled rows of THIS rank's panel against the CPU oracle, every N (worst over ranks)
    parity = None
    try:
        err = sampled_parity(cx, rp_l, ci_l, va_l, Ml, K, Bd, Cd) if fmt == "csr" else \
            sampled_parity(cx, rp_l, ci_l, va_l, Ml, K, Bd, Cd)
        worst = sh.reduce_max(err, device="cuda")
        # a checksum over all panels: sum(C) in fp64 against (1^T A) B, both reduced over the ranks
        colsum = torch.zeros(K, dtype=torch.float64, device="cuda")
        colsum.index_add_(0, ci_l.to(torch.int64), va_l.to(torch.float64))
        want = float(sh.reduce_sum(float((colsum @ Bd.to(torch.float64)).sum().item()), device="cuda"))
        got = float(sh.reduce_sum(float(Cd.to(torch.float64).sum().item()), device="cuda"))
        scale = float(sh.reduce_sum(float((colsum.abs() @ Bd.abs().to(torch.float64)).sum().item()), device="cuda"))
        parity = {"max_rel_err_vs_oracle": worst, "tolerance": 1e-5, "sample": "8 rows of every rank's panel",
                  "checksum_rel_diff": abs(got - want) / max(scale, 1e-30), "ok": bool(worst <= 1e-5 and abs(got - want) <= 1e-6 * scale)}
    except Exception as ex:
        parity = {"error": str(ex)[:200], "ok": False}

    # ---- e2e: host buffers through the C ABI, copies inside the timed region, wall clock, max over ranks
    e2e = None
    try:
        if fmt != "csr":
            raise RuntimeError("the default e2e leg is CSR; COO / ELL / BSR host entries are timed in the 'formats' sub-records")
        rp_h, ci_h, va_h = rp_l.cpu().pin_memory(), ci_l.cpu().pin_memory(), va_l.cpu().pin_memory()
        C_h = torch.empty((Ml, N), dtype=torch.float32).pin_memory()
        if world == 1:
            B_h = Bd.cpu().pin_memory()
            e2e_step = lambda: b.spmm_csr_host(rp_h, ci_h, va_h, Ml, K, B_h, C_h, variant=args.variant)
            api = "cuspmm_spmm_csr_host (pinned host CSR + B in, C out; wall clock around the call)"
            h2d = 4 * (M + 1) + 8 * nnz_total + 4 * K * N
        else:
            ks = sh.b_slice_rows(K, world)                  # B rows per rank (padded so that the slices are equal)
            B_full = torch.zeros((ks * world, N), dtype=torch.float32, device="cuda")
            B_slice = torch.empty((ks, N), dtype=torch.float32, device="cuda")
            Bs_h = torch.zeros((ks, N), dtype=torch.float32).pin_memory()
            k0, k1 = sh.local_b_slice(None, K, rank, world)
            if k1 > k0:
                Bs_h[:k1 - k0] = Bd[k0:k1].cpu()

            def e2e_step():
                B_slice.copy_(Bs_h, non_blocking=True)                      # this rank's slice of B over its own PCIe link
                sh.allgather_B(B_full, B_slice)                             # completed over NVLink
                return b.spmm_csr_host_devB(rp_h, ci_h, va_h, Ml, K, B_full[:K], C_h, variant=args.variant)
            api = ("per rank: H2D of its A panel + its 1/N slice of B, NCCL all_gather of B over NVLink, "
                   "cuspmm_spmm_csr_host_devB, D2H of its C rows; wall clock, max over ranks")
            h2d = 4 * (M + world) + 8 * nnz_total + 4 * ks * world * N
        e2e_step()                                                            # warm-up (allocations, NCCL channels)
        barrier()
        t0 = time.perf_counter()
        dev_ms = 0.0
        for _ in range(args.e2e_steps):
            dev_ms += e2e_step()
        barrier()
        e2e_ms = sh.reduce_max((time.perf_counter() - t0) * 1e3 / args.e2e_steps, device="cuda")
        # the fp32 kernels are bit-identical whatever the panel split; the tensor-core kernel (split products, red.add of partial
        # tiles) agrees to rounding: compare the host path's C with the device path's, normalised by |A||B| (the parity metric)
        C_e = C_h.cuda()
        same = bool((C_e == Cd).all().item())
        same = bool(sh.reduce_max(0.0 if same else 1.0, device="cuda") == 0.0)
        den = b.spmm_csr(rp_l, ci_l, va_l.abs(), Ml, K, Bd.abs(), variant=1).clamp_min(1e-30)
        diff = sh.reduce_max(float(((C_e - Cd).abs() / den).max().item()), device="cuda")
        del C_e, den
        e2e = {"value": flops_total / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e2e_ms,
               "device_ms_per_step_rank0": dev_ms / args.e2e_steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * M * N,
               "api": api, "same_result_as_device_path": same, "max_rel_diff_vs_device_path": diff,
               "result_agrees_with_device_path": bool(diff <= 2e-6)}
    except Exception as ex:      # report, never hide
        e2e = {"value": None, "unit": UNIT, "error": str(ex)[:300]}

    # ---- same-run cuSPARSE on this rank's panel (max over ranks)
    cusparse = None
    if not args.no_cusparse and fmt in ("csr", "coo"):
        try:
            tmp = torch.empty_like(Cd)
            which, rws = (0, rp_l) if fmt == "csr" else (1, rows_l)
            if l2_flush is None:
                avg, mn = b.cusparse_spmm(which, rws, ci_l, va_l, Ml, K, Bd, tmp, warmup=2, iters=5)
            else:     # same protocol as ours: L2 evicted before every timed launch
                b.cusparse_spmm(which, rws, ci_l, va_l, Ml, K, Bd, tmp, warmup=2, iters=1)
                ts = []
                for _ in range(7):
                    l2_flush.sum()
                    ts.append(b.cusparse_spmm(which, rws, ci_l, va_l, Ml, K, Bd, tmp, warmup=0, iters=1)[0])
                avg, mn = statistics.median(ts), min(ts)
            avg = sh.reduce_max(avg, device="cuda")
            cusparse = {"alg": "CSR_ALG2" if fmt == "csr" else "COO_ALG4", "ms_avg": avg, "ms_min_rank0": mn,
                        "gflops": flops_total / (avg * 1e-3) / 1e9, "speedup_vs_cusparse": avg / ms_per_step,
                        "max_abs_diff_vs_ours_rank0": float((tmp - Cd).abs().max().item()),
                        "note": "the same row panels through cusparseSpMM on every rank, max over ranks"}
            del tmp
        except Exception as ex:
            cusparse = {"error": str(ex)[:200]}

    # ---- the fp32 FMA kernels on the same panels (tensor mode off), when the selector chose the tensor-core kernel
    fp32_path = None
    if fmt == "csr" and kernel_name == "csr_tensor" and not args.variant:
        prev = b.set_csr_tensor_mode(0)
        try:
            kv = b.csr_selected_variant(Ml, K, nnz, N)
            tmp = torch.empty_like(Cd)
            run32 = lambda: b.spmm_csr(rp_l, ci_l, va_l, Ml, K, Bd, variant=0, out=tmp)
            run32(); run32()
            ts = []
            for _ in range(7):
                if l2_flush is not None:
                    l2_flush.sum()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); run32(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms32 = sh.reduce_max(statistics.median(ts), device="cuda")
            den = (Cd.abs() + tmp.abs()).clamp_min(1e-30)
            fp32_path = {"kernel_rank0": b.CSR_KERNEL_NAMES.get(kv), "ms": ms32, "gflops": flops_total / (ms32 * 1e-3) / 1e9,
                         "tensor_path_speedup": ms32 / ms_per_step,
                         "max_abs_diff_vs_tensor_path_rank0": float((tmp - Cd).abs().max().item()),
                         "note": "cuspmm_set_csr_tensor_mode(0): the selector restricted to the fp32 FMA kernels, same timing protocol"}
            del tmp, den
        except Exception as ex:
            fp32_path = {"error": str(ex)[:200]}
        finally:
            b.set_csr_tensor_mode(prev)

    out = None
    if rank == 0:
        sm_clock = (clocks or {}).get("sm_mhz")
        traffic = lookup_traffic(args.workload, fmt, kernel_name) if world == 1 else None
        if kernel_name == "csr_tensor":
            roof = tensor_roofline_record(pk, Ml, K, N, alg_l, ms_per_step, nnz, sm_clock, traffic)
            roof["note"] = ("per step of rank 0 = csr_tc_prepare_B + memset of C + csr_tc_kernel (+ the fallback kernel, which exits at "
                            "once); csr_tc_kernel is ~96 % of it (profiles/r02_ncu_launches_bench.csv).  A tiles are made dense in "
                            "shared memory and multiplied on the tensor cores with a tf32 + bf16 three-product split (fp32-grade "
                            "result, see parity); what bounds it is shared-memory bandwidth and the builders, see DESIGN.md section 4")
        else:
            roof = roofline_record(pk, alg_l, ms_per_step, flops_l, nnz, N, sm_clock, traffic)
            roof["note"] = ("per launch of rank 0's kernel (algorithmic bytes of ITS panel / its device time; peak = one GPU). fp32 CUDA-core "
                            "SpMM at this density is bound by SM-local operand bandwidth (one distinct B element per FMA), not HBM: "
                            "smem_operand_bw = all B reads through LDS at 128 B/clk/SM; see DESIGN.md")
        roof["frac_of_binding_bound"] = max(roof["bounds_ms"].values()) / ms_per_step
        roof["kernel"] = kernel_name
        imb = max(rank_nnz) * world / max(sum(rank_nnz), 1)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None,
            "dtype": "f32 (tensor cores: tf32 + bf16 three-product split, fp32 accumulate)" if kernel_name == "csr_tensor" else "f32",
            "data": "synthetic",
            "config": {"workload": f"{args.workload}: ONE {M}x{K} A, density {density} (nnz {nnz_total}), B {K}x{N}, "
                                   f"{fmt.upper()} variant {args.variant} (0 = selector -> {kernel_name})",
                       "format": fmt, "M": M, "K": K, "N": N, "nnz": nnz_total,
                       "parallelism": (f"{world} nnz-balanced contiguous row panels (cuspmm_partition_rows_by_nnz), one per GPU; "
                                       "B replicated; C left sharded by rows; no data-path collective") if world > 1 else "single GPU",
                       "row_splits": [int(x) for x in splits], "nnz_per_rank": rank_nnz, "nnz_imbalance": imb,
                       "ms_per_rank": rank_ms,
                       "l2": ("inputs_larger_than_l2 (A+B+C per rank = %.0f MB vs 126 MB L2; no flush)" % (alg_l / 1e6)) if l2_flush is None
                             else ("l2_flushed_between_steps (A+B+C per rank = %.0f MB; 256 MB read before every timed step, outside the events)" % (alg_l / 1e6)),
                       "seed": 618},
            "gpu_launches": launches,
            "wall_ms_timed_region": wall_ms,
            "roofline": roof,
            "e2e": e2e,
            "parity": parity,
            "cusparse": cusparse,
            "fp32_path": fp32_path,
            "clocks": clocks,
        }
    # ---- N = 1 only: CPU baseline and the other formats / configs (rank 0 alone exists)
    if world == 1 and rank == 0:
        try:
            threads = os.cpu_count() or 1
            rows_cap = min(M, 256 * threads)
            srp, sci, sva = wl.csr_sample_to_host(rp_l, ci_l, va_l, 0, rows_cap)
            a_host = orc.CSR(rows_cap, K, srp, sci, sva)
            B_host = Bd.cpu().numpy()
            run_all = lambda r: cpu_spmm_rows(orc, a_host, B_host, r, threads)
            rows_s = calibrated_rows(run_all, args.cpu_seconds, min(rows_cap, 2 * threads), rows_cap)
            dt, kind, cores, snnz = run_all(rows_s)
            out["cpu_baseline"] = {"value": 2.0 * snnz * N / dt / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
                                   "sample": f"rows [0,{rows_s}) of the workload ({snnz} nnz, {dt:.1f} s): the reference's spmmCSRCpu "
                                             f"(oracle/_ref, built -O2) on one nnz-balanced row block per host thread",
                                   "host_cpus": os.cpu_count()}
            run_one = lambda r: cpu_spmm_rows(orc, a_host, B_host, r, 1)
            rows_1 = calibrated_rows(run_one, args.cpu_seconds / 2, 8, rows_cap)
            dt, kind, cores, snnz = run_one(rows_1)
            out["cpu_baseline_single_thread"] = {"value": 2.0 * snnz * N / dt / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
                                                 "sample": f"rows [0,{rows_1}) ({snnz} nnz, {dt:.1f} s): as the reference runs it (one thread)"}
        except Exception as ex:
            out["cpu_baseline"] = {"value": None, "error": str(ex)[:200]}
        if not args.no_extras and args.workload == "large_25605" and fmt == "csr":
            t_ex = time.perf_counter()
            try:
                out["convert"] = convert_bench(cx, rp_l, ci_l, va_l, M, K)
            except Exception as ex:
                out["convert"] = {"error": str(ex)[:200]}
            formats = {}
            try:      # the headline shape in the other formats (same matrix), then the tensor-core BSR rows of configs[3]
                formats.update(sparse_case(cx, args.workload, M, K, density, N, ("coo", "ell"), 5, pre=(rp_l, ci_l, va_l)))
            except Exception as ex:
                formats["error_coo_ell"] = str(ex)[:200]
            del rp_l, ci_l, va_l
            torch.cuda.empty_cache()
            for bs in (16, 32):
                try:
                    formats[f"bsr{bs}"] = bsr_case(cx, M, K, density, N, bs, 618, 10)
                except Exception as ex:
                    formats[f"bsr{bs}"] = {"error": str(ex)[:200]}
            out["formats"] = formats
            configs = {}
            plan = [("medium_4096", ("csr", "coo", "ell"), None), ("medium_4000_s99", ("csr", "ell"), 2048),
                    ("medium_4000_s90", ("csr", "ell"), 128), ("medium_4000_s50", ("csr", "ell"), 2048),
                    ("large_20000", ("csr",), None), ("large_25605_s50", ("csr",), None),
                    ("ffn_11008x4096_s90", ("csr",), None), ("ffn_11008x4096_s50", ("csr",), None)]
            for name, fmts, n_over in plan:
                try:
                    m_, k_, d_, n_ = wl.NAMED[name]
                    n_ = n_over or n_
                    configs[f"{name}_N{n_}"] = sparse_case(cx, name, m_, k_, d_, n_, fmts, 5)
                    torch.cuda.empty_cache()
                except Exception as ex:
                    configs[name] = {"error": str(ex)[:200]}
            out["configs"] = configs
            out["extras_wall_s"] = time.perf_counter() - t_ex
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
