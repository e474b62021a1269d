#!/usr/bin/env python
"""Text summary of one or more .ncu-rep files (ncu --set full ... --import-source on), for profiles/:
selected raw metrics per report side by side, then per report the opcode mix and the top stall sites from the SASS page.

    python scripts/ncu_summary.py gpurun_out/a.ncu-rep [gpurun_out/b.ncu-rep ...] > profiles/rNN_ncu_<what>_summary.txt
"""
import csv
import io
import subprocess
import sys
from collections import Counter

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared_op_utccp.sum", "smsp__sass_inst_executed_op_utccp.sum",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    reps = sys.argv[1:]
    cols = []
    for rep in reps:
        rows = ncu_csv(rep, "raw")
        hdr, units, vals = rows[0], rows[1], rows[2]
        cols.append({h: (v, u) for h, u, v in zip(hdr, units, vals)})
    print("# ncu --set full --clock-control none, one launch each; columns: " + " | ".join(r.split("/")[-1] for r in reps))
    for c in cols:
        print("# kernel: " + c.get("Kernel Name", ("?", ""))[0][:200])
    for m in METRICS:
        if not any(m in c for c in cols):
            continue
        line = f"{m:95s}"
        unit = ""
        for c in cols:
            v, unit = c.get(m, ("-", ""))
            line += f" {v:>18s}"
        print(line + "  " + unit)
    for rep in reps:
        rows = ncu_csv(rep, "source", ("--print-source", "sass"))
        if len(rows) < 3:
            continue
        hdr = rows[1]
        try:
            isrc, ie, iss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
        except ValueError:
            continue
        stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        data = [r for r in rows[2:] if len(r) > iss]
        tot = sum(int(r[ie] or 0) for r in data) or 1
        tots = sum(int(r[iss] or 0) for r in data) or 1
        ops, opst = Counter(), Counter()
        for r in data:
            s = r[isrc].split()
            op = (s[1] if s and s[0].startswith("@") and len(s) > 1 else (s[0] if s else "?")).split(".")[0]
            ops[op] += int(r[ie] or 0)
            opst[op] += int(r[iss] or 0)
        print(f"\n## {rep.split('/')[-1]}: opcode mix (share of executed warp instructions | share of stall samples)")
        for op, e in ops.most_common(18):
            print(f"  {op:12s} {100 * e / tot:5.1f} %   {100 * opst[op] / tots:5.1f} %")
        agg = {hdr[i][6:]: sum(int(r[i] or 0) for r in data) for i in stall_cols}
        print("## stall reasons (% of samples): " + ", ".join(f"{k} {100 * v / tots:.1f}" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
        print("## top stall sites")
        for r in sorted(data, key=lambda r: -int(r[iss] or 0))[:10]:
            st = sorted(((hdr[i][6:], int(r[i] or 0)) for i in stall_cols if int(r[i] or 0) > 0), key=lambda x: -x[1])[:2]
            print(f"  {100 * int(r[iss]) / tots:5.2f} %  {r[isrc][:70]:70s} {st}")


if __name__ == "__main__":
    main()
