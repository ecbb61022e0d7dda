#!/usr/bin/env python3
"""Summarise an Nsight Compute report for profiles/:   python tools/ncu_summary.py <file.ncu-rep> [title...]
Prints the launch facts, pipe/issue/stall metrics and DRAM traffic of the (first) captured kernel from `ncu --page raw`,
then -- when the report has the source page (`--import-source on`) -- the executed-instruction mix by opcode from
`ncu --page source`, with the share of the "wait" stall samples each opcode carries."""
import csv, io, re, subprocess, sys
from collections import Counter

rep = sys.argv[1]
title = " ".join(sys.argv[2:])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
KEEP = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_global_loads", "sass__inst_executed_global_stores",
        "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores"]
print("# " + (title or rep))
print("# ncu --set full --clock-control none --import-source on ; summarised by tools/ncu_summary.py")
for h, u, v in zip(hdr, units, vals):
    if h in KEEP or ("issue_stalled" in h and "per_issue_active" in h and float(v or 0) >= 0.02):
        print("%s [%s] = %s" % (h, u, v))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 3:
    rows = rows[1:]                    # first line names the kernel
    h = rows[0]
    def col(name):
        for i, x in enumerate(h):
            if x.strip() == name:
                return i
        return None
    ci, cs = col("Instructions Executed"), col("Source")
    cw = col("stall_wait")
    if ci is not None and cs is not None:
        cnt, wait = Counter(), Counter()
        for r in rows[1:]:
            if len(r) <= max(ci, cs):
                continue
            m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[cs])
            if not m:
                continue
            op = m.group(1)
            op = ".".join(op.split(".")[:3]) if op.startswith("IMAD") else op.split(".")[0] if not op.startswith(("LDG", "STG", "LDS", "STS", "SHFL", "IADD3")) else ".".join(op.split(".")[:2])
            try:
                cnt[op] += int(float(r[ci] or 0))
                if cw is not None and len(r) > cw:
                    wait[op] += int(float(r[cw] or 0))
            except ValueError:
                pass
        tot, wt = sum(cnt.values()), sum(wait.values()) or 1
        wide = sum(v for k, v in cnt.items() if k.startswith("IMAD.WIDE")) or 1
        print("\n# executed warp instructions by opcode: share | per 1000 IMAD.WIDE | share of 'wait' stall samples")
        for op, c in cnt.most_common(22):
            print("%-18s %6.2f%%  %8.1f  %6.2f%%" % (op, 100.0 * c / tot, 1000.0 * c / wide, 100.0 * wait[op] / wt))
        print("%-18s %6.2f%%  %8.1f" % ("(all)", 100.0, 1000.0 * tot / wide))
