#!/usr/bin/env python
"""Summarise an .ncu-rep (read here with `ncu -i`): per kernel the duration, DRAM bytes,
L2 hit rate, pipe utilisation and top stall reasons.  python tools/ncu_summary.py rep [--stalls]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__warps_issue_stalled_long_scoreboard_per_warp_active.pct",
        "launch__shared_mem_per_block_dynamic", "derived__smsp__sass_thread_inst_executed_op_dfma_pred_on_x2"]
for d in data:
    print("==", d[idx["Kernel Name"]][:110])
    for w in want:
        if w in idx:
            print("   %-78s %12s %s" % (w, d[idx[w]][:14], units[idx[w]]))
    if "--stalls" in sys.argv:
        st = [(float(d[i]), h) for h, i in idx.items()
              if h.startswith("smsp__average_warp") and h.endswith("_per_issue_active.ratio") and d[i]]
        if not st:
            st = [(float(d[i]), h) for h, i in idx.items()
                  if "warps_issue_stalled" in h and h.endswith("per_warp_active.pct") and d[i]]
        for v, h in sorted(st, reverse=True)[:7]:
            print("   stall %-72s %12.3f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", ""), v))
