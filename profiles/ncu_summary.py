"""Summarise an .ncu-rep (one kernel): headline metrics + per-source-file / per-line hot spots.
Usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep [nlines]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
nlines = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
M = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sass__inst_executed_shared_loads",
        "sass__inst_executed_shared_stores", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second"]
for k in keys:
    if k in M:
        print("%-70s %s %s" % (k, M[k][0], M[k][1]))
print("--- warp stall reasons (cycles per issued instruction) ---")
st = [(float(M[h][0]), h) for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
for v, h in sorted(st, reverse=True)[:8]:
    print("  %-28s %.3f" % (h.split("stalled_")[1].split("_per_issue")[0], v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur, h2 = None, None
agg = collections.defaultdict(lambda: [0, 0])
ft = collections.defaultdict(lambda: [0, 0])
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 2 and r[0] == "Line No":
        h2 = r
        ii, isamp = h2.index("Instructions Executed"), h2.index("# Samples")
        continue
    if h2 and len(r) > ii and r[0] != "":
        try:
            ln, inst, smp = int(r[0]), int(r[ii]), int(r[isamp])
        except ValueError:
            continue
        key = (cur, ln, r[1].strip()[:80])
        agg[key][0] += inst
        agg[key][1] += smp
        ft[cur][0] += inst
        ft[cur][1] += smp
tot = sum(v[0] for v in ft.values()) or 1
ts = sum(v[1] for v in ft.values()) or 1
print("--- instructions / stall samples per source file ---")
for f, v in sorted(ft.items(), key=lambda kv: -kv[1][0]):
    print("  %-22s inst %5.1f%%  samples %5.1f%%" % (f, 100 * v[0] / tot, 100 * v[1] / ts))
print("--- hottest source lines by stall samples ---")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:nlines]:
    print("  %-16s %4d inst %5.2f%% smp %5.2f%%  %s" % (k[0], k[1], 100 * v[0] / tot, 100 * v[1] / ts, k[2]))
