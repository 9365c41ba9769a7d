"""Extract dram__bytes_read.sum / dram__bytes_write.sum of the one kernel in an .ncu-rep into
profiles/traffic_recross.json (read by bench.py for roofline.traffic).
Usage: python profiles/ncu_traffic.py gpurun_out/x.ncu-rep child_steps child_pairs"""
import csv
import json
import os
import subprocess
import sys

rep, steps, pairs = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
M = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def b(key):
    v, u = M[key]
    return float(v) * scale[u]


out = dict(kernel=M["Kernel Name"][0], child_steps=steps, child_pairs=pairs, dram_bytes_read=b("dram__bytes_read.sum"),
           dram_bytes_write=b("dram__bytes_write.sum"), duration_under_ncu=float(M["gpu__time_duration.sum"][0]), duration_unit=M["gpu__time_duration.sum"][1],
           source=os.path.basename(rep))
# the hardware's view of the same launch, next to the census-based roofline fraction of bench.py
for key, name in (("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_active_pct"),
                  ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "dmma_pipe_active_pct"),
                  ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
                  ("smsp__inst_executed.sum", "warp_instructions"),
                  ("sass__inst_executed_local_loads", "local_loads"), ("sass__inst_executed_local_stores", "local_stores")):
    if key in M:
        out[name] = float(M[key][0])
# every pipe the capture reports (the DMMA of the tensor-core transform is not part of sm__pipe_fp64_cycles_active)
out["pipes"] = {k: M[k][0] + " " + M[k][1] for k in sorted(M)
                if k in ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
                         "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
                         "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
                         "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic_recross.json"), "w"), indent=1)
print(out)
