# round 2, capture U (1 GPU): the bead transform from shared memory, FMA form against DMMA form (north_star: tensor cores for the
# transform only if ncu shows they beat the FMA path) -- timings, then ncu --set full of the four kernels
set -x
O=gpurun_out/r2u
mkdir -p $O
python profiles/bench_transform.py $O/bench_transform.json > $O/bench_transform.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:transform_bench -c 4 -f -o $O/prof_transform python profiles/bench_transform.py - ncu > $O/prof_transform.log 2>&1
ncu -i $O/prof_transform.ncu-rep --page raw --csv > $O/prof_transform_raw.csv 2> /dev/null
python - <<'PY' > $O/transform_ncu_summary.txt 2>&1
import csv
rows = list(csv.reader(open("gpurun_out/r2u/prof_transform_raw.csv")))
hdr = rows[0]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed_pipe_fp64.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"]
for r in rows[2:]:
    M = dict(zip(hdr, r))
    for k in keys:
        if k in M:
            print("%-70s %s" % (k, M[k]))
    st = sorted(((float(M[h]), h) for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and M[h] not in ("", "n/a")), reverse=True)[:6]
    for v, hh in st:
        print("   stall %-28s %.3f" % (hh.split("stalled_")[1].split("_per_issue")[0], v))
    print()
PY
rm -f $O/prof_transform.ncu-rep
ls -la $O
