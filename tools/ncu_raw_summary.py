"""Key metrics per kernel from `ncu -i X.ncu-rep --page raw --csv` (one block per profiled launch)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
units = dict(zip(hdr, rows[1]))
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("----")
    for w in want:
        if w in d:
            print("  %-68s %s %s" % (w, d[w], units.get(w, "") if w != "Kernel Name" else ""))
    st = {k.split("smsp__average_warps_issue_stalled_")[1].split("_per_issue_active")[0]: float(v) for k, v in d.items()
          if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v not in ("", "n/a")}
    print("  stall reasons (warps per issue-active cycle):", " ".join("%s=%.2f" % kv for kv in sorted(st.items(), key=lambda x: -x[1])[:8]))
