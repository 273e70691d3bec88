"""Summarise an .ncu-rep (ncu --set full) into a small text table for profiles/."""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",           # non-sampled: MMA count x clocks per MMA x 4 sub-pipes
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "smsp__inst_executed.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__sass_inst_executed_op_global_ld.sum",
        "smsp__sass_inst_executed_op_global_st.sum"]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units, data = rows[0], rows[1], rows[2:]
idx = {c: i for i, c in enumerate(h)}
print("# %s : %d launches" % (rep.split("/")[-1], len(data)))
names = [r[idx["Kernel Name"]] for r in data]
for j, n in enumerate(names):
    print("# launch %d: %s grid=%s block=%s" % (j, n[:80], data[j][idx.get("Grid Size", 0)], data[j][idx.get("Block Size", 0)]))
for k in KEYS:
    cols = [c for c in h if c == k or c.endswith("." + k)]
    for c in cols[:1]:
        i = idx[c]
        print("%-80s %-10s %s" % (k, units[i], "  ".join(r[i] for r in data)))
# tensor-pipe utilisation from the non-sampled counter: (hmma cycles active / 4 sub-pipes) / elapsed cycles
try:
    hk = [c for c in h if c.endswith("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg")][0]
    ek = [c for c in h if c == "sm__cycles_elapsed.avg"][0]
    print("%-80s %-10s %s" % ("tensor pipe active = hmma_cycles_active_realtime / 4 / sm__cycles_elapsed", "%",
                               "  ".join("%.1f" % (100.0 * float(r[idx[hk]].replace(",", "")) / 4 / float(r[idx[ek]].replace(",", ""))) for r in data)))
except Exception:
    pass
