"""Condense an ncu report into the per-launch table committed under profiles/.
    ncu -i gpurun_out/X.ncu-rep --page raw --csv > /tmp/x.csv ; python profiles/ncu_extract.py /tmp/x.csv [row ...] > profiles/Y.csv
Keeps the metrics DESIGN.md / r1_summary.md quote (duration, DRAM bytes, pipe utilisation, issue, occupancy, stalls,
instruction and lane counts); one column per selected launch (default: all).
"""
import csv
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "sm__cycles_elapsed.avg.per_second",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")
STALL = "smsp__average_warps_issue_stalled_"

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
sel = [int(a) for a in sys.argv[2:]] or list(range(len(data)))
ix = {h: i for i, h in enumerate(hdr)}
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit"] + [f"launch {i}" for i in sel])
w.writerow(["Kernel Name", ""] + [data[i][ix["Kernel Name"]] for i in sel])
for h in hdr:
    if h in KEEP or (h.startswith(STALL) and h.endswith("_per_issue_active.ratio")):
        vals = [data[i][ix[h]] for i in sel]
        if any(v not in ("", "0", "n/a") for v in vals):
            w.writerow([h, units[ix[h]]] + vals)
