"""Extracts the judged metrics of one kernel from `ncu -i rep --page raw --csv` into a small JSON.
usage: python tools/ncu_summary.py raw.csv out.json"""
import csv
import json
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg", "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
out = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")]}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            d[k] = "%s %s" % (r[i], units[i])
    out.append(d)
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out[0], indent=1))
