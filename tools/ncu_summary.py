"""Summarise an .ncu-rep (read on the CPU box): python tools/ncu_summary.py gpurun_out/x.ncu-rep [more patterns]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
extra = sys.argv[2:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
EXACT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
         "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
         "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
         "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
         "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
         "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
         "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
         "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg.per_cycle_active",
         "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum",
         "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
         "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
         "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_shared_ld.sum"]
for r in rows[2:]:
    print("=" * 100)
    stalls = []
    for i, h in enumerate(hdr):
        if h in EXACT or any(e in h for e in extra):
            print(f"{h:80s} {units[i]:14s} {r[i]}")
        m = __import__("re").match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", h)
        if m:
            try:
                stalls.append((float(r[i]), m.group(1)))
            except ValueError:
                pass
    print("stall (warps per issue-active cycle): " + ", ".join(f"{n}={v:.2f}" for v, n in sorted(stalls, reverse=True)[:8]))
