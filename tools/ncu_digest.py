"""Digest ncu CSV exports (raw + source pages) into a short text summary."""
import csv, re, sys
raw, src = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
EX = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
      "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
      "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
      "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
      "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
      "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
      "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
      "lts__t_sectors_op_red.sum", "lts__t_sectors.sum", "lts__t_sectors_srcunit_tex_op_red.sum", "sm__cycles_elapsed.avg.per_second"]
for r in rows[2:]:
    st = []
    for i, h in enumerate(hdr):
        if h in EX:
            print(f"{h:78s} {units[i]:14s} {r[i]}")
        m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", h)
        if m:
            try:
                st.append((float(r[i]), m.group(1)))
            except ValueError:
                pass
    print("stall (warps per issue-active cycle): " + ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
if src:
    rows = list(csv.reader(open(src)))
    hdr = rows[1]
    ia, isrc, iss, iinst = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    data, tot, ti, ops = [], 0, 0, {}
    for r in rows[2:]:
        try:
            s, n = int(r[iss]), int(r[iinst])
        except ValueError:
            continue
        tot += s
        ti += n
        data.append((s, r[ia][-5:], r[isrc][:84], n))
        op = r[isrc].split()[1] if r[isrc].startswith("@") else r[isrc].split()[0]
        ops[op] = ops.get(op, 0) + n
    print(f"stall samples {tot}, warp instructions {ti}")
    print("top opcodes: " + ", ".join(f"{k}={v / ti * 100:.1f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14]))
    for s, a, so, n in sorted(data, reverse=True)[:int(sys.argv[3]) if len(sys.argv) > 3 else 14]:
        print(f"{s:7d} {s / tot * 100:5.1f}% inst={n:9d} {a} {so}")
