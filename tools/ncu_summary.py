#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU): key raw metrics + hottest CUDA source lines.
usage: tools/ncu_summary.py gpurun_out/prof_raster_X.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread',
 'smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__throughput.avg.pct_of_peak_sustained_active',
 'lts__throughput.avg.pct_of_peak_sustained_elapsed','smsp__thread_inst_executed_per_inst_executed.ratio','l1tex__t_sector_hit_rate.pct',
 'lts__t_sector_hit_rate.pct','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
 'sm__warps_active.avg.per_cycle_active','sm__inst_executed_pipe_lsu.sum','sm__inst_executed_pipe_alu.sum','sm__inst_executed_pipe_fma.sum',
 'sm__inst_executed_pipe_xu.sum','sm__inst_executed_pipe_uniform.sum','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active']
for k in keys:
    if k in hdr:
        i = hdr.index(k); print("%-70s %-12s %s" % (k, units[i], vals[i]))
stall = [(float(vals[i]), hdr[i]) for i in range(len(hdr)) if 'warp_issue_stalled' in hdr[i] and hdr[i].endswith('per_warp_active.pct') and vals[i]]
for v, k in sorted(stall, reverse=True)[:8]:
    print("stall %-64s %.2f" % (k.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''), v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(i for i, r in enumerate(rows[:6]) if 'Source' in r)
hdr = rows[h]; iS = hdr.index('Source'); iI = hdr.index('Instructions Executed'); iSm = hdr.index('# Samples')
agg = []
for r in rows[h + 1:]:
    if len(r) <= iI or not r[0].strip().isdigit(): continue      # keep CUDA source lines (have a line number)
    try: agg.append((int(r[iI]), int(r[iSm] or 0), r[0], r[iS].strip()[:100]))
    except ValueError: pass
tot = sum(a[0] for a in agg)
print("total inst attributed to CUDA lines:", tot)
for a in sorted(agg, reverse=True)[:topn]:
    print("%6.2f%% inst %5d smp  L%-4s %s" % (100.0 * a[0] / max(tot, 1), a[1], a[2], a[3]))
