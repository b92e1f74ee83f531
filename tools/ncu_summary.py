#!/usr/bin/env python
"""Summarise an ncu report (raw + source pages): key metrics, instruction counts and stall hot spots.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [min_samples]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; thr = int(sys.argv[2]) if len(sys.argv) > 2 else 300
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, units, r = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'smsp__inst_executed.sum',
        'launch__registers_per_thread', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.sum']
for i, h in enumerate(hdr):
    if h in keys: print(f"{h:75s} {r[i]:>16s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
tot = sum(int(x[idx['# Samples']]) for x in data); toti = sum(int(x[idx['Instructions Executed']]) for x in data)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print("total samples", tot, "warp instructions", toti)
cum = 0; cumi = 0
marks = ('UTCBAR', 'USETMAXREG', 'EXIT')
for i, x in enumerate(data):
    n = int(x[idx['# Samples']]); cum += n; cumi += int(x[idx['Instructions Executed']])
    s = x[idx['Source']].strip()
    if n >= thr or any(m in s for m in marks):
        st = sorted([(int(x[idx[h]]), h[6:]) for h in stalls], reverse=True)[:2]
        print(f"{i:5d} cum={100*cum/tot:5.1f}% inst={100*cumi/toti:5.1f}% n={n:6d} exec={x[idx['Instructions Executed']]:>9s}  {s[:64]:64s} {st}")
