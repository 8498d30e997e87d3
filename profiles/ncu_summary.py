#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + source page) into text: key metrics, stall reasons, opcode mix, hot source lines."""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum',
        'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed', 'sm__cycles_elapsed.max',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__shared_mem_per_block_dynamic']
print("== key metrics:", rep)
for i, h in enumerate(hdr):
    if h in keep: print(f"{h} = {vals[i]} {units[i]}")
print("== stall reasons (warps per issue-active cycle)")
for i, h in enumerate(hdr):
    if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
        print(f"  {h.split('stalled_')[1].split('_per_')[0]:24s} {float(vals[i]):.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(h2)}
ti = sum(int(r[ix["Instructions Executed"]]) for r in data); ts = sum(int(r[ix["# Samples"]]) for r in data)
opc = collections.Counter(); ops = collections.Counter()
for r in data:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ix["Source"]])
    op = m.group(2).split('.')[0] if m else r[ix["Source"]]
    opc[op] += int(r[ix["Instructions Executed"]]); ops[op] += int(r[ix["# Samples"]])
print(f"== opcode mix: {len(data)} SASS instrs, {ti} warp-instr executed")
for op, c in opc.most_common(22):
    print(f"  {op:10s} {c/ti*100:6.2f}% inst   {ops[op]/ts*100:6.2f}% stall-samples")
