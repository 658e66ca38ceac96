"""Condense an .ncu-rep (raw + source pages) into the handful of numbers the roofline argument needs.
usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/NAME.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__inst_executed.avg.per_cycle_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_static", "inst_executed", "thread_inst_executed",
    "smsp__sass_inst_executed_op_tma_ld.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print("kernel:", name)
    for i, h in enumerate(hdr):
        if h in WANT:
            print(f"  {h:75s} {r[i]:>18s} {units[i]}")
    print("  warp stall reasons (avg warps stalled per issue-active cycle):")
    st = [(float(r[i] or 0), h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    for v, h in sorted(st, reverse=True)[:8]:
        print(f"    {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v:8.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
data = rows[2:]
isrc, ie, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[ie]) for r in data) or 1
tots = sum(int(r[ismp]) for r in data) or 1
bars = [i for i, r in enumerate(data) if "BAR.SYNC" in r[isrc]]
print("  instruction / stall-sample share between barriers (SASS index ranges):")
prev = 0
for b in bars + [len(data)]:
    rs = data[prev:b]
    print(f"    sass[{prev:4d}:{b:4d}]  inst {sum(int(r[ie]) for r in rs) / tot * 100:5.1f}%   samples {sum(int(r[ismp]) for r in rs) / tots * 100:5.1f}%")
    prev = b
ops = {}
for r in data:
    op = r[isrc].split()[0] if r[isrc].split() and not r[isrc].split()[0].startswith("@") else (r[isrc].split()[1] if len(r[isrc].split()) > 1 else "?")
    op = op.split(".")[0]
    ops[op] = ops.get(op, 0) + int(r[ie])
print("  executed warp-instructions by opcode:")
for op, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14]:
    print(f"    {op:10s} {v / tot * 100:5.1f}%")
