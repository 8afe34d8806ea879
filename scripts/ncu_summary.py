#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + source page) into text: key metrics, opcode mix per sample, stall reasons."""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]; nsamples = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    print("== kernel:", d.get("Kernel Name"), "grid", d.get("Grid Size"), "block", d.get("Block Size"))
    keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            # tensor-core side (tcgen05 kernels): operand reads through the shared-memory data pipe, tensor pipe activity
            "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active"]
    for k in keys:
        if k in d: print("  %-80s %s %s" % (k, d[k], u[k]))
    st = [(k, float(v)) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled") and v]
    for k, v in sorted(st, key=lambda x: -x[1])[:8]:
        print("  stall %-60s %.3f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
    if nsamples:
        print("  lane-instr/sample %.2f   smem wavefronts/sample %.3f" % (
            float(d["smsp__inst_executed.sum"]) * 32 / nsamples,
            float(d["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]) / nsamples))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = sum(int(r[iE]) for r in data); ts = sum(int(r[iN]) for r in data) or 1
ops, samp = collections.Counter(), collections.Counter()
for r in data:
    parts = r[iS].split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    op = op.split(".")[0]
    ops[op] += int(r[iE]); samp[op] += int(r[iN])
print("== opcode mix (dynamic)")
for op, c in ops.most_common(24):
    line = "  %-8s %5.1f%% instr  %5.1f%% stall-samples" % (op, 100 * c / tot, 100 * samp[op] / ts)
    if nsamples: line += "  %6.2f lane-instr/sample" % (c * 32 / nsamples)
    print(line)
print("== regions (400-instruction blocks in address order)")
for i in range(0, len(data), 400):
    blk = data[i:i + 400]
    e = sum(int(r[iE]) for r in blk); s = sum(int(r[iN]) for r in blk)
    if e * 200 > tot or s * 200 > ts:
        line = "  @%5d %5.1f%% instr %5.1f%% samples" % (i, 100 * e / tot, 100 * s / ts)
        if nsamples: line += " %6.2f lane-instr/sample" % (e * 32 / nsamples)
        print(line, " ", blk[0][iS].strip()[:50])
