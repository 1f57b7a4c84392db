#!/usr/bin/env python3
"""Summarise an ncu report (read here, no GPU): key raw metrics + hot/cold split of the SASS of the first kernel.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep [launches.csv]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__grid_size", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "sm__inst_executed_pipe_fmaheavy.sum", "sm__inst_executed_pipe_fmalite.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum"]
print("kernel:", rows[2][h.index("Kernel Name")] if "Kernel Name" in h else "?")
for name in want:
    if name in h:
        i = h.index(name)
        print("%-95s %-12s %s" % (name, rows[1][i], " | ".join(r[i] for r in rows[2:])))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
ia, ie, it, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
data = []
for r in rows[hi + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    data.append(r)
tot = sum(int(r[ie]) for r in data); tots = sum(int(r[isamp]) for r in data)
ex = sorted(int(r[ie]) for r in data)
mx = ex[-3]
b, bs, bn = collections.Counter(), collections.Counter(), collections.Counter()
for r in data:
    e = int(r[ie])
    key = "A hot loops (>=0.5 max)" if e >= 0.5 * mx else ("B warm (0.05-0.5)" if e >= 0.05 * mx else ("C lukewarm (0.005-0.05)" if e >= 0.005 * mx else "D cold"))
    b[key] += e; bs[key] += int(r[isamp]); bn[key] += 1
print("\nSASS split: %d instructions, %d warp-instructions executed, %d samples" % (len(data), tot, tots))
for k in sorted(b):
    print("  %-26s n_instr %4d  executed %5.1f%%  samples %5.1f%%" % (k, bn[k], 100 * b[k] / tot, 100 * bs[k] / max(1, tots)))
ops = collections.Counter()
for r in data:
    if int(r[ie]) >= 0.5 * mx:
        op = r[ia].split()[0] if not r[ia].strip().startswith("@") else r[ia].split()[1]
        ops[op] += int(r[ie])
print("hot-loop opcode mix (share of hot warp-instructions):")
th = sum(ops.values())
print("  " + "  ".join("%s %.1f%%" % (k, 100 * v / th) for k, v in ops.most_common(14)))
if "--top" in sys.argv:
    print("\ntop sampled instructions:")
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:40]:
        print("  %-70s exec %5.2f%%max thr/inst %.1f samples %s" % (r[ia].strip()[:70], 100 * int(r[ie]) / mx, int(r[it]) / max(1, int(r[ie])), r[isamp]))
