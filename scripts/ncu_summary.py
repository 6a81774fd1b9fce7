#!/usr/bin/env python
"""Text + JSON summary of one kernel of an .ncu-rep for profiles/ (the raw page filtered to the metric families
the DESIGN cites).   python scripts/ncu_summary.py REP "header line" UNITS OUT_PREFIX"""
import csv, io, json, subprocess, sys
rep, header, units, out = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
fam = ("dram__bytes", "gpu__time_duration", "l1tex__data_pipe_lsu_wavefronts_mem_shared", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared",
       "launch__", "sm__cycles_elapsed", "sm__inst_executed_pipe_alu.", "sm__inst_executed_pipe_fp64", "sm__inst_executed_pipe_lsu",
       "sm__inst_executed_pipe_tmem", "sm__inst_executed_pipe_uniform.", "sm__throughput", "sm__warps_active", "smsp__issue_active",
       "smsp__inst_executed.sum", "smsp__average_warp", "smsp__sass_inst_executed_op_tmem", "smsp__inst_executed_op_shared")
val = {}
lines = [header]
for i, n in enumerate(h):
    val[n] = v[i]
    if n.startswith(fam) and "pcsamp" not in n:
        lines.append(f"{n} [{u[i]}] = {v[i]}")
open(out + "_ncu_full_summary.txt", "w").write("\n".join(lines) + "\n")
f = lambda k: float(val[k].replace(",", ""))
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
ub = dict(zip(h, u))
rd = f("dram__bytes_read.sum") * scale[ub["dram__bytes_read.sum"]]
wr = f("dram__bytes_write.sum") * scale[ub["dram__bytes_write.sum"]]
json.dump({"kernel": val.get("Kernel Name", ""), "command": header, "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
           "smem_wavefronts_per_pivot": f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / units,
           "warp_instructions_per_pivot": f("smsp__inst_executed.sum") / units,
           "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "gpu_time_ms": f("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[ub["gpu__time_duration.sum"]]},
          open(out + "_ncu_summary.json", "w"), indent=1)
print(open(out + "_ncu_summary.json").read())
