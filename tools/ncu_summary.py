#!/usr/bin/env python
"""Concise text summary of one `ncu --set full` report (key counters, stall reasons, opcode mix).
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title line" > profiles/xxx.txt"""
import csv, io, subprocess, sys
rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h, u, v = r[0], r[1], r[2]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "sm__cycles_active.avg", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
        "smsp__sass_inst_executed_op_shared_st.sum", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum", "smsp__sass_inst_executed_op_global_ld.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
kn = v[h.index("Kernel Name")] if "Kernel Name" in h else "?"
print(title)
print(f"report: {rep}; kernel: {kn.replace(' ', '')}")
for i, n in enumerate(h):
    if n in keys or ("issue_stalled" in n and "per_issue_active" in n) or \
       ("sass_thread_inst_executed_op_f" in n and n.endswith("sum.per_cycle_elapsed")):
        print(f"{n} [{u[i]}] = {v[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
mix = subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_opmix.py")], input=src,
                     stdout=subprocess.PIPE, text=True).stdout
print("\ndynamic SASS opcode mix (warp instructions, share, share of stall samples):")
print(mix)
