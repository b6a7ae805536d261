"""Text summary of one .ncu-rep (details lines, pipe utilisation, DRAM bytes, stall reasons, hottest SASS lines, opcode mix):
    python tools/ncu_summary.py file.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top_n = sys.argv[2] if len(sys.argv) > 2 else "12"
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
keys = ("Duration", "Elapsed Cycles", "Executed Ipc Active", "Issue Slots Busy", "Registers Per", "Achieved Occ", "Executed Instructions  ",
        "Dynamic Shared Memory Per Block", "DRAM Throughput", "Memory Throughput")
for l in det.splitlines():
    if any(k in l for k in keys) or l.strip().startswith(("void ", "lavt::")):
        print(" ".join(l.split()))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
want = ["sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__sass_inst_executed_op_tmem_ldt.sum", "smsp__sass_inst_executed_op_tmem_stt.sum"]
for k, u, v in zip(rows[0], rows[1], rows[2]):
    if k in want:
        print(f"  {k} = {v} {u}")
print(subprocess.run([sys.executable, __file__.replace("ncu_summary", "ncu_src"), rep, top_n, "ops"], capture_output=True, text=True).stdout)
