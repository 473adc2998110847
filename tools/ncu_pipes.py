"""Selected pipe / memory counters of one kernel capture (ncu --set full), for the FP64-utilisation claims in DESIGN.md.
    python tools/ncu_pipes.py X.ncu-rep > profiles/..._pipes.txt"""
import csv, io, subprocess, sys
KEYS = ["sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "smsp__pipe_tensor_subpipe_dmma_cycles_active.avg",
        "sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "sm__inst_executed.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"]
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h, u = rows[0], rows[1]
for r in rows[2:]:
    d = {k.split("TriageCompute.")[-1]: (v, un) for k, un, v in zip(h, u, r)}
    print(d.get("Kernel Name", ("?",))[0])
    for k in KEYS:
        if k in d:
            print(f"  {k:90s} {d[k][0]:>16s} {d[k][1]}")
