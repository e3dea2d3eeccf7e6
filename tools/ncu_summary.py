"""Summarise an .ncu-rep (ncu --set full) into the small CSV kept under profiles/: per captured kernel launch the duration,
DRAM bytes, pipe and issue activity, occupancy, launch shape, stall reasons per issued instruction and FP64 instruction rates.
    python tools/ncu_summary.py gpurun_out/r2_k1_final.ncu-rep profiles/r2_k1_fused_ncu_full_bench_c5.csv "comment ..."
"""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block", "launch__waves_per_multiprocessor",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_src_fp64.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed"]
rep, out = sys.argv[1], sys.argv[2]
comment = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
with open(out, "w") as f:
    f.write("kernel,metric,unit,value\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        for i, h in enumerate(hdr):
            stall = h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
            if h in KEYS or (stall and float(r[i] or 0) >= 0.02):
                f.write(f"\"{name}\",{h},{units[i]},{r[i]}\n")
    if comment:
        f.write("# " + comment + "\n")
print(out)
