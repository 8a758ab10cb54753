"""Summarise an .ncu-rep (--set full; one or more kernel launches) into a small text file for profiles/.
usage: summarize_ncu.py report.ncu-rep out.txt [launch index, default all]"""
import csv, subprocess, sys, io
rep, out = sys.argv[1], sys.argv[2]
only = int(sys.argv[3]) if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = [
    "Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
]
with open(out, "w") as f:
    f.write(f"# ncu --set full --clock-control none summary of {rep.split('/')[-1]}\n")
    for li, vals in enumerate(rows[2:]):
        if only is not None and li != only:
            continue
        d = dict(zip(hdr, zip(vals, units)))
        f.write(f"## launch {li}\n")
        for k in keys:
            if k in d:
                f.write(f"{k:75s} {d[k][0]} {d[k][1]}\n")
        st = []
        for k in hdr:
            if "smsp__average_warps_issue_stalled" in k and "per_issue_active" in k:
                try: st.append((float(d[k][0].replace(',', '')), k))
                except ValueError: pass
        f.write("# warp stall reasons (warps per issue-active cycle), top 8\n")
        for v, k in sorted(st, reverse=True)[:8]:
            f.write(f"{k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):30s} {v:.3f}\n")
print(open(out).read())
