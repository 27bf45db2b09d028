"""Key metrics of one `ncu --set full` report: python tools/ncu_keys.py file.ncu-rep -> markdown rows."""
import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_issued.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "smsp__inst_executed_op_local_st.sum", "smsp__inst_executed_op_local_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg", "smsp__cycles_active.avg", "launch__grid_size", "launch__block_size"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")][:60]
    print(f"**{name}**  (grid {r[hdr.index('Grid Size')]}, block {r[hdr.index('Block Size')]})\n")
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"| `{k}` | {r[i]} | {units[i]} |")
    print()
