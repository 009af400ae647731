"""Key raw metrics per kernel of an ncu report.  python scripts/ncu_kernel_metrics.py report.ncu-rep"""
import csv, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_sectors_op_red.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__waves_per_multiprocessor', 'sm__cycles_active.avg', 'sm__cycles_elapsed.max']
for r in rows[2:]:
    print(r[h.index('Kernel Name')][:100])
    for w in want:
        if w in h:
            print("   %-62s %s %s" % (w, r[h.index(w)], rows[1][h.index(w)]))
