"""Print the metrics of interest from `ncu -i X.ncu-rep --page raw --csv` (stdin), one launch per column."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg ',
        'smsp__inst_executed.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fp64', 'sm__pipe_alu_cycles_active.avg.pct', 'sm__pipe_fma_cycles_active.avg.pct',
        'sm__pipe_fmaheavy_cycles_active.avg.pct', 'sm__pipe_xu_cycles_active.avg.pct',
        'smsp__average_warps_issue_stalled', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__cycles_active.avg', 'smsp__cycles_active.avg', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum',
        'lts__t_sectors_srcunit_tex.sum', 'dram__throughput']
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
for i, h in enumerate(hdr):
    if any(h.startswith(w.strip()) for w in WANT) or 'tensor' in h or 'utc' in h.lower() or 'tmem' in h.lower():
        print(", ".join([h, units[i]] + [r[i] for r in rows[2:]]))
