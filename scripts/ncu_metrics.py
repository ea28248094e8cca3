"""Print the roofline-relevant metrics of an `ncu --page raw --csv` dump (dev helper)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'l1tex__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__grid_size', 'sm__cycles_elapsed.avg.per_second', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'smsp__inst_executed.sum']
for r in rows[2:]:
    print('-----')
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w} = {r[i]} {units[i]}")
