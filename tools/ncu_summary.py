#!/usr/bin/env python
"""Condense `ncu --page raw --csv` dumps (profiles/*_ncu_raw_*.csv) into one markdown table."""
import csv
import glob
import os
import sys

WANT = [
    ('gpu__time_duration.sum', 'time'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM %'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'LTS %'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM %'),
    ('l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum', 'atom sectors'),
    ('l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum', 'red sectors'),
    ('lts__t_sectors.sum', 'LTS sectors'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occupancy %'),
    ('launch__registers_per_thread', 'regs'),
    ('smsp__inst_executed.sum', 'warp inst'),
    ('launch__grid_size', 'grid'),
]

# derived columns: sectors per request (32 = perfectly scattered, 4 = one sector per 8 lanes ...) and the share
# of every fetched sector's 32 bytes that the program actually used -- the "sector efficiency" north_star names
DERIVED = [
    ('ld sectors/req', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum'),
    ('st sectors/req', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum'),
    ('atom sectors/req', 'l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_atom.sum'),
    ('red sectors/req', 'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_red.sum'),
]
RATIO = [
    ('ld bytes used/sector', 'smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.ratio'),
    ('st bytes used/sector', 'smsp__sass_average_data_bytes_per_sector_mem_global_op_st.ratio'),
]


def main(pattern):
    labels = [label for _, label in WANT] + [d[0] for d in DERIVED] + [r[0] for r in RATIO]
    print('| kernel | ' + ' | '.join(labels) + ' |')
    print('|---|' + '---|' * len(labels))
    for path in sorted(glob.glob(pattern)):
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            name = vals[hdr.index('Kernel Name')].split('(')[0].replace('void ', '')
            cells = []
            for metric, _ in WANT:
                if metric in hdr:
                    i = hdr.index(metric)
                    try:
                        cells.append('{:.4g} {}'.format(float(vals[i].replace(',', '')), units[i]).strip())
                    except ValueError:
                        cells.append(vals[i])
                else:
                    cells.append('-')
            def num(metric):
                try:
                    return float(vals[hdr.index(metric)].replace(',', ''))
                except (ValueError, IndexError):
                    return None
            for _, sectors, requests in DERIVED:
                a, b = num(sectors), num(requests)
                cells.append('{:.2f}'.format(a / b) if a is not None and b else '-')
            for _, metric in RATIO:
                a = num(metric)
                cells.append('{:.1f} B'.format(a) if a is not None else '-')
            print('| `{}` ({}) | '.format(name, os.path.basename(path)) + ' | '.join(cells) + ' |')


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'profiles/*_ncu_raw_*.csv')
