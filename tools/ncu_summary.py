#!/usr/bin/env python
"""Summarise an `ncu --set full` raw CSV page (ncu -i X.ncu-rep --page raw --csv) into the handful
of numbers the roofline discussion needs.  usage: ncu_summary.py raw.csv [row_index]"""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'launch__grid_size', 'launch__block_size',
        'sm__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__inst_executed_pipe_fmaheavy.sum', 'sm__inst_executed_pipe_fmalite.sum', 'sm__inst_executed_pipe_fp64.sum',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active',
        'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_lgds.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_per_inst_issued.ratio']


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    vals = rows[2 + idx]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    print("kernel:", d.get('Kernel Name', ('', '?'))[1][:140])
    for k in WANT:
        if k in d:
            print(f'{k:72s} {d[k][0]:12s} {d[k][1]}')
    for h in hdr:
        if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio') or ('warps_issue_stalled' in h and 'pct' in h):
            print(f'{h:72s} {d[h][0]:12s} {d[h][1]}')


if __name__ == '__main__':
    main()
