#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics per kernel + warp-stall breakdown.
usage: tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_name.txt"""
import csv, subprocess, sys, io

KEYS = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.sum',
        'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum', 'smsp__inst_executed_op_global_ld.sum']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('== kernel:', d.get('Kernel Name', '?')[:100], '| grid', d.get('Grid Size'), 'block', d.get('Block Size'))
        for i, h in enumerate(hdr):
            if h in KEYS:
                print('  %-85s %-12s %s' % (h, units[i], r[i]))
        st = [(float(r[i] or 0), h) for i, h in enumerate(hdr)
              if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') or
              h.startswith('smsp__average_warp_latency_issue_stalled_') and h.endswith('.ratio')]
        st.sort(reverse=True)
        print('  -- warp stall reasons (warps stalled per issue-active cycle), top 10')
        for v, h in st[:10]:
            print('  %-85s %.3f' % (h, v))


if __name__ == '__main__':
    main(sys.argv[1])
