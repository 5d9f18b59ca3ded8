'''
Turn the .ncu-rep captures brought back in gpurun_out/ into the small text/CSV summaries committed here.
Usage: python profiles/summarize_ncu.py gpurun_out/r01_rk4.ncu-rep profiles/r01_ncu_rk4_cells.txt
'''
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__sass_thread_inst_executed_op_dfma_pred_on.sum', 'sm__sass_thread_inst_executed_op_dmul_pred_on.sum',
        'sm__sass_thread_inst_executed_op_dadd_pred_on.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def main(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, 'w') as fh:
        fh.write(f'# ncu --set full --clock-control none summary of {rep.split("/")[-1]} (values per launch)\n')
        for r in rows[2:]:
            name = r[hdr.index('Kernel Name')]
            fh.write(f'\n== {name}\n')
            for k in KEYS:
                if k in hdr:
                    fh.write(f'{k:75s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}\n')
            fh.write('warp stall reasons (warps per issue-active cycle):\n')
            for i, k in enumerate(hdr):
                if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k:
                    try:
                        v = float(r[i])
                    except ValueError:
                        continue
                    if v >= 0.1:
                        short = k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')
                        fh.write(f'    {short:40s} {v:8.2f}\n')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
