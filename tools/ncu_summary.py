"""Summarise an .ncu-rep (raw page) into the handful of counters the roofline discussion needs.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/xxx.md]
"""
import csv
import io
import subprocess
import sys

WANT = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'),
    ('launch__registers_per_thread', 'regs/thread'),
    ('launch__shared_mem_per_block_dynamic', 'dyn smem/block'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM % of peak'),
    ('lts__t_bytes.sum', 'L2 bytes'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor pipe active % (elapsed)'),
    ('sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed', 'UTCHMMA bf16 ops % of peak'),
    ('l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'MMA operand smem wavefronts % of peak'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'LSU smem wavefronts % of peak'),
    ('smsp__sass_inst_executed_op_utcmma.sum', 'UTCMMA instructions'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_elapsed', 'XU (MUFU) pipe %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_elapsed', 'issue slots busy %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
    ('sm__cycles_elapsed.avg', 'SM cycles'),
]


def main(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f'# ncu --set full summary: {path}\n')
    for r in rows[2:]:
        name = r[idx['Kernel Name']].split('(')[0].replace('void ynet::', '')
        print(f'## `{name}`  (id {r[idx["ID"]]})\n')
        print('| counter | value |')
        print('|---|---:|')
        for key, label in WANT:
            if key in idx and r[idx[key]] not in ('', 'n/a'):
                print(f'| {label} (`{key}`) | {r[idx[key]]} {units[idx[key]]} |')
        print()


if __name__ == '__main__':
    main(sys.argv[1])
