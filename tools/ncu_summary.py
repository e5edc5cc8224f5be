#!/usr/bin/env python
"""Summarise an .ncu-rep (read here on the CPU box with `ncu -i`) into the markdown kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep "title" > profiles/r01_x.md
"""
import csv
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'), ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
    ('launch__registers_per_thread', 'registers/thread'), ('launch__shared_mem_per_block_dynamic', 'dynamic smem/block'),
    ('launch__occupancy_limit_registers', 'occupancy limit (regs), CTAs/SM'), ('launch__occupancy_limit_shared_mem', 'occupancy limit (smem), CTAs/SM'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('dram__bytes_read.sum', 'DRAM bytes read'), ('dram__bytes_write.sum', 'DRAM bytes written'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput % of ncu peak'),
    ('dram__bytes_read.sum.per_second', 'DRAM read rate'), ('dram__bytes_write.sum.per_second', 'DRAM write rate'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'),
    ('smsp__inst_executed.sum', 'warp instructions executed'),
    ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'LSU pipe %'),
    ('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'L1 data-pipe LSU wavefronts %'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'shared-memory wavefronts'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'shared-memory bank conflicts'),
    ('smsp__inst_executed_op_shared_atom.sum', 'shared atomic instructions'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit rate %'), ('l1tex__t_sector_hit_rate.pct', 'L1 hit rate %'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe % (expected 0)'),
    ('sm__cycles_elapsed.avg.per_second', 'SM clock'), ('dram__cycles_elapsed.avg.per_second', 'DRAM clock'),
]
STALLS = ['long_scoreboard', 'short_scoreboard', 'wait', 'not_selected', 'math_pipe_throttle', 'mio_throttle', 'lg_throttle', 'barrier',
          'branch_resolving', 'no_instruction', 'dispatch_stall', 'drain', 'imc_miss', 'membar', 'tex_throttle', 'sleeping']


def main():
    rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f'# {title}\n\nSource: `{rep}` (ncu --set full --clock-control none --import-source on; read with `ncu -i ... --page raw --csv`).\n')
    for r in rows[2:]:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        print(f"## {d.get('Kernel Name', '?')}\n")
        print('| metric | value | unit |\n|---|---|---|')
        for k, label in KEYS:
            if k in d and d[k] != '': print(f'| {label} (`{k}`) | {d[k]} | {u.get(k, "")} |')
        print('\nWarp stall reasons (cycles per issued instruction):\n')
        print('| ' + ' | '.join(STALLS) + ' |\n|' + '---|' * len(STALLS))
        print('| ' + ' | '.join((d.get(f'smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio', '') or '-')[:6] for s in STALLS) + ' |\n')


if __name__ == '__main__':
    main()
