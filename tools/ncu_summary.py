#!/usr/bin/env python3
"""Summarise .ncu-rep captures (read here with `ncu -i ... --page raw --csv`) into a small text table.

    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [...] > profiles/r1_x.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('sm__cycles_elapsed.max', 'cycles'),
    ('dram__bytes_read.sum', 'dram read'),
    ('dram__bytes_write.sum', 'dram write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram % of peak'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm throughput %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'),
    ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'fma pipe active %'),
    ('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'alu pipe %'),
    ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lsu pipe %'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe active %'),
    ('sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active', 'tensor inst %'),
    ('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex lsu wavefronts %'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smem wavefronts'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem bank conflicts'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('launch__registers_per_thread', 'registers/thread'),
    ('launch__shared_mem_per_block_dynamic', 'dyn smem/block'),
    ('launch__occupancy_limit_registers', 'occupancy limit (regs)'),
    ('sm__maximum_warps_per_active_cycle_pct', 'theoretical occupancy %'),
]


def rows_of(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        yield dict(zip(hdr, r)), dict(zip(hdr, units))


def main():
    for path in sys.argv[1:]:
        print("# {}".format(path))
        for d, u in rows_of(path):
            print("== {}  grid {} block {}".format(d.get('Kernel Name', '?')[:90], d.get('Grid Size'), d.get('Block Size')))
            for key, label in KEYS:
                if key in d and d[key] != '':
                    print("   {:28s} {:>18s} {}".format(label, d[key], u.get(key, '')))
        print()


if __name__ == '__main__':
    main()
