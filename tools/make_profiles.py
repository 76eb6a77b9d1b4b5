#!/usr/bin/env python3
"""Turn the captures of tools/profile_step.sh (gpurun_out/<tag>_*.ncu-rep, <tag>_launches.csv) into the tracked
evidence under profiles/: a text summary of the `--set full` captures, the launch list, DRAM bytes per launch per
kernel class (what `roofline.traffic` of bench.py reads) and the SASS instruction counts of the built library.

    python tools/make_profiles.py r2a r2      # capture tag, name of the files under profiles/
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import ncu_summary          # noqa: E402

CLASSES = [('gru_seq_kernel', 'gru_seq'), ('block_layout', 'block_layout'), ('gru_fused_kernel', 'gru_fused'), ('gru_tc_kernel', 'gru_recurrence'), ('gru_h16_kernel', 'gru_recurrence'), ('viterbi', 'viterbi'),
           ('conv1d', 'conv1d')]


def kernel_class(name, dram_write):
    for key, cls in CLASSES:
        if key in name:
            return cls
    if 'gemm_tf32x3' in name:
        return 'softmax' if ', 1, ' in name or '<0, 1' in name else 'gru_projection'
    return None


def main():
    tag, out = sys.argv[1], sys.argv[2]
    src = os.path.join(ROOT, 'gpurun_out')
    dst = os.path.join(ROOT, 'profiles')
    reps = [os.path.join(src, f) for f in sorted(os.listdir(src)) if f.startswith(tag + '_') and f.endswith('.ncu-rep')]
    traffic = {}
    with open(os.path.join(dst, out + '_step_kernels_ncu.txt'), 'w') as fh:
        fh.write("# ncu --set full --clock-control none summaries (tools/profile_step.sh -> tools/make_profiles.py); one step of\n"
                 "# `python bench.py` (raw_rgrgr, 1024 chunks x 4000 samples), kernels serialised by the profiler.\n"
                 "# gru_if1 = recurrence as launched with one batch in flight (G = 1 group of 8 sequences per CTA, 128 CTAs),\n"
                 "# gru_if4 = first GRU layer with four batches in flight (G = 2 groups of 16 sequences per CTA, 32 CTAs per batch),\n"
                 "# gru_fused = layers 2-5 with four batches in flight: projection inside the launch (3-CTA clusters, 48 CTAs),\n"
                 "# gru_seq = the layers with >= 6 batches in flight: 128 sequences on the TMEM lanes (8 CTAs per batch), layout = the\n"
                 "# row-major <-> blocked conversions around them.\n")
        for path in reps:
            fh.write("# {}\n".format(os.path.basename(path)))
            for d, u in ncu_summary.rows_of(path):
                name = d.get('Kernel Name', '?')
                fh.write("== {}  grid {} block {}\n".format(name[:100], d.get('Grid Size'), d.get('Block Size')))
                for key, label in ncu_summary.KEYS:
                    if key in d and d[key] != '':
                        fh.write("   {:28s} {:>18s} {}\n".format(label, d[key], u.get(key, '')))
                try:
                    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
                    rd = float(d['dram__bytes_read.sum']) * scale[u['dram__bytes_read.sum']]
                    wr = float(d['dram__bytes_write.sum']) * scale[u['dram__bytes_write.sum']]
                    cls = kernel_class(name, wr)
                    if 'gemm_tf32x3' in name:
                        cls = 'softmax' if wr > 2e9 else ('gru_projection' if wr > 1e8 else None)
                    if cls and os.path.basename(path) != tag + '_gru_if4.ncu-rep':
                        traffic.setdefault(cls, []).append(rd + wr)
                except (KeyError, ValueError):
                    pass
            fh.write("\n")
    with open(os.path.join(dst, out + '_traffic.json'), 'w') as fh:
        json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, fh, indent=1)
    launches = os.path.join(src, tag + '_launches.csv')
    if os.path.exists(launches):
        with open(launches) as fi, open(os.path.join(dst, out + '_launches.csv'), 'w') as fo:
            for line in fi:
                if line.startswith('"') or line.startswith('=='):
                    fo.write(line)
    # SASS evidence: which kernels carry Blackwell tensor-core / TMA instructions
    lib = os.path.join(ROOT, 'sloika_b200', 'libsloika_b200.so')
    sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    pats = ['UTCHMMA', 'UTCQMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'HMMA', 'LDGSTS', 'SYNCS']
    counts, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = dict.fromkeys(pats, 0)
            continue
        if cur:
            for p in pats:
                if re.search(r'\b' + p + r'\b|\b' + p + r'\.', line):
                    counts[cur][p] += 1
    with open(os.path.join(dst, out + '_sass_counts.txt'), 'w') as fh:
        fh.write("# cuobjdump -sass sloika_b200/libsloika_b200.so: instruction counts per kernel (sm_100a)\n"
                 "# UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store,\n"
                 "# UBLKCP = cp.async.bulk (1-D TMA), HMMA = legacy mma.sync, LDGSTS = cp.async, SYNCS = mbarrier ops\n")
        fh.write("{:>8s} {:>6s} {:>6s} {:>8s} {:>8s} {:>7s} {:>6s} {:>7s} {:>6s}  kernel\n".format(
            'UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'HMMA', 'LDGSTS', 'SYNCS'))
        for name in sorted(counts):
            c = counts[name]
            demangled = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip()
            fh.write("{:8d} {:6d} {:6d} {:8d} {:8d} {:7d} {:6d} {:7d} {:6d}  {}\n".format(
                c['UTCHMMA'] + c['UTCQMMA'], c['LDTM'], c['STTM'], c['UTMALDG'], c['UTMASTG'], c['UBLKCP'], c['HMMA'],
                c['LDGSTS'], c['SYNCS'], demangled[:110]))
    print("wrote", out + '_step_kernels_ncu.txt', out + '_traffic.json', out + '_launches.csv', out + '_sass_counts.txt')


if __name__ == '__main__':
    main()
