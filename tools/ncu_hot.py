#!/usr/bin/env python3
"""Top SASS instructions of a kernel by warp-stall samples, from `ncu --page source --csv` of a .ncu-rep.

    python tools/ncu_hot.py prof.ncu-rep regex:kernel [launch-skip] [top-n]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else '0'
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', kern, '--launch-skip', skip,
                          '--launch-count', '1'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    i_src, i_smp, i_ins = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    data = []
    for idx, r in enumerate(rows[2:]):
        if len(r) <= i_ins:
            continue
        try:
            smp = float(r[i_smp] or 0)
        except ValueError:
            continue
        stalls = sorted(((float(r[i] or 0), h[6:]) for i, h in stall_cols), reverse=True)[:2]
        data.append((smp, idx, float(r[i_ins] or 0), r[i_src].strip(), stalls))
    total = sum(d[0] for d in data) or 1.0
    print("kernel {} : {} SASS lines, {} samples".format(kern, len(data), int(total)))
    for smp, idx, ins, src, stalls in sorted(data, reverse=True)[:top]:
        st = ', '.join('{}:{:.0f}'.format(n, v) for v, n in stalls if v > 0)
        print("{:6.2f}%  line {:5d}  exec {:11.0f}  {:70s} [{}]".format(100 * smp / total, idx, ins, src[:70], st))


if __name__ == '__main__':
    main()
