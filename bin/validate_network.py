#!/usr/bin/env python3
"""Validate a network on an HDF5 file of labelled chunks, forward pass on the B200 (reference
`bin/validate_network.py`): same options, same progress / final lines (mean loss, accuracy, kev/s)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class AutoBool(argparse.Action):
    def __init__(self, option_strings, dest, default=None, required=False, help=None):
        strings = []
        for s in option_strings:
            strings.append(s)
            if s.startswith('--'):
                strings.append('--no-' + s[2:])
        super().__init__(option_strings=strings, dest=dest, nargs=0, const=None, default=default, type=bool,
                         choices=None, required=required, help=help)

    def __call__(self, parser, namespace, values, option_string=None):
        setattr(namespace, self.dest, not option_string.startswith('--no-'))


def build_parser():
    parser = argparse.ArgumentParser(description='Validate a simple neural network (B200)',
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('--bad', default=True, action=AutoBool, help='Use bad events as a separate state')
    parser.add_argument('--batch', default=200, metavar='size', type=int,
                        help='Batch size (number of chunks to run in parallel)')
    parser.add_argument('--transducer', default=True, action=AutoBool, help='Model is a transducer')
    parser.add_argument('model', help='File to read model description from')
    parser.add_argument('input', help='HDF5 file containing chunks')
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    from sloika_b200 import helpers, validate
    from sloika_b200.fast5 import H5File
    sys.stdout.write('* Loading network from {}\n'.format(args.model))
    fv = validate.wrap_network(helpers.load_calc_post(args.model))
    sys.stdout.write('* Loading data from {}\n'.format(args.input))
    h5 = H5File(args.input)
    full_chunks = h5.dataset('/chunks')
    full_labels = np.array(h5.dataset('/labels'))
    full_bad = np.array(h5.dataset('/bad')).astype(bool)
    if not args.transducer:
        validate.remove_blanks(full_labels)
    if args.bad:
        full_labels[full_bad] = 0
    total_ev = line_ev = 0
    score = acc = wacc = wscore = 0.0
    t1 = t0 = time.time()
    sys.stdout.write('* Validating\n')
    nbatch = len(full_chunks) // args.batch
    for i in range(nbatch):
        idx = i * args.batch
        events = np.ascontiguousarray(full_chunks[idx:idx + args.batch].transpose((1, 0, 2)))
        labels = np.ascontiguousarray(full_labels[idx:idx + args.batch].transpose())
        fval, ncorr = fv(events, labels)
        nev = np.size(labels)
        line_ev += nev
        total_ev += nev
        score += fval
        wscore += 1
        acc += ncorr
        wacc += nev
        sys.stdout.write('.')
        if (i + 1) % 50 == 0:
            tn = time.time()
            dt = tn - t1
            sys.stdout.write(' {:5d} {:5.3f}  {:5.2f}%  {:5.2f}s ({:.2f} kev/s)\n'.format(
                (i + 1) // 50, score / wscore, 100.0 * acc / wacc, dt, line_ev / 1000.0 / dt))
            line_ev = 0
            t1 = tn
    dt = time.time() - t0
    sys.stdout.write('\nFinal {:5.3f}  {:5.2f}%  {:5.2f}s ({:.2f} kev/s)\n'.format(
        score / max(wscore, 1), 100.0 * acc / max(wacc, 1), dt, total_ev / 1000.0 / dt))


if __name__ == '__main__':
    main()
