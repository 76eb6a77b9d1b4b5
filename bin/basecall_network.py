#!/usr/bin/env python3
"""1D basecaller for RNNs on B200 -- the `raw` sub-command of the reference's `bin/basecall_network.py`.

Same command line (reference `bin/basecall_network.py:17-75`):

    basecall_network.py raw [--alphabet ACGT] [--compile FILE] [--input_strand_list FILE] [--jobs n]
                            [--kmer_len 5] [--limit reads] [--min_prob 1e-5] [--skip 0.0]
                            [--transducer | --no-transducer] [--bad | --no-bad]
                            [--open_pore_fraction 0] [--trim 200 10]  model input_folder

FASTA records go to stdout and the throughput line to stderr exactly as the reference prints them
(`:102-111`).  Differences: the model is any Sloika model pickle (loaded without Theano); reads are fed
to the GPU in ragged whole-read batches of `--batch` files instead of one per call; `--jobs` is accepted
for compatibility (one process drives one GPU; under torchrun each rank takes a shard of the files and
rank 0 prints the gathered calls).  The `events` sub-command is not on the B200 path.
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class AutoBool(argparse.Action):
    """--flag / --no-flag pair (reference `sloika/cmdargs.py:137-171`)."""

    def __init__(self, option_strings, dest, default=None, required=False, help=None):
        if default is None:
            raise ValueError('You must provide a default with AutoBool action')
        if len(option_strings) != 1:
            raise ValueError('Only single argument is allowed with AutoBool action')
        opt = option_strings[0]
        if not opt.startswith('--'):
            raise ValueError('AutoBool arguments must be prefixed with --')
        name = opt[2:]
        opts = ['--' + name, '--no-' + name]
        default_opt = opts[0] if default else opts[1]
        super(AutoBool, self).__init__(opts, dest, nargs=0, const=None, default=default, required=required,
                                       help='{} (Default: {})'.format(help, default_opt))

    def __call__(self, parser, namespace, values, option_strings=None):
        setattr(namespace, self.dest, not option_strings.startswith('--no-'))


def positive_int(x):
    v = int(x)
    if v <= 0:
        raise argparse.ArgumentTypeError('{} is not positive'.format(x))
    return v


def non_negative(kind):
    def conv(x):
        v = kind(x)
        if v < 0:
            raise argparse.ArgumentTypeError('{} is negative'.format(x))
        return v
    return conv


def proportion(x):
    v = float(x)
    if not 0.0 <= v <= 1.0:
        raise argparse.ArgumentTypeError('{} not in [0, 1]'.format(x))
    return v


def existing(path):
    if not os.path.exists(path):
        raise argparse.ArgumentTypeError("File/path does not exist, {}".format(path))
    return path


def _common(p):
    """Options shared by both sub-commands (reference `bin/basecall_network.py:22-48`)."""
    p.add_argument('--alphabet', default='ACGT', help='Alphabet of the sequences')
    p.add_argument('--compile', default=None, help='File output compiled model')
    p.add_argument('--input_strand_list', default=None, type=existing, help='Strand summary file containing subset')
    p.add_argument('--jobs', default=1, metavar='n', type=positive_int,
                   help='Processes parsing fast5 files (the network itself runs on the GPU of this process)')
    p.add_argument('--kmer_len', default=5, metavar='length', type=positive_int, help='Length of kmer')
    p.add_argument('--limit', default=None, metavar='reads', type=positive_int, help='Limit number of reads to process')
    p.add_argument('--min_prob', metavar='proportion', default=1e-5, type=proportion,
                   help='Minimum allowed probabiility for basecalls')
    p.add_argument('--skip', default=0.0, type=non_negative(float), help='Skip penalty')
    p.add_argument('--trans', default=None, type=proportion, nargs=3, metavar=('stay', 'step', 'skip'),
                   help='Base transition probabilities (non-transducer models)')
    p.add_argument('--transducer', default=True, action=AutoBool, help='Model is transducer')
    p.add_argument('model', type=existing, help='Pickled model file')
    p.add_argument('input_folder', type=existing, help='Directory containing single-read fast5 files')


def build_parser():
    parser = argparse.ArgumentParser(description='1D basecaller for RNNs (B200)',
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    sub = parser.add_subparsers(help='command', dest='command')
    sub.required = True
    ev = sub.add_parser('events', help='basecall from events', formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    _common(ev)
    ev.add_argument('--bad', default=True, action=AutoBool, help='Model emits bad events as a separate state')
    ev.add_argument('--section', default='template', choices=['template', 'complement'], help='Section to call')
    ev.add_argument('--segmentation', default='Segmentation', metavar='location',
                    help='Location of segmentation information')
    ev.add_argument('--trim', default=(50, 1), nargs=2, type=non_negative(int), metavar=('beginning', 'end'),
                    help='Number of events to trim off start and end')
    ev.set_defaults(datatype='events')
    raw = sub.add_parser('raw', help='basecall from raw signal', formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    _common(raw)
    raw.add_argument('--bad', default=True, action=AutoBool, help='Model emits bad signal blocks as a separate state')
    raw.add_argument('--open_pore_fraction', metavar='proportion', default=0, type=proportion,
                     help='Max fraction of signal to trim due to open pore')
    raw.add_argument('--trim', default=(200, 10), nargs=2, type=non_negative(int), metavar=('beginning', 'end'),
                     help='Number of samples to trim off start and end')
    raw.add_argument('--batch', default=64, type=positive_int, help='Reads per device batch')
    raw.set_defaults(datatype='samples')
    return parser


def main(argv=None):
    args = build_parser().parse_args(argv)
    if os.path.exists(args.compile or ''):
        raise RuntimeError("File/path for 'compile' exists, {}".format(args.compile))       # FileAbsent
    from sloika_b200 import basecall
    reader_pool = basecall.make_reader_pool(args.jobs)         # forked before CUDA is initialised
    import torch
    from sloika_b200 import helpers, sharding
    from sloika_b200.fast5 import iterate_fast5

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group('nccl')

    # under torchrun every rank would check and write the same --compile file: rank 0 writes it, the others compile
    # into a private temporary file
    compiled_file = helpers.compile_model(args.model, args.compile if rank == 0 else None)
    basecall.init_worker(compiled_file)
    seq_printer = basecall.SeqPrinter(args.kmer_len, datatype=args.datatype, transducer=args.transducer,
                                      alphabet=args.alphabet)
    files = list(iterate_fast5(args.input_folder, paths=True, limit=args.limit, strand_list=args.input_strand_list))
    shards = sharding.partition_reads([os.path.getsize(f) for f in files], world)
    mine = shards[rank]

    nbases = nevents = 0
    t0 = time.time()
    results = []
    if args.command == 'events':
        # the events route (reference `basecall.events_worker`): one read per call, as the reference does
        for i in mine:
            results.append(basecall.events_worker(files[i], section=args.section, segmentation=args.segmentation,
                                                  trim=tuple(args.trim), kmer_len=args.kmer_len,
                                                  transducer=args.transducer, bad=args.bad, min_prob=args.min_prob,
                                                  alphabet=args.alphabet, skip=args.skip, trans=args.trans))
    chunks = [] if args.command == 'events' else \
        [[files[i] for i in mine[lo:lo + args.batch]] for lo in range(0, len(mine), args.batch)]
    # with --jobs the files of the next batch are parsed by the pool while this batch is on the device
    ahead = basecall.read_files(chunks[0], reader_pool, wait=False) if (reader_pool is not None and chunks) else None
    for k, chunk in enumerate(chunks):
        loaded = None
        if ahead is not None:
            loaded = ahead.get()
            ahead = basecall.read_files(chunks[k + 1], reader_pool, wait=False) if k + 1 < len(chunks) else None
        results.extend(basecall.raw_batch(chunk, trim=tuple(args.trim), open_pore_fraction=args.open_pore_fraction,
                                          kmer_len=args.kmer_len, transducer=args.transducer, bad=args.bad,
                                          min_prob=args.min_prob, alphabet=args.alphabet, skip=args.skip,
                                          trans=args.trans, loaded=loaded))
    results = sharding.gather_results(mine, results, len(files))
    if rank == 0:
        for res in results:
            if res is None:
                continue
            read, score, call, nev = res
            nbases += seq_printer.write(read, score, call, nev)
            nevents += nev
        dt = time.time() - t0
        t = 'Called {} bases in {:.1f} s ({:.1f} bases/s or {:.1f} {}/s)\n'
        sys.stderr.write(t.format(nbases, dt, nbases / dt, nevents / dt, args.datatype))
    if reader_pool is not None:
        reader_pool.close()
    if compiled_file != args.compile:
        os.remove(compiled_file)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
