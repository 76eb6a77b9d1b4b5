"""The drop-in driver `bin/basecall_network.py raw` and the fast5 reader on fabricated single-read files
(written from the committed DAQ fixtures by tests/h5write.py, so this also runs on the GPU box)."""
import io
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, scaled_signal
from h5write import write_fast5

sys.path.insert(0, os.path.join(ROOT, 'bin'))
NAMES = ['read{}'.format(i) for i in range(1, 9)]


@pytest.fixture(scope='module')
def fast5_dir(tmp_path_factory, reads_daq):
    d = tmp_path_factory.mktemp('reads')
    for i, name in enumerate(NAMES):
        off, rng, digi = reads_daq[name + '_scaling']
        write_fast5(str(d / (name + '.fast5')), reads_daq[name], off, rng, digi, read_number=100 + i)
    (d / 'broken.fast5').write_bytes(b'\x89HDF\r\n\x1a\n' + b'\0' * 40)
    return d


def test_fast5_roundtrip_and_iteration(fast5_dir, reads_daq, tmp_path):
    from sloika_b200.fast5 import Fast5, iterate_fast5
    files = list(iterate_fast5(str(fast5_dir), paths=True))
    assert [os.path.basename(f) for f in files] == sorted(n + '.fast5' for n in NAMES + ['broken'])
    for name in ('read1', 'read7'):
        with Fast5(str(fast5_dir / (name + '.fast5'))) as f5:
            assert f5.filename_short == name
            np.testing.assert_array_equal(f5.get_read(raw=True, scale=False), reads_daq[name])
            np.testing.assert_array_equal(f5.get_read(raw=True), scaled_signal(reads_daq, name))
    strands = tmp_path / 'strands.txt'
    strands.write_text('filename\tother\nread3.fast5\tx\nread8.fast5\ty\n')
    picked = list(iterate_fast5(str(fast5_dir), paths=True, strand_list=str(strands)))
    assert [os.path.basename(f) for f in picked] == ['read3.fast5', 'read8.fast5']


def test_cli_arguments_match_reference_defaults(tmp_path):
    import basecall_network as cli
    model = tmp_path / 'm.pkl'
    model.write_bytes(b'x')
    args = cli.build_parser().parse_args(['raw', str(model), str(tmp_path)])
    # defaults of bin/basecall_network.py:24-75
    assert (args.kmer_len, args.min_prob, args.skip, args.transducer, args.bad) == (5, 1e-5, 0.0, True, True)
    assert tuple(args.trim) == (200, 10) and args.open_pore_fraction == 0 and args.jobs == 1
    assert args.alphabet == 'ACGT' and args.limit is None and args.datatype == 'samples'
    args = cli.build_parser().parse_args(['raw', '--no-transducer', '--trim', '50', '1', '--skip', '2.5', '--limit', '3',
                                          str(model), str(tmp_path)])
    assert args.transducer is False and list(args.trim) == [50, 1] and args.skip == 2.5 and args.limit == 3
    with pytest.raises(SystemExit):
        cli.build_parser().parse_args(['raw', '--min_prob', '1.5', str(model), str(tmp_path)])
    with pytest.raises(SystemExit):
        cli.build_parser().parse_args(['raw', str(tmp_path / 'missing.pkl'), str(tmp_path)])


@pytest.mark.gpu
def test_cli_end_to_end_matches_golden_fasta(fast5_dir, pretrained, read_basecalls, tmp_path):
    """`basecall_network.py raw model dir`: FASTA on stdout identical to the golden records (made with the
    reference's decode.py / bio.py), unreadable file reported on stderr and skipped, summary line on stderr."""
    model = tmp_path / 'pretrained.b200.pkl'
    with open(model, 'wb') as fh:
        pickle.dump(pretrained, fh)
    proc = subprocess.run([sys.executable, os.path.join(ROOT, 'bin', 'basecall_network.py'), 'raw', '--batch', '5',
                           str(model), str(fast5_dir)], capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stderr[-2000:]
    expect = ''.join(read_basecalls[n]['header'] + '\n' + read_basecalls[n]['seq'] + '\n' for n in sorted(NAMES))
    assert proc.stdout == expect
    # --jobs: the files are parsed by a forked pool, same records in the same order
    par = subprocess.run([sys.executable, os.path.join(ROOT, 'bin', 'basecall_network.py'), 'raw', '--batch', '3',
                          '--jobs', '3', str(model), str(fast5_dir)], capture_output=True, text=True, timeout=600)
    assert par.returncode == 0, par.stderr[-2000:]
    assert par.stdout == expect and 'broken.fast5' in par.stderr
    assert 'Error getting raw data for file' in proc.stderr and 'broken.fast5' in proc.stderr
    nbases = sum(len(read_basecalls[n]['seq']) for n in NAMES)
    assert 'Called {} bases in'.format(nbases) in proc.stderr and 'samples/s' in proc.stderr
