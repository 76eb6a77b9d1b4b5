#!/usr/bin/env python3
"""Golden fixtures for the remap decode from the REFERENCE itself.

Run in the build container, where /root/reference is mounted:

    python tools/make_golden_remap.py

`sloika/transducer.py` (map_to_sequence) is executed UNMODIFIED from /root/reference.  It needs the
reference's only native module, `sloika/viterbi_helpers.pyx`, which is compiled with Cython into a
scratch directory OUTSIDE the repository (default /tmp/refbuild); the one edit made to the scratch copy is
`np.int` -> `np.int64` (the alias was removed from NumPy 1.24+; same type on this platform).  The
reference's `util.geometric_prior` (util.py) supplies the priors exactly as `tools/chunkify_raw.py:268-274`
builds them.  Stored: inputs and the reference's outputs (score, path; slip_update's from_score / from_pos),
`tests/golden/remap_cases.npz` + `.json`.  Nothing is read from /root/reference at test time.
"""
import importlib.util
import json
import os
import subprocess
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
SCRATCH = os.environ.get('SLOIKA_REFBUILD', '/tmp/refbuild')
GOLD = os.path.join(ROOT, 'tests', 'golden')


def build_reference_helpers():
    os.makedirs(os.path.join(SCRATCH, 'sloika'), exist_ok=True)
    src = open(os.path.join(REF, 'sloika', 'viterbi_helpers.pyx')).read()
    src = src.replace('ITYPE = np.int\n', 'ITYPE = np.int64\n').replace('ctypedef np.int_t ITYPE_t', 'ctypedef np.int64_t ITYPE_t')
    with open(os.path.join(SCRATCH, 'sloika', 'viterbi_helpers.pyx'), 'w') as fh:
        fh.write(src)
    with open(os.path.join(SCRATCH, 'setup.py'), 'w') as fh:
        fh.write("from setuptools import setup, Extension\nfrom Cython.Build import cythonize\nimport numpy as np\n"
                 "setup(ext_modules=cythonize([Extension('viterbi_helpers', ['sloika/viterbi_helpers.pyx'], "
                 "include_dirs=[np.get_include()])], language_level=3))\n")
    subprocess.run([sys.executable, 'setup.py', 'build_ext', '--inplace'], cwd=SCRATCH, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    sys.path.insert(0, SCRATCH)
    import viterbi_helpers
    return viterbi_helpers


def load_reference(vh):
    """Import the reference's transducer.py / util.geometric_prior with a stub `sloika` package around them."""
    pkg = types.ModuleType('sloika')
    cfg = types.ModuleType('sloika.config')
    cfg.sloika_dtype = np.float32                      # config.py: THEANO floatX = float32 on the basecall path
    pkg.config, pkg.viterbi_helpers = cfg, vh
    sys.modules.update({'sloika': pkg, 'sloika.config': cfg, 'sloika.viterbi_helpers': vh})
    spec = importlib.util.spec_from_file_location('sloika.transducer', os.path.join(REF, 'sloika', 'transducer.py'))
    tr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tr)
    # geometric_prior is a self-contained function of util.py: execute just its source
    usrc = open(os.path.join(REF, 'sloika', 'util.py')).read()
    start = usrc.index('def geometric_prior')
    end = usrc.index('\ndef ', start + 1)
    ns = {'np': np}
    exec(usrc[start:end], ns)
    return tr, ns['geometric_prior']


def main():
    vh = build_reference_helpers()
    tr, geometric_prior = load_reference(vh)
    rng = np.random.default_rng(20260101)
    arrays, meta = {}, []

    # ---- slip_update: the reference test's own case (test/unit/test_viterbi.py:10-16) + tie-heavy ones
    np.random.seed(0xdeadbeef)
    xs = [np.random.normal(size=10).astype(np.float32)]
    xs.append(rng.integers(-3, 4, size=40).astype(np.float32))           # quantised: >= ties
    xs.append((rng.standard_normal(300) * 20).astype(np.float32))
    xs.append(np.zeros(7, dtype=np.float32))
    for n, x in enumerate(xs):
        for slip in (5.0, 0.0, 0.5):
            fs, fp = vh.slip_update(x, np.float32(slip))
            key = 'slip{}_{}'.format(n, str(slip).replace('.', 'p'))
            arrays[key + '_x'], arrays[key + '_score'], arrays[key + '_pos'] = x, fs, fp.astype(np.int64)
            meta.append({'kind': 'slip_update', 'key': key, 'slip': slip})

    # ---- map_to_sequence
    def case(name, nev, nstate, npos, slip, prior, log, quant=False, peaky=1.0):
        logits = rng.standard_normal((nev, nstate)) * peaky
        if quant:
            logits = np.round(logits * 2) / 2
        p = np.exp(logits - logits.max(axis=1, keepdims=True))
        p = (p / p.sum(axis=1, keepdims=True)).astype(np.float32)
        if quant:
            trans = (np.round(np.log(p) * 2) / 2).astype(np.float32)      # many exact ties
            is_log = True
        elif log:
            trans, is_log = np.log(p).astype(np.float32), True
        else:
            trans, is_log = (np.float32(1e-5) + np.float32(1 - 1e-5) * p).astype(np.float32), False
        seq = rng.integers(1, nstate, size=npos)
        p0 = geometric_prior(npos, prior[0]) if prior[0] is not None else None
        p1 = geometric_prior(npos, prior[1], rev=True) if prior[1] is not None else None
        score, path = tr.map_to_sequence(trans, list(seq), slip=slip, prior_initial=p0, prior_final=p1, log=is_log)
        arrays[name + '_trans'], arrays[name + '_seq'] = trans, seq.astype(np.int32)
        arrays[name + '_path'], arrays[name + '_score'] = np.asarray(path, dtype=np.int32), np.float32(score)
        if p0 is not None:
            arrays[name + '_prior0'] = p0
        if p1 is not None:
            arrays[name + '_prior1'] = p1
        meta.append({'kind': 'map_to_sequence', 'key': name, 'slip': slip, 'log': is_log,
                     'prior': [prior[0], prior[1]], 'nev': nev, 'npos': npos, 'nstate': nstate})
        print('{:14s} nev {:4d} npos {:4d} slip {} score {:.4f} moves {}'.format(
            name, nev, npos, slip, float(score), int(np.sum(np.diff(path) != 0))))

    case('tiny_noslip', 12, 17, 6, None, (None, None), True)
    case('tiny_slip', 12, 17, 6, 5.0, (None, None), True)
    case('small_slip0', 40, 65, 25, 0.0, (None, None), True)
    case('small_prior', 60, 65, 30, 5.0, (25.0, 25.0), False)
    case('ties', 80, 17, 40, 1.0, (None, None), True, quant=True)
    case('ties_noslip', 50, 17, 30, None, (None, None), True, quant=True)
    case('mid', 300, 65, 150, 5.0, (25.0, None), False, peaky=3.0)
    case('mid_log', 250, 65, 200, 2.5, (None, 10.0), True, peaky=3.0)
    case('long_seq', 64, 65, 333, 5.0, (None, None), True)
    case('three_pos', 20, 17, 3, 5.0, (None, None), True)

    np.savez_compressed(os.path.join(GOLD, 'remap_cases.npz'), **arrays)
    with open(os.path.join(GOLD, 'remap_cases.json'), 'w') as fh:
        json.dump(meta, fh, indent=1)
    print('wrote', len(meta), 'cases,', os.path.getsize(os.path.join(GOLD, 'remap_cases.npz')), 'bytes')


if __name__ == '__main__':
    main()
