"""Model file -> `calc_post` (reference `sloika/helpers.py:10-79`).

The reference compiles the Theano graph in a child process and hands back the name of a pickle of the
compiled function; `basecall.init_worker` then unpickles it into the module-global `calc_post`
(`sloika/basecall.py:12-23`).  Here "compiling" is loading the weights: `compile_model` validates the
model file (same ValueError on a bad file, `helpers.py:73-75`) and returns a path that
`basecall.init_worker` can open; `--compile FILE` persists a Theano-free copy of the model
(layer tree + float32 weights) that loads faster than walking the Theano stubs.
"""
import os
import pickle
import tempfile

from sloika_b200 import model_io


def load_calc_post(model_file, device=None):
    """Model pickle -> compiled network callable on `device` (default: current CUDA device)."""
    network = model_io.load_model(model_file)
    compiled = network.compile()
    if device is not None:
        compiled.to(device)
    return compiled


def compile_model(model_file, output_file=None):
    """Validate `model_file` and write the B200 form of the model next to it.

    :returns: name of the file to give to `basecall.init_worker` (a temporary file unless
        `output_file` is given; the caller removes temporaries, `bin/basecall_network.py:113-114`)
    """
    try:
        network = model_io.load_model(model_file)
    except Exception as err:
        raise ValueError("model in file {} could not be loaded: {}".format(model_file, err))
    if output_file is None:
        fd, output_file = tempfile.mkstemp(suffix='.b200.pkl', dir=os.environ.get('TMPDIR'))
        os.close(fd)
    with open(output_file, 'wb') as fh:
        pickle.dump(network, fh, protocol=pickle.HIGHEST_PROTOCOL)
    return output_file
