"""Load existing Sloika model pickles without Theano.

A Sloika model file is a pickle of a `sloika.layers.Layer` tree whose weights are Theano shared
variables (`sloika/helpers.py:31-39` loads it; `models/pretrained.pkl` is protocol 3).  The objects
are plain `__dict__` pickles, so they can be rebuilt with a custom `find_class`:

* `sloika.layers.<Cls>`            -> the same-named class in `sloika_b200.layers`
* `sloika.activation.<fn>`         -> the same-named activation token
* `theano.tensor.sharedvar.TensorSharedVariable`, `theano.sandbox.cuda.var.CudaNdarraySharedVariable`
  (GPU-trained pickles, `misc/model_convert.py:7-8`), `theano.tensor.type.TensorType`,
  `theano.gof.utils.scratchpad`, `theano.gof.link.Container`, ... -> inert stubs that just keep
  their `__dict__`; a weight is `var.container.storage[0]` (float32 ndarray)

after which every shared-variable stub is replaced by a `layers.Param` (host array + device mirror).
Pickles of *compiled* `theano.compile.function_module.Function` objects cannot be rebuilt without
Theano and are rejected with a clear error.
"""
import io
import pickle

import numpy as np

from sloika_b200 import activation, layers
from sloika_b200.config import sloika_dtype


class ModelFormatError(ValueError):
    pass


class _Stub(object):
    """Inert stand-in for a Theano object: keeps whatever state the pickle carries."""

    def __init__(self, *args, **kwargs):
        self._args = args

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        else:
            self._state = state


def _make_stub(module, name):
    return type(name, (_Stub,), {'__module__': module})


def _identity_unpickler(value, *rest):
    return value


_SHARED_NAMES = ('TensorSharedVariable', 'CudaNdarraySharedVariable', 'ScalarSharedVariable',
                 'SharedVariable', 'GpuArraySharedVariable')


_ALLOWED_GLOBALS = {
    ('numpy._core.multiarray', '_reconstruct'), ('numpy._core.multiarray', 'scalar'),
    ('numpy._core.numeric', '_frombuffer'),                    # how NumPy 2 pickles arrays (files we wrote ourselves)
    ('numpy', 'ndarray'), ('numpy', 'dtype'), ('copyreg', '_reconstructor'), ('copy_reg', '_reconstructor'),
    ('collections', 'OrderedDict'), ('builtins', 'object'), ('builtins', 'tuple'), ('builtins', 'list'),
    ('builtins', 'dict'), ('builtins', 'set'), ('builtins', 'frozenset'), ('builtins', 'bytearray'),
    ('__builtin__', 'object'), ('_codecs', 'encode'),
    ('sloika_b200.layers', 'Param'),
}


class _SloikaUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == 'sloika.layers' or module == 'sloika_b200.layers':
            cls = getattr(layers, name, None)
            if cls is None:
                raise ModelFormatError(
                    "model uses sloika.layers.{}, which is not on the raw basecall path "
                    "implemented for B200".format(name))
            return cls
        if module == 'sloika.activation' or module == 'sloika_b200.activation':
            fun = getattr(activation, name, None)
            if fun is None:
                raise ModelFormatError("unknown activation sloika.activation.{}".format(name))
            return fun
        if module.startswith('theano'):
            if 'function_module' in module or module.startswith('theano.compile'):
                raise ModelFormatError(
                    "model file is a pickled *compiled* Theano function ({}.{}); only "
                    "uncompiled sloika.layers model pickles can be loaded without Theano"
                    .format(module, name))
            if name.endswith('_unpickler'):
                return _identity_unpickler
            return _make_stub(module, name)
        if module.startswith('numpy.core'):
            # numpy >= 2 moved numpy.core -> numpy._core; keep old pickles quiet
            module = 'numpy._core' + module[len('numpy.core'):]
        # everything else a Sloika model pickle legitimately names; any other global (os.system, builtins.eval, ...)
        # is refused, so unpickling cannot be made to call arbitrary code
        if (module, name) in _ALLOWED_GLOBALS:
            return super().find_class(module, name)
        raise ModelFormatError("model pickle names {}.{}, which a Sloika model does not need".format(module, name))


def _is_shared(obj):
    return isinstance(obj, _Stub) and type(obj).__name__ in _SHARED_NAMES


def _shared_value(var):
    try:
        value = var.container.storage[0]
    except (AttributeError, IndexError, TypeError):
        raise ModelFormatError("shared variable without container.storage[0]")
    value = np.asarray(value)
    if value.dtype != np.dtype(sloika_dtype):
        value = value.astype(sloika_dtype)
    return np.ascontiguousarray(value)


def _adopt(layer):
    """Replace shared-variable stubs by Params, recursively; normalise numpy ints to python ints."""
    if not isinstance(layer, layers.Layer):
        raise ModelFormatError("model pickle does not contain a sloika.layers.Layer "
                               "(got {})".format(type(layer).__name__))
    for key, val in list(layer.__dict__.items()):
        if _is_shared(val):
            layer.__dict__[key] = layers.Param(_shared_value(val), getattr(val, 'name', None) or key)
        elif isinstance(val, np.integer):
            layer.__dict__[key] = int(val)
        elif isinstance(val, _Stub):
            raise ModelFormatError("unexpected Theano object {} in layer attribute {}"
                                   .format(type(val).__name__, key))
    if isinstance(layer, layers.Convolution):
        layer.padding = tuple(int(p) for p in layer.padding)
    for child in (layer._children() if isinstance(layer, layers._Container) else []):
        _adopt(child)
    return layer


def loads(data):
    """Rebuild a B200 layer tree from the bytes of a Sloika model pickle."""
    last = None
    for kwargs in ({}, {'encoding': 'latin1'}):   # py2 pickles: helpers.py:33-39
        try:
            obj = _SloikaUnpickler(io.BytesIO(data), **kwargs).load()
            break
        except ModelFormatError:
            raise
        except UnicodeDecodeError as err:
            last = err
    else:
        raise ModelFormatError("cannot decode model pickle: {!r}".format(last))
    return _adopt(obj)


def load_model(filename):
    """Load a Sloika model `.pkl` (uncompiled `sloika.layers` tree) as B200 layers."""
    with open(filename, 'rb') as fh:
        data = fh.read()
    try:
        return loads(data)
    except ModelFormatError:
        raise
    except Exception as err:
        raise ModelFormatError("failed to load model {}: {!r}".format(filename, err))


def network_from_script(path, **kwargs):
    """Execute a reference model script (`models/*.py`) against the B200 layer classes and call its
    `network(klen, sd, ...)` factory.  `import sloika.module_tools` inside the script resolves to
    `sloika_b200.module_tools` (same trick as `bin/train_network.py:253-268` importing the module)."""
    import sloika_b200
    sloika_b200.install_as_sloika()
    scope = {'__name__': '__sloika_model__', '__file__': path}
    with open(path, 'r') as fh:
        code = compile(fh.read(), path, 'exec')
    exec(code, scope)
    return scope['network'](**kwargs)


def weights_of(layer, prefix=''):
    """Flat `{dotted.name: ndarray}` view of a layer tree's parameters (for fixtures/tests)."""
    out = {}
    if isinstance(layer, layers._Container):
        for i, child in enumerate(layer._children()):
            out.update(weights_of(child, '{}{}.'.format(prefix, i)))
    else:
        for key in layer._ordered():
            out[prefix + key] = getattr(layer, key).get_value()
    return out
