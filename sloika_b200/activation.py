"""Activation tokens for the B200 path.

The reference's `sloika/activation.py` holds Theano expressions; model pickles refer to them *by
name* (`sloika.activation.elu` etc.), and the layers only ever call `fun(x)` on symbolic tensors.
Here an activation is a named token whose `code` selects the fused epilogue inside the CUDA
kernels (see `include/sloika_b200.h`, `SLOIKA_ACT_*`).  Only the four activations on the raw
basecall path are implemented on the device (`activation.py:8, 38-42, 52, 56`); the others are
declared so that pickles/models naming them fail with a clear message instead of an AttributeError.
"""

ACT_LINEAR, ACT_TANH, ACT_SIGMOID, ACT_ELU = 0, 1, 2, 3


class Activation(object):
    def __init__(self, name, code):
        self.__name__ = name
        self.code = code

    def __call__(self, x):
        raise TypeError("sloika_b200 activations are kernel-epilogue tokens; "
                        "'{}' cannot be applied to host data".format(self.__name__))

    def __repr__(self):
        return "<activation {}>".format(self.__name__)

    def __reduce__(self):
        # pickles by name, exactly like the reference's module-level functions
        return (_lookup, (self.__name__,))


def _lookup(name):
    return globals()[name]


linear = Activation('linear', ACT_LINEAR)
tanh = Activation('tanh', ACT_TANH)
sigmoid = Activation('sigmoid', ACT_SIGMOID)
elu = Activation('elu', ACT_ELU)

# Named by the reference but not on the raw basecall path: no kernel epilogue.
_UNSUPPORTED = ['relu', 'relu_smooth', 'softplus', 'exp', 'erf', 'L1mL2', 'fair', 'retu', 'tanh_pm',
                'sigmoid_pm', 'bounded_linear', 'sin', 'cauchy', 'geman_mcclure', 'welsh']
for _n in _UNSUPPORTED:
    globals()[_n] = Activation(_n, None)
del _n

__all__ = ['linear', 'tanh', 'sigmoid', 'elu'] + _UNSUPPORTED


def code_of(fun):
    """Kernel epilogue code for an activation token; raises for activations with no device epilogue."""
    code = getattr(fun, 'code', None)
    if code is None:
        raise NotImplementedError("activation '{}' has no sm_100a epilogue (only linear/tanh/sigmoid/elu "
                                  "are on the raw basecall path)".format(getattr(fun, '__name__', fun)))
    return code
