"""Global dtype of the path.

Mirrors `sloika/config.py:1-3` (`sloika_dtype = theano.config.floatX`); the basecall wrapper
(`bin/basecall_network:5-7`) pins floatX=float32, and every pickled weight is float32, so the
B200 path is float32 end to end.
"""
sloika_dtype = 'float32'
