"""Namespace that `models/*.py` import (`sloika/module_tools.py:1-13`): lets the reference's model
factories (`network(klen, sd, ...)`) build B200 layers unchanged."""
from functools import partial

import numpy as np
from scipy.stats import truncnorm

from sloika_b200.config import sloika_dtype
from sloika_b200.activation import *
from sloika_b200.layers import *
from sloika_b200.variables import *


def truncated_normal(size, sd):
    ''' Truncated normal for Xavier style initiation (`module_tools.py:9-13`)
    '''
    res = sd * truncnorm.rvs(-2, 2, size=size)
    return res.astype(sloika_dtype)
