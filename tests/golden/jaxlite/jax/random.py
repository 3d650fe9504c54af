"""`jax.random` facade of jaxlite: NumPy generators behind the two calls the reference's examples make at import
time (`examples/minpack.py:89-91`); the streams are NOT JAX's."""
import numpy as _np
from ._impl import Array as _Array


def PRNGKey(seed):
    return int(seed)


key = PRNGKey


def split(key, num=2):
    return [key * 7919 + i + 1 for i in range(num)]


def uniform(key, shape=(), dtype=None, minval=0., maxval=1.):
    return _Array(_np.random.default_rng(key).uniform(minval, maxval, shape))


def normal(key, shape=(), dtype=None):
    return _Array(_np.random.default_rng(key).normal(size=shape))
