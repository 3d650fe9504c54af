"""`jax.numpy` facade of jaxlite."""
import numpy as _np
from ._impl import (add, subtract, multiply, divide, true_divide, arctan2, maximum, minimum, greater, less, equal,
                    power, negative, abs, absolute, sqrt, exp, log, log1p, sin, cos, tan, arcsin, arccos, arctan,
                    sinh, cosh, tanh, arcsinh, arccosh, arctanh, square, sum, max, min, amax, amin, mean, dot, matmul,
                    transpose, reshape, squeeze, expand_dims, concatenate, zeros, ones, zeros_like, ones_like, full,
                    split, array, asarray, eye, tile, diagonal, where, allclose, logical_and, arange, einsum, diag, all, Array as ndarray)

float32 = _np.float32
float64 = _np.float64
int32 = _np.int32
int64 = _np.int64
bool_ = _np.bool_
pi = _np.pi
newaxis = None
inf = _np.inf
