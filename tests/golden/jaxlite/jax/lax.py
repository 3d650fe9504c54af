"""`jax.lax` facade of jaxlite."""
from ._impl import (neg_p, abs_p, sqrt_p, exp_p, log_p, log1p_p, sin_p, cos_p, tan_p, asin_p, acos_p, atan_p, sinh_p,
                    cosh_p, tanh_p, asinh_p, acosh_p, atanh_p, square_p, logistic_p, erf_p, stop_gradient_p,
                    device_put_p, add_p, sub_p, mul_p, div_p, atan2_p, pow_p, max_p, min_p, eq_p, gt_p, lt_p,
                    integer_pow_p, reduce_sum_p, reduce_max_p, reduce_min_p, transpose_p, reshape_p, squeeze_p,
                    convert_element_type_p, iota_p, slice_p, concatenate_p, broadcast_in_dim_p, dot_general_p,
                    select_n_p, dot_general, stop_gradient, logistic, broadcast_in_dim, convert_element_type,
                    slice_in_dim, scatter, ScatterDimensionNumbers, multiply as mul, add, subtract as sub,
                    divide as div, log, exp, arctan as atan, tanh, sqrt, transpose, reshape, squeeze, expand_dims,
                    concatenate)
from ._impl import lax_slice as slice   # noqa: A001

stop_grad = stop_gradient
