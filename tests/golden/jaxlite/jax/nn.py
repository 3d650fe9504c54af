"""`jax.nn` facade of jaxlite."""
from ._impl import logistic as sigmoid, tanh, _unavailable, exp, stop_gradient, max as _max, sum as _sum, subtract, divide


def softmax(x, axis=-1):
    """`jax.nn.softmax` without its `custom_jvp` wrapper (graphax has no rule for `custom_jvp_call`): the body JAX
    differentiates when `jax_softmax_custom_jvp` is off, `exp(x - stop_gradient(max(x))) / sum(...)`."""
    x_max = _max(x, axis=axis, keepdims=True)
    unnormalized = exp(subtract(x, stop_gradient(x_max)))
    return divide(unnormalized, _sum(unnormalized, axis=axis, keepdims=True))


relu = _unavailable("nn.relu")
