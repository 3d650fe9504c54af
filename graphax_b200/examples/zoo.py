"""A function that touches every elementwise elemental rule the engine lowers (reference
``primitives.py:150-239``), on mixed rank-0 / rank-1 operands.  Not a reference workload: a coverage
fixture for the rule table (the reference tests these rules one primitive at a time in
``tests/core/primitive_test.py``)."""
from ..tracer import jnp


def ElementalZoo(x, y, v):
    """x, y: scalars with 0 < x < 1 < y; v: shape (3,)."""
    a = jnp.exp(x) * jnp.log(y)
    b = jnp.tanh(v) + jnp.sigmoid(x)
    c = jnp.maximum(v, y) - jnp.minimum(v, x)
    d = jnp.arctan(a) + jnp.sinh(x) * jnp.cosh(y)
    e = jnp.arcsin(x * .5) + jnp.arccos(y * .25) + jnp.tan(x)
    f = jnp.log1p(jnp.square(b)) + jnp.erf(c)
    g = jnp.power(y, x) + jnp.square(v) + jnp.arcsinh(v) + jnp.arctanh(x * .3) + jnp.arccosh(y + 1.)
    h = jnp.abs(v - 1.5) + jnp.sqrt(y) + (-v) ** 3 / y
    return d + e, f * g + h, jnp.sum(c) * a
