"""Small scalar functions used for plumbing checks.

``readme_f`` stands in for the ``f`` of the reference README call
``graphax.jacve(f, order=[1, 3, 2, 4], argnums=(0, 1, 2, 3))(1., 1., 2., 7.)``
(``README.rst:41-42``), which the reference never defines: four inputs, four
eliminable vertices (1-4) and one output vertex (5).
"""
from ..tracer import jnp


def readme_f(x, y, z, w):
    a = x * y
    b = jnp.sin(a)
    c = a * z
    d = b + c
    return d * w


def Simple(x, y):
    """Reference ``examples/easy.py`` style two-input scalar function."""
    z = x * y
    w = jnp.sin(z)
    return w + z, jnp.log(w)


def Lighthouse(nu, gamma, omega, t):
    """The lighthouse example of Griewank & Walther (two outputs)."""
    y1 = nu * jnp.tan(omega * t) / (gamma - jnp.tan(omega * t))
    y2 = gamma * y1
    return y1, y2


def Hole(x, y, z, w):
    a = y * z
    b = a + x
    c = a + w
    d = jnp.cos(b)
    e = jnp.exp(c)
    return d - e, d / e, d * e


def Helmholtz(x):
    """Helmholtz free energy density of a mixture, x of shape (n,): a rank-0 intermediate broadcast
    back into a vector (reference ``examples/easy.py``, tested with order [2, 5, 4, 3, 1])."""
    return x * jnp.log(x / (1. + -jnp.sum(x)))
