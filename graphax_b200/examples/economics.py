"""Black-Scholes price and its Jacobian-of-Jacobian (second-order greeks) - the reference's nested-jacve
example (/root/reference/src/graphax/examples/economics.py:8-22), restated against the jnp shim."""
from ..tracer import jnp


def Phi(z):
    return (1. + jnp.erf(z / jnp.sqrt(2.))) / 2


def Black76(F, K, r, sigma, T):
    d1 = (jnp.log(F / K) + (sigma ** 2) / 2. * T) / (sigma * jnp.sqrt(T))
    d2 = d1 - sigma * jnp.sqrt(T)
    return jnp.exp(-r * T) * (F * Phi(d1) - K * Phi(d2))


def BlackScholes(S, K, r, sigma, T):
    F = jnp.exp(r * T) * S
    return Black76(F, K, r, sigma, T)


def BlackScholes_Jacobian(S, K, r, sigma, T):
    """The first-order greeks as a traced function: differentiating it gives the second-order greeks."""
    from ..api import jacve
    return jacve(BlackScholes, order="rev", argnums=(0, 1, 2, 3, 4))(S, K, r, sigma, T)
