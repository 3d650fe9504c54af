"""Option pricing: the Black-Scholes price of a European call through Black's 1976 formula, and its first-order
greeks as a traced function - differentiating that again with jacve gives the second-order greeks (the
reference's nested-jacve example of the same name, test point in tests/examples/example_test.py:351-363)."""
from ..tracer import jnp


def _normal_cdf(z):
    return (1. + jnp.erf(z / jnp.sqrt(2.))) / 2


def _black76(forward, strike, rate, vol, maturity):
    d_plus = (jnp.log(forward / strike) + (vol ** 2) / 2. * maturity) / (vol * jnp.sqrt(maturity))
    d_minus = d_plus - vol * jnp.sqrt(maturity)
    discount = jnp.exp(-rate * maturity)
    return discount * (forward * _normal_cdf(d_plus) - strike * _normal_cdf(d_minus))


def BlackScholes(S, K, r, sigma, T):
    """Call price for spot ``S``, strike ``K``, rate ``r``, volatility ``sigma`` and maturity ``T``."""
    return _black76(jnp.exp(r * T) * S, K, r, sigma, T)


def BlackScholes_Jacobian(S, K, r, sigma, T):
    """(delta, dual delta, rho, vega, theta-like) of the call as one traced function."""
    from ..api import jacve
    greeks = jacve(BlackScholes, order="rev", argnums=(0, 1, 2, 3, 4))
    return greeks(S, K, r, sigma, T)
