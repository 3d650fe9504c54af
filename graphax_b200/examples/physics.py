"""Scalar test functions from physics: a charged rotating black-hole metric, a mixture free energy and one step of
a warm-rain cloud microphysics scheme (the reference's ``examples/easy.py`` beyond Simple / Lighthouse / Hole /
Helmholtz).

The vertex ids of a custom elimination order are the positions of the traced operations, so every function performs
its operations in the reference's order (``tests/test_reference_goldens.py::test_further_examples_trace_like_the_
reference`` compares the traces equation by equation with the jaxprs the reference's own sources produced).
"""
from ..tracer import jnp
from .simple import Helmholtz

# Kerr-Sen black hole: rotation parameter, charge parameter and mass (geometric units)
SPIN = .5
CHARGE = .9
MASS = 1.


def KerrSenn_metric(t, r, theta, phi):
    """Non-zero components (g_tt, g_rr, g_theta_theta, g_phi_phi, g_phi_t) of the Kerr-Sen metric in Boyer-Lindquist
    coordinates; stationary and axisymmetric, so ``t`` and ``phi`` do not enter."""
    sin2 = jnp.sin(theta) ** 2
    sigma = r ** 2 + 2. * CHARGE * r + SPIN ** 2 * jnp.cos(theta) ** 2
    delta = r ** 2 + 2 * CHARGE * r - 2. * MASS * r + SPIN ** 2
    g_tt = -(1. - 2. * MASS * r / sigma)
    g_rr = sigma / delta
    g_thth = sigma
    g_phph = (sigma + SPIN ** 2 * sin2 + 2. * MASS * r * SPIN ** 2 * sin2 / sigma) * sin2
    g_pht = -2. * MASS * r * SPIN / sin2
    return g_tt, g_rr, g_thth, g_phph, g_pht


def KerrSenn_Jacobian(t, r, theta, phi):
    """First derivatives of the metric components (the ingredients of the Christoffel symbols) as one traced
    function: differentiate it again with ``jacve`` for second derivatives."""
    from ..api import jacve
    return jacve(KerrSenn_metric, order="fwd", argnums=(0, 1, 2, 3))(t, r, theta, phi)


def FreeEnergy(x):
    """Total Helmholtz free energy of the mixture ``x`` (shape (n,)): a scalar function of a vector."""
    return jnp.sum(Helmholtz(x))


# state and constants of the cloud scheme (mixing ratios of cloud water, rain and vapour; saturation terms)
CLOUD_WATER, RAIN, VAPOUR = 1., 1., 1.
COND_RATE = 1.
SUPERSATURATION = 1.
RAIN_SOURCE = 0.


def _condensation(qc):
    return COND_RATE * SUPERSATURATION * qc ** 0.33333


def _accretion(a2, bc, br, qc, qr):
    return a2 * qc ** bc * qr ** br


def _autoconversion(a1, gamma, qc):
    return a1 * qc ** gamma


def _evaporation(e1, d1, e2, d2, qr):
    return e1 * qr ** d1 + e2 * qr ** d2


def CloudSchemes_step(a1, a2, e1, e2, delta, gamma, bc, br, d1, d2, chi):
    """Tendencies (d qc, d qr, d qv) of a Kessler-type warm-rain scheme as a function of its eleven tunable
    parameters, at the fixed state above (the parameter-sensitivity study of gmd-2019-140)."""
    qc, qr = CLOUD_WATER, RAIN
    dqc = _condensation(qc) - _autoconversion(a1, gamma, qc) - _accretion(a2, bc, br, qc, qr)
    dqr = _autoconversion(a1, gamma, qc) + _evaporation(e1, d1, e2, d2, qr) + RAIN_SOURCE - delta * qr ** chi
    dqv = -_condensation(qc) - _evaporation(e1, d1, e2, d2, qr)
    return dqc, dqr, dqv
