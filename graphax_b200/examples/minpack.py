"""Test problems of the MINPACK-2 collection as traced scalar functions (the reference's ``examples/minpack.py``):
residuals of nonlinear systems whose Jacobians a Newton solver needs, and one discretised variational problem.

Operations are performed in the reference's order, because the vertex ids of a custom elimination order are the
positions of the traced operations (checked equation by equation against the jaxprs traced from the reference's own
sources, ``tests/test_reference_goldens.py::test_further_examples_trace_like_the_reference``).
"""
import numpy as np

from ..tracer import jnp

# ---- propane combustion in air, full formulation: 11 equilibrium equations in 11 unknowns ----
PRESSURE = 40.          # atmospheres
AIR_FUEL_RATIO = 10.    # relative amount of air


def PropaneCombustion(x1, x2, x3, x4, x5, x6, x7, x8, x9, x10, x11):
    """Residuals of the chemical-equilibrium system: x1..x10 are the moles of the ten combustion products, x11 their
    sum; four element balances, six equilibrium relations and the closure."""
    root_p = jnp.sqrt(PRESSURE / x11)
    carbon = x1 + x4 - 3.
    oxygen = 2 * x1 + x2 + x4 + x7 + x8 + x9 + 2 * x10 - AIR_FUEL_RATIO
    hydrogen = 2 * x2 + 2 * x5 + x6 + x7 - 8.
    nitrogen = 2 * x3 + x9 - 4 * AIR_FUEL_RATIO
    eq5 = x2 * x4 - x1 * x5
    eq6 = jnp.sqrt(x2 * x4) - jnp.sqrt(x1) * root_p * x6
    eq7 = jnp.sqrt(x1 * x2) - jnp.sqrt(x4) * x7 * root_p
    eq8 = x1 - x4 * x8 * PRESSURE / x11
    eq9 = x1 * jnp.sqrt(x3) - x4 * x9 * root_p
    eq10 = x1 ** 2 - x4 ** 2 * x10 * PRESSURE / x11
    closure = x11 - x1 - x2 - x3 - x4 - x5 - x6 - x7 - x8 - x9 - x10
    return carbon, oxygen, hydrogen, nitrogen, eq5, eq6, eq7, eq8, eq9, eq10, closure


# ---- human heart dipole: moments of two dipoles from body-surface measurements (all measured sums set to 1) ----
SUM_MX = SUM_MY = SUM_A = SUM_B = SUM_C = SUM_D = SUM_E = 1.


def HumanHeartDipole(x1, x2, x3, x4, x5, x6, x7, x8):
    """Residuals of the eight moment equations: (x1, x3) and (x2, x4) are the dipole strengths, (x5, x7) and (x6, x8)
    their positions; the equations are the real and imaginary parts of the first three complex moments."""
    m0_re = x1 + x2 - SUM_MX
    m0_im = x3 + x4 - SUM_MY
    m1_re = x5 * x1 + x6 * x2 - x7 * x3 - x8 * x4 - SUM_A
    m1_im = x7 * x1 + x8 * x2 + x5 * x3 + x6 * x4 - SUM_B

    sq_a = (x5 ** 2 - x7 ** 2)          # Re z^2 of the two positions
    sq_b = (x6 ** 2 - x8 ** 2)
    m2_re = x1 * sq_a - 2 * x3 * x5 * x7 + x2 * sq_b - 2 * x4 * x6 * x8 - SUM_C
    m2_im = x3 * sq_a - 2 * x1 * x5 * x7 + x4 * sq_b - 2 * x2 * x6 * x8 - SUM_D

    cube_a_re = x5 * (x5 ** 2 - 3 * x7 ** 2)     # Re / -Im z^3
    cube_a_im = x7 * (x7 ** 2 - 3 * x5 ** 2)
    cube_b_re = x6 * (x6 ** 2 - 3 * x8 ** 2)
    cube_b_im = x8 * (x8 ** 2 - 3 * x6 ** 2)
    m3_re = x1 * cube_a_re + x3 * cube_a_im + x2 * cube_b_re + x4 * cube_b_im - SUM_E
    m3_im = x3 * cube_a_re - x1 * cube_a_im + x4 * cube_b_re - x2 * cube_b_im - SUM_E
    return m0_re, m0_im, m1_re, m1_im, m2_re, m2_im, m3_re, m3_im


# ---- steady-state combustion (Bratu problem) on a 2 x 2 periodic grid: the discretised energy functional ----
HX, HY = .5, .5         # grid spacings
BRATU_LAMBDA = 6.81


def _mean_exp(centre, first, second):
    return 2. / 3. * (jnp.exp(centre) + jnp.exp(first) + jnp.exp(second))


def _triangle_energy(centre, first, second):
    """Contribution of one triangle: squared gradient of the piecewise-linear interpolant minus the source term."""
    return HX * HY / 4 * (((first - centre) / HX) ** 2 + ((second - centre) / HY) ** 2
                          - BRATU_LAMBDA * _mean_exp(centre, first, second))


def SteadyStateCombustion(v00, v01, v10, v11):
    """Energy of the four grid values: every node contributes its lower-left and its upper-right triangle."""
    e00 = _triangle_energy(v00, v10, v01) + _triangle_energy(v00, v01, v10)
    e01 = _triangle_energy(v01, v11, v00) + _triangle_energy(v01, v00, v11)
    e10 = _triangle_energy(v10, v00, v11) + _triangle_energy(v10, v11, v00)
    e11 = _triangle_energy(v11, v01, v10) + _triangle_energy(v11, v10, v01)
    return e00 + e01 + e10 + e11


# ---- Gaussian data fitting: 65 residuals of a sum of an exponential and three Gaussians ----
N_DATA = 65
FIT_TIMES = np.arange(N_DATA, dtype=np.float32) / 10
# the reference draws its data from jax.random.uniform(PRNGKey(123)); without JAX the same role is played by a fixed
# low-discrepancy sequence in [0, 1) (the values only shift the residuals, not the Jacobian)
FIT_DATA = ((np.arange(1, N_DATA + 1, dtype=np.float64) * 0.6180339887498949) % 1.0).astype(np.float32)


def _residuals(x1, x2, x3, x4, x5, x6, x7, x8, x9, x10, x11, y, t):
    out = y - x1 * jnp.exp(-t * x5)
    out -= x2 * jnp.exp(-(t - x9) ** 2 * x6)
    out -= x3 * jnp.exp(-(t - x10) ** 2 * x7)
    out -= x4 * jnp.exp(-(t - x11) ** 2 * x8)
    return out


def GaussianDataFitting(x1, x2, x3, x4, x5, x6, x7, x8, x9, x10, x11):
    """Residual vector (shape (65,)) of the model ``x1 exp(-t x5) + sum_k x_{k+1} exp(-(t - x_{k+8})^2 x_{k+5})``."""
    return _residuals(x1, x2, x3, x4, x5, x6, x7, x8, x9, x10, x11, jnp.asarray(FIT_DATA), jnp.asarray(FIT_TIMES))
