"""Roe approximate Riemann flux (Roe 1981) for the 1-D and 3-D Euler equations.

These are the *inputs* of the hot path: the functions whose Jacobians the named
configs accumulate.  They restate the maths of the reference workloads
(``src/graphax/examples/roe.py:18-70`` for 1-D, ``:85-148`` for 3-D; adiabatic
constant ``roe.py:5``) against ``graphax_b200.jnp`` and keep the same sequence
of array operations, because the vertex ids of an elimination order are the
positions of those operations.
"""
from ..tracer import jnp

GAMMA = .5  # sic: the reference uses an unphysical 0.5 (roe.py:5)


def _pressure_1d(rho, mom, ener):
    return (GAMMA - 1.) * (ener - .5 * mom**2 / rho)


def _enthalpy(rho, ener, p):
    return (ener + p) / rho


def _euler_flux_1d(rho, mom, ener, p):
    vel = mom / rho
    return mom, p + mom * vel, vel * (p + ener)


def RoeFlux_1d(ul0, ul1, ul2, ur0, ur1, ur2):
    """6 scalars (left/right density, momentum, energy) -> 3 scalar fluxes; 101 equations."""
    d_rho = ur0 - ul0
    rho_mean = jnp.sqrt(ul0 * ur0)
    wsum = jnp.sqrt(ul0) + jnp.sqrt(ur0)

    vel_l = ul1 / ul0
    p_l = _pressure_1d(ul0, ul1, ul2)
    h_l = _enthalpy(ul0, ul2, p_l)

    vel_r = ur1 / ur0
    p_r = _pressure_1d(ur0, ur1, ur2)
    h_r = _enthalpy(ur0, ur2, p_r)

    d_p = p_r - p_l
    d_vel = vel_r - vel_l

    # Roe averages
    u = (jnp.sqrt(ul0) * vel_l + jnp.sqrt(ur0) * vel_r) / wsum
    h = (jnp.sqrt(ul0) * h_l + jnp.sqrt(ur0) * h_r) / wsum

    q2 = u**2
    a2 = (GAMMA - 1.) * (h - .5 * q2)
    a = jnp.sqrt(a2)

    lam_p = jnp.abs(u + a)
    lam_0 = jnp.abs(u)
    lam_m = jnp.abs(u)

    # wave strengths
    n = a * rho_mean
    c0 = (d_rho - d_p / a2) * lam_0
    c1 = (d_vel + d_p / n) * lam_p
    c2 = (d_vel - d_p / n) * lam_m

    fl0, fl1, fl2 = _euler_flux_1d(ul0, ul1, ul2, p_l)
    fr0, fr1, fr2 = _euler_flux_1d(ur0, ur1, ur2, p_r)

    f0 = fl0 + fr0
    f1 = fl1 + fr1
    f2 = fl2 + fr2

    c = .5 * rho_mean / a

    df0 = c0 + c * c1 - c * c2
    df1 = c0 * u + c * c1 * (u + a) - c * c2 * (u - a)
    df2 = c0 * .5 * q2 + c * c1 * (h + u * a) - c * c2 * (h - u * a)

    return .5 * (f0 - df0), .5 * (f1 - df1), .5 * (f2 - df2)


def _pressure_3d(rho, mom, ener):
    return (GAMMA - 1.) * (ener - .5 * jnp.sum(mom**2) / rho)


def _euler_flux_3d(rho, mom, vel, ener, p):
    vel_x = vel[0:1]
    p_vec = jnp.concatenate([p, jnp.zeros(2, dtype=jnp.float32)])
    return mom[0:1], p_vec + mom * vel, vel_x * (p + ener)


def RoeFlux_3d(ul0, ul, ul4, ur0, ur, ur4):
    """(1,),(3,),(1,) left and right states -> fluxes (1,),(3,),(1,); 138 equations."""
    d_rho = ur0 - ul0
    d_mom = ur - ul
    d_m1 = d_mom[0:1]
    d_m2 = d_mom[1:2]
    d_m3 = d_mom[2:]
    d_ener = ur4 - ul4

    vel_l = ul / ul0
    vel_r = ur / ur0
    wsum = jnp.sqrt(ul0) + jnp.sqrt(ur0)

    # Roe averages
    uvw = (jnp.sqrt(ul0) * vel_l + jnp.sqrt(ur0) * vel_r) / wsum
    u = uvw[0:1]
    v = uvw[1:2]
    w = uvw[2:]

    p_l = _pressure_3d(ul0, ul, ul4)
    h_l = _enthalpy(ul0, ul4, p_l)

    p_r = _pressure_3d(ur0, ur, ur4)
    h_r = _enthalpy(ur0, ur4, p_r)

    h = (jnp.sqrt(ul0) * h_l + jnp.sqrt(ur0) * h_r) / wsum

    q2 = jnp.sum(uvw**2)
    a2 = (GAMMA - 1.) * (h - .5 * q2)
    a = jnp.sqrt(a2)

    lam_p = u + a
    lam_0 = u
    lam_m = u - a

    # wave strengths
    c3 = ((GAMMA - 1.) / a2 * ((h - q2) * d_rho + jnp.dot(uvw, d_mom) - d_ener)) * lam_0
    c2 = (d_m3 / w - d_m1) * lam_0
    c1 = (d_m2 / v - d_rho) * lam_0
    k1 = d_rho - c3
    k2 = (d_m1 - u * d_rho) / a
    c0 = .5 * (k1 - k2) * lam_m
    c4 = .5 * (k1 + k2) * lam_p

    fl0, fl, fl4 = _euler_flux_3d(ul0, ul, vel_l, ul4, p_l)
    fr0, fr, fr4 = _euler_flux_3d(ur0, ur, vel_r, ur4, p_r)

    f0 = fl0 + fr0
    f = fl + fr
    f4 = fl4 + fr4

    df0 = c0 + c3 + c4 * lam_p
    df1 = c0 * lam_m + c3 * u + c4 * lam_p
    df2 = c0 * v + c1 * v + c3 * v + c4 * v
    df3 = c0 * w + c2 * w + c3 * w + c4 * w
    df4 = c0 * (h - u * a) + c1 * v**2 + c2 * w**2 + c3 * .5 * q2 + c4 * (h + u * a)

    df = jnp.concatenate([df1, df2, df3])

    return .5 * (f0 - df0), .5 * (f - df), .5 * (f4 - df4)
