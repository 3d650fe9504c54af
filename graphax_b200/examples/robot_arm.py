"""Forward kinematics of a 6-DOF manipulator (position + Tait-Bryan angles).

Restates the reference workload ``src/graphax/examples/differential_kinematics.py:9-55``
(a deep chain of scalar sin/cos products; 114 equations, outputs at vertices
85, 102, 114, 62, 66, 68 - SURVEY.md Appendix A.6) against ``graphax_b200.jnp``.
"""
from ..tracer import jnp


def _sin_of_sum(a, b):
    return jnp.cos(a) * jnp.sin(b) + jnp.sin(a) * jnp.cos(b)


def _cos_of_sum(a, b):
    return jnp.cos(a) * jnp.cos(b) - jnp.sin(a) * jnp.sin(b)


# link lengths / offsets in mm
_A1, _A2, _A3 = 175., 890., 50.
_D1, _D4, _D6 = 575., 1035., 185.


def RobotArm_6DOF(t1, t2, t3, t4, t5, t6):
    c1 = jnp.cos(t1)
    c2 = jnp.cos(t2)
    c4 = jnp.cos(t4)
    c5 = jnp.cos(t5)
    c6 = jnp.cos(t6)

    c23 = _cos_of_sum(t2, t3)

    s1 = jnp.sin(t1)
    s2 = jnp.sin(t2)
    s4 = jnp.sin(t4)
    s5 = jnp.sin(t5)
    s6 = jnp.sin(t6)

    s23 = _sin_of_sum(t2, t3)

    # approach vector and two rotation-matrix entries
    ax = s5 * (c1 * c23 * c4 + s1 * s4) + c1 * s23 * c5
    ay = s5 * (s1 * c23 * c4 - c1 * s4) + s1 * s23 * c5
    az = s23 * c4 * s5 - c23 * c5

    nz = c6 * (c23 * s5 + s23 * c4 * c5) - s23 * s4 * s6
    oz = -s6 * (c23 * s5 + s23 * c4 * c5) - s23 * s4 * c6

    # Tait-Bryan angles
    z = jnp.arctan2(ay, ax)
    y_ = jnp.arctan2(jnp.sqrt((1. - az**2)), az)
    z__ = jnp.arctan2(oz, -nz)

    x1 = _D6 * (s5 * (c1 * c23 * c4 + s1 * s4) + c1 * s23 * c5)
    x2 = c1 * (_A1 + _A2 * c2 + _A3 * c23 + _D4 * s23)
    px = x1 + x2

    y1 = _D6 * (s5 * (s1 * c23 * c4 + c1 * s4) + s1 * s23 * c5)
    y2 = s1 * (_A1 + _A2 * c2 + _A3 * c23 + _D4 * s23)
    py = y1 + y2

    pz = _A2 * s2 + _D1 + _A3 * s23 - _D4 * c23 + _D6 * (s23 * c4 * s5 - c23 * c5)
    return px, py, pz, z, y_, z__
