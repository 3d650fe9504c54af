"""Forward kinematics of a six-axis manipulator: tool position and Tait-Bryan orientation angles.

The workload behind BASELINE.json config 4 (the reference's ``RobotArm_6DOF``; geometry from
https://dergipark.org.tr/en/download/article-file/2289226): a deep chain of scalar sine / cosine products with no
vector vertices at all, which makes it the stress case for fusion and launch overhead - 114 equations, every edge
Jacobian a scalar, six outputs.

Vertex numbering (it is what an elimination order refers to, so the evaluation order below is part of the
contract; tests/test_reference_goldens.py compares every equation with the reference's own trace):

    1-5    cos of joints 1, 2, 4, 5, 6          6-12   cos(t2 + t3) by the addition theorem
    13-17  sin of joints 1, 2, 4, 5, 6          18-24  sin(t2 + t3)
    25-61  approach vector (ax, ay, az) and the two rotation-matrix entries nz, oz
    62-68  the three angles: atan2(ay, ax) = 62, atan2(sqrt(1 - az^2), az) = 66, atan2(oz, -nz) = 68
    69-114 position: px = 85, py = 102, pz = 114 (the wrist terms are evaluated again - there is no common
           subexpression elimination in a jaxpr, and the vertex ids count on that)
"""
from ..tracer import jnp

# link lengths and joint offsets in millimetres
LINK_1, LINK_2, LINK_3 = 175., 890., 50.
OFFSET_1, OFFSET_4, OFFSET_6 = 575., 1035., 185.


def _cos_sum(p, q):
    """cos(p + q) = cos p cos q - sin p sin q (seven equations)."""
    return jnp.cos(p) * jnp.cos(q) - jnp.sin(p) * jnp.sin(q)


def _sin_sum(p, q):
    """sin(p + q) = cos p sin q + sin p cos q (seven equations)."""
    return jnp.cos(p) * jnp.sin(q) + jnp.sin(p) * jnp.cos(q)


def RobotArm_6DOF(t1, t2, t3, t4, t5, t6):
    """Joint angles (six scalars) -> (px, py, pz, yaw, pitch, roll)."""
    cos1, cos2, cos4, cos5, cos6 = jnp.cos(t1), jnp.cos(t2), jnp.cos(t4), jnp.cos(t5), jnp.cos(t6)
    cos23 = _cos_sum(t2, t3)
    sin1, sin2, sin4, sin5, sin6 = jnp.sin(t1), jnp.sin(t2), jnp.sin(t4), jnp.sin(t5), jnp.sin(t6)
    sin23 = _sin_sum(t2, t3)

    # every helper below is *re-evaluated* at each use, on purpose (see the module docstring)
    def wrist_along_x():
        return sin5 * (cos1 * cos23 * cos4 + sin1 * sin4) + cos1 * sin23 * cos5

    def wrist_along_z():
        return sin23 * cos4 * sin5 - cos23 * cos5

    def elbow_mix():
        return cos23 * sin5 + sin23 * cos4 * cos5

    def planar_reach():
        return LINK_1 + LINK_2 * cos2 + LINK_3 * cos23 + OFFSET_4 * sin23

    approach_x = wrist_along_x()
    approach_y = sin5 * (sin1 * cos23 * cos4 - cos1 * sin4) + sin1 * sin23 * cos5
    approach_z = wrist_along_z()

    normal_z = cos6 * elbow_mix() - sin23 * sin4 * sin6
    orient_z = -sin6 * elbow_mix() - sin23 * sin4 * cos6

    yaw = jnp.arctan2(approach_y, approach_x)
    pitch = jnp.arctan2(jnp.sqrt((1. - approach_z**2)), approach_z)
    roll = jnp.arctan2(orient_z, -normal_z)

    px = OFFSET_6 * wrist_along_x() + cos1 * planar_reach()
    # note the '+' where approach_y has a '-': the published position formula, kept as the reference has it
    py = OFFSET_6 * (sin5 * (sin1 * cos23 * cos4 + cos1 * sin4) + sin1 * sin23 * cos5) + sin1 * planar_reach()
    pz = LINK_2 * sin2 + OFFSET_1 + LINK_3 * sin23 - OFFSET_4 * cos23 + OFFSET_6 * wrist_along_z()
    return px, py, pz, yaw, pitch, roll
