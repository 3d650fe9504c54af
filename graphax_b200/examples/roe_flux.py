"""Roe's approximate Riemann flux (Roe 1981) across one cell face for the Euler equations, in one and three dimensions.

These functions are the *inputs* of the hot path - the headline benchmark accumulates the Jacobian of
``RoeFlux_3d`` for a batch of left / right states - and correspond to the reference's workloads of the same names
(its adiabatic constant is an unphysical 0.5, kept here because the test points are only in-domain for it).

A state is (density, momentum, total energy); the flux is the average of the two physical fluxes minus a
dissipation term assembled from the eigen-decomposition of the Roe-averaged flux Jacobian:

    Roe averages   velocity and enthalpy weighted by sqrt(density) of the two sides
    sound speed    a^2 = (gamma - 1) (h - q^2 / 2)
    wave strengths projections of the jump (d rho, d mom, d E) onto the eigenvectors, times |eigenvalue|
    dissipation    sum over waves of strength x eigenvector

The sequence of array operations is part of the contract: vertex i of an elimination order is the i-th operation
(101 for the 1-D flux with outputs 97 / 99 / 101, 138 for the 3-D flux with outputs 134 / 136 / 138; repeated
expressions such as sqrt(density) are deliberately evaluated again every time they are needed, a jaxpr has no common
subexpression elimination).  tests/test_reference_goldens.py compares every equation, operand and literal with the
reference's own trace.
"""
from ..tracer import jnp

GAMMA = .5
_GM1 = GAMMA - 1.


def _specific_enthalpy(density, energy, pressure):
    return (energy + pressure) / density


# ---- one dimension: three scalars per side --------------------------------------------------------------------------
def _pressure_scalar(density, momentum, energy):
    return _GM1 * (energy - .5 * momentum**2 / density)


def _physical_flux_scalar(density, momentum, energy, pressure):
    speed = momentum / density
    return momentum, pressure + momentum * speed, speed * (pressure + energy)


def RoeFlux_1d(ul0, ul1, ul2, ur0, ur1, ur2):
    """(density, momentum, energy) left, then right -> the three flux components."""
    jump_density = ur0 - ul0
    mean_density = jnp.sqrt(ul0 * ur0)
    weight_sum = jnp.sqrt(ul0) + jnp.sqrt(ur0)

    speed_left = ul1 / ul0
    pressure_left = _pressure_scalar(ul0, ul1, ul2)
    enthalpy_left = _specific_enthalpy(ul0, ul2, pressure_left)

    speed_right = ur1 / ur0
    pressure_right = _pressure_scalar(ur0, ur1, ur2)
    enthalpy_right = _specific_enthalpy(ur0, ur2, pressure_right)

    jump_pressure = pressure_right - pressure_left
    jump_speed = speed_right - speed_left

    speed = (jnp.sqrt(ul0) * speed_left + jnp.sqrt(ur0) * speed_right) / weight_sum
    enthalpy = (jnp.sqrt(ul0) * enthalpy_left + jnp.sqrt(ur0) * enthalpy_right) / weight_sum

    kinetic = speed**2
    sound2 = _GM1 * (enthalpy - .5 * kinetic)
    sound = jnp.sqrt(sound2)

    fast, entropy_wave, slow = jnp.abs(speed + sound), jnp.abs(speed), jnp.abs(speed)

    impedance = sound * mean_density
    strength_entropy = (jump_density - jump_pressure / sound2) * entropy_wave
    strength_fast = (jump_speed + jump_pressure / impedance) * fast
    strength_slow = (jump_speed - jump_pressure / impedance) * slow

    left = _physical_flux_scalar(ul0, ul1, ul2, pressure_left)
    right = _physical_flux_scalar(ur0, ur1, ur2, pressure_right)
    total_mass, total_momentum, total_energy = left[0] + right[0], left[1] + right[1], left[2] + right[2]

    half_ratio = .5 * mean_density / sound

    damp_mass = strength_entropy + half_ratio * strength_fast - half_ratio * strength_slow
    damp_momentum = (strength_entropy * speed + half_ratio * strength_fast * (speed + sound)
                     - half_ratio * strength_slow * (speed - sound))
    damp_energy = (strength_entropy * .5 * kinetic + half_ratio * strength_fast * (enthalpy + speed * sound)
                   - half_ratio * strength_slow * (enthalpy - speed * sound))

    return .5 * (total_mass - damp_mass), .5 * (total_momentum - damp_momentum), .5 * (total_energy - damp_energy)


# ---- three dimensions: density (1,), momentum (3,), energy (1,) per side --------------------------------------------
def _pressure_vector(density, momentum, energy):
    return _GM1 * (energy - .5 * jnp.sum(momentum**2) / density)


def _physical_flux_vector(density, momentum, velocity, energy, pressure):
    normal_speed = velocity[0:1]
    pressure_on_normal = jnp.concatenate([pressure, jnp.zeros(2, dtype=jnp.float32)])
    return momentum[0:1], pressure_on_normal + momentum * velocity, normal_speed * (pressure + energy)


def RoeFlux_3d(ul0, ul, ul4, ur0, ur, ur4):
    """Left state (density, momentum[3], energy), then the right one -> flux (mass, momentum[3], energy)."""
    jump_density = ur0 - ul0
    jump_momentum = ur - ul
    jump_m1, jump_m2, jump_m3 = jump_momentum[0:1], jump_momentum[1:2], jump_momentum[2:]
    jump_energy = ur4 - ul4

    velocity_left = ul / ul0
    velocity_right = ur / ur0
    weight_sum = jnp.sqrt(ul0) + jnp.sqrt(ur0)

    velocity = (jnp.sqrt(ul0) * velocity_left + jnp.sqrt(ur0) * velocity_right) / weight_sum
    vel_n, vel_t1, vel_t2 = velocity[0:1], velocity[1:2], velocity[2:]

    pressure_left = _pressure_vector(ul0, ul, ul4)
    enthalpy_left = _specific_enthalpy(ul0, ul4, pressure_left)
    pressure_right = _pressure_vector(ur0, ur, ur4)
    enthalpy_right = _specific_enthalpy(ur0, ur4, pressure_right)

    enthalpy = (jnp.sqrt(ul0) * enthalpy_left + jnp.sqrt(ur0) * enthalpy_right) / weight_sum

    kinetic2 = jnp.sum(velocity**2)
    sound2 = _GM1 * (enthalpy - .5 * kinetic2)
    sound = jnp.sqrt(sound2)

    fast = vel_n + sound
    convected = vel_n
    slow = vel_n - sound

    # strengths of the entropy wave (s_e), the two shear waves (s_t2, s_t1) and the acoustic pair (s_slow, s_fast)
    s_e = (_GM1 / sound2 * ((enthalpy - kinetic2) * jump_density + jnp.dot(velocity, jump_momentum) - jump_energy)) * convected
    s_t2 = (jump_m3 / vel_t2 - jump_m1) * convected
    s_t1 = (jump_m2 / vel_t1 - jump_density) * convected
    rest = jump_density - s_e
    acoustic = (jump_m1 - vel_n * jump_density) / sound
    s_slow = .5 * (rest - acoustic) * slow
    s_fast = .5 * (rest + acoustic) * fast

    left = _physical_flux_vector(ul0, ul, velocity_left, ul4, pressure_left)
    right = _physical_flux_vector(ur0, ur, velocity_right, ur4, pressure_right)
    total_mass, total_momentum, total_energy = left[0] + right[0], left[1] + right[1], left[2] + right[2]

    damp_mass = s_slow + s_e + s_fast * fast
    damp_normal = s_slow * slow + s_e * vel_n + s_fast * fast
    damp_t1 = s_slow * vel_t1 + s_t1 * vel_t1 + s_e * vel_t1 + s_fast * vel_t1
    damp_t2 = s_slow * vel_t2 + s_t2 * vel_t2 + s_e * vel_t2 + s_fast * vel_t2
    damp_energy = (s_slow * (enthalpy - vel_n * sound) + s_t1 * vel_t1**2 + s_t2 * vel_t2**2 + s_e * .5 * kinetic2
                   + s_fast * (enthalpy + vel_n * sound))

    damp_momentum = jnp.concatenate([damp_normal, damp_t1, damp_t2])

    return .5 * (total_mass - damp_mass), .5 * (total_momentum - damp_momentum), .5 * (total_energy - damp_energy)
