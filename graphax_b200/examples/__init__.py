"""The reference's example functions (``graphax/examples/*.py``), written against ``graphax_b200.jnp``.

Every function performs its operations in the reference's order, so the vertex ids of a custom elimination order mean
the same vertices (checked equation by equation against jaxprs traced from the reference's own sources).

=====================  ==========================================  ===================================================
module                 functions                                   reference
=====================  ==========================================  ===================================================
``roe_flux``           RoeFlux_1d, RoeFlux_3d                      ``roe.py`` (configs 2 and 3 of the baseline)
``robot_arm``          RobotArm_6DOF                               ``differential_kinematics.py`` (config 4)
``simple``             Simple, Lighthouse, Hole, Helmholtz,        ``easy.py``; ``readme_f`` stands in for the README's
                       readme_f                                    undefined ``f`` (config 1)
``physics``            KerrSenn_metric (+ _Jacobian), FreeEnergy,  ``easy.py``
                       CloudSchemes_step
``minpack``            PropaneCombustion, HumanHeartDipole,        ``minpack.py`` (GaussianDataFitting closes over its
                       SteadyStateCombustion, GaussianDataFitting  65 data points: constants of the trace)
``economics``          BlackScholes, BlackScholes_Jacobian         ``economics.py`` (nested jacve: second-order greeks)
``deep_learning``      Perceptron, Encoder, EncoderDecoder         ``deep_learning.py`` (matrix-valued vertices)
``mlp``                MLP_loss                                    the dense weight-Jacobian workload (config 5)
``zoo``                ElementalZoo                                every elementwise rule of ``primitives.py``
=====================  ==========================================  ===================================================

Not here: ``neuromorphic.py`` (weight matrices of 10^4 elements with vector outputs: outside the per-sample engine and
the scalar-output tensor path) and ``randoms.py`` (machine-generated functions; the tests rebuild them from the jaxprs
the reference traced).
"""
from .roe_flux import RoeFlux_1d, RoeFlux_3d
from .robot_arm import RobotArm_6DOF
from .simple import readme_f, Simple, Lighthouse, Hole, Helmholtz
from .zoo import ElementalZoo
from .mlp import MLP_loss
from .deep_learning import Perceptron, Encoder, EncoderDecoder
from .economics import BlackScholes, BlackScholes_Jacobian
from .physics import KerrSenn_metric, KerrSenn_Jacobian, FreeEnergy, CloudSchemes_step
from .minpack import PropaneCombustion, HumanHeartDipole, SteadyStateCombustion, GaussianDataFitting

__all__ = ["RoeFlux_1d", "RoeFlux_3d", "RobotArm_6DOF", "readme_f", "Simple", "Lighthouse", "Hole", "Helmholtz",
           "ElementalZoo", "MLP_loss", "Perceptron", "Encoder", "EncoderDecoder", "BlackScholes", "BlackScholes_Jacobian",
           "KerrSenn_metric", "KerrSenn_Jacobian", "FreeEnergy", "CloudSchemes_step", "PropaneCombustion", "HumanHeartDipole",
           "SteadyStateCombustion", "GaussianDataFitting"]
