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
