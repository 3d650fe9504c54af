"""NumPy stand-in for the slice of JAX that graphax uses — see `_impl.py`.  Test infrastructure only."""
from . import _impl
from ._impl import make_jaxpr, jit, vmap, Array, custom_jvp, ShapeDtypeStruct
from ._autodiff import jacfwd, jacrev, grad
from . import numpy, lax, tree_util, core, nn, random   # noqa: F401
from . import _src                                      # noqa: F401
from .scipy import special as _special                   # noqa: F401

__version__ = "0.0-jaxlite"
