"""graphax_b200 - B200-native engine for graphax's ``jacve`` hot path.

Same call shape as the reference (``graphax/__init__.py:2-4``)::

    from graphax_b200 import jacve, jnp
    jac = jacve(f, order="rev", argnums=(0, 1))(x, y)
"""
from .api import (jacve, vmap, tree_allclose, JacveFunction, SparseJacobian, sparse_tensor_zeros_like,
                  get_output_vertices, get_valid_vertices)
from .tracer import jnn, jnp, lax, make_jaxpr
from .primitives import Primitive, defelemental, defelemental2, elemental_rules
from .filtering import filter_jacve

__version__ = "0.1.0"
__all__ = ["jacve", "vmap", "tree_allclose", "jnn", "jnp", "lax", "make_jaxpr", "JacveFunction", "SparseJacobian",
           "Primitive", "defelemental", "defelemental2", "elemental_rules", "filter_jacve", "sparse_tensor_zeros_like",
           "get_output_vertices", "get_valid_vertices"]
