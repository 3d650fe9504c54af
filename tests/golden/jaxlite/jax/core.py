"""`jax.core` facade of jaxlite."""
from ._impl import ShapedArray, Var, Literal, JaxprEqn, Jaxpr, ClosedJaxpr, Primitive, Tracer
