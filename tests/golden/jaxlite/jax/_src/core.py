from .._impl import ShapedArray, Var, Literal, JaxprEqn, Jaxpr, ClosedJaxpr, Primitive, Tracer  # noqa: F401
