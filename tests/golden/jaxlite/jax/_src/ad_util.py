from .._impl import add_any_p, stop_gradient_p  # noqa: F401
