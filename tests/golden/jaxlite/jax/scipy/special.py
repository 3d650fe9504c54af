from .._impl import erf  # noqa: F401
