from .._impl import pjit_p  # noqa: F401
