from .._impl import safe_map  # noqa: F401
