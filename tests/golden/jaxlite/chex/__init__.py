"""chex stand-in: graphax only uses `chex.Array` as a type annotation (`sparse/tensor.py:14`)."""
from typing import Any
Array = Any
