"""`jax.nn` facade of jaxlite (only what the reference's examples reach at import time)."""
from ._impl import logistic as sigmoid, tanh, _unavailable
softmax = _unavailable("nn.softmax")
relu = _unavailable("nn.relu")
