"""Two-layer perceptron loss over a batch - the dense weight-Jacobian workload (config 5).

The reference's version (``tests/examples/example_test.py:204-215``) vmaps a per-sample network and sums;
here the batch is an explicit leading axis, so the graph's vertices are [B, H] matrices and the edges into
the weights are the dense ``[B,H]^T [B,H']`` contractions that go to the tensor cores."""
from ..tracer import jnp, lax

_NT = (((1,), (1,)), ((), ()))          # x[B,I] . W[H,I]^T


def MLP_loss(x, W1, b1, W2, b2, y):
    z1 = lax.dot_general(x, W1, _NT) + b1
    a1 = jnp.tanh(z1)
    z2 = lax.dot_general(a1, W2, _NT) + b2
    d = jnp.tanh(z2) - y
    return .5 * jnp.sum(d ** 2)
