"""Small attention networks on matrix-valued vertices: the workloads of the reference's Perceptron / Encoder /
EncoderDecoder tests (functions of the same name in its examples package; shapes and test points in
tests/examples/example_test.py:240-349), written here against the engine's jnp / jnn stand-ins.

Building blocks, in the order they are evaluated (the order fixes the vertex numbering):
  normalise   : layer normalisation over the last axis, variance with the 1/n convention and eps = 1e-6
  attend      : softmax(q^T k) v, softmax over axis 1
  swish       : t * logistic(t)
  cross_entropy : -sum(targets * log softmax(scores))
"""
from ..tracer import jnn, jnp

_EPS = 1e-6


def _logistic(t):
    return 1. / (1. + jnp.exp(-t))


def _swish(t):
    return t * _logistic(t)


def _cross_entropy(scores, targets):
    log_p = jnp.log(jnn.softmax(scores, axis=-1))
    return -jnp.sum(targets * log_p, axis=-1)


def _normalise(t, scale, shift):
    centre = jnp.mean(t, axis=-1)
    spread = jnp.sum((t - centre) ** 2, axis=-1) / t.shape[-1]
    return (t - centre) / jnp.sqrt(spread + _EPS) * scale + shift


def _attend(queries, keys, values):
    weights = jnn.softmax(queries.T @ keys, axis=1)
    return weights @ values


def _self_attention_block(h, w_q, w_k, w_v, w_out, bias, scale, shift):
    """Residual self-attention, normalisation, one swish layer."""
    mixed = h + _attend(w_q @ h, w_k @ h, w_v @ h)
    return _swish(w_out @ _normalise(mixed, scale, shift) + bias)


def _cross_attention_block(h, memory_q, memory_k, self_w, cross_w, w_out, bias, scales, shifts):
    """Residual self-attention on ``h``, then attention whose queries and keys come from ``memory_*``."""
    (sq, sk, sv), (cq, ck, cv) = self_w, cross_w
    first = _normalise(h + _attend(sq @ h, sk @ h, sv @ h), scales[0], shifts[0])
    second = _normalise(first + _attend(cq @ memory_q, ck @ memory_k, cv @ first), scales[1], shifts[1])
    return _swish(w_out @ second + bias)


def Perceptron(x, y, W1, b1, W2, b2, gamma, beta):
    """Two tanh layers with a layer norm in between, softmax cross-entropy against ``y``."""
    hidden = _normalise(jnp.tanh(W1 @ x + b1), gamma, beta)
    return _cross_entropy(jnp.tanh(W2 @ hidden + b2), y)


def Encoder(x, y, WQ1, WQ2, WK1, WK2, WV1, WV2, W1, W2, b1, b2, gamma0, beta0, gamma1, beta1):
    """Two stacked self-attention blocks, softmax cross-entropy against ``y``."""
    h = _self_attention_block(x, WQ1, WK1, WV1, W1, b1, gamma0, beta0)
    h = _self_attention_block(h, WQ2, WK2, WV2, W2, b2, gamma1, beta1)
    return _cross_entropy(h, y)


def EncoderDecoder(x, y, WQ1, WQ2, WQ3, WK1, WK2, WK3, WV1, WV2, WV3, W1, W2, b1, b2,
                   gamma0, beta0, gamma1, beta1, gamma2, beta2):
    """One encoder block whose output feeds the cross-attention of a decoder block; squared error against ``y``."""
    memory = _self_attention_block(x, WQ1, WK1, WV1, W1, b1, gamma0, beta0)
    # weight roles as the reference wires them (its decoder takes WQ2, WQ3, WK2 for the self-attention and
    # WK3, WV2, WV3 for the cross-attention: the positional call in its EncoderDecoder)
    decoded = _cross_attention_block(x, memory, memory, (WQ2, WQ3, WK2), (WK3, WV2, WV3), W2, b2,
                                     (gamma1, gamma2), (beta1, beta2))
    return .5 * (decoded - y) ** 2
