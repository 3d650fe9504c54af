"""`filter_jacve` (reference `equinox_bindings.py:44-135`, `tests/eqx_test.py`): differentiate with respect to the
floating-point array leaves of a model pytree; the result is shaped like the model."""
import dataclasses
from typing import Any, Callable

import numpy as np
import pytest

import graphax_b200 as gx
from graphax_b200 import jnp
from graphax_b200.filtering import filter_jacve, is_inexact_array, tree_flatten, tree_unflatten


@dataclasses.dataclass(frozen=True)
class Linear:                      # what an equinox Module is: a frozen dataclass of arrays and static fields
    weight: Any
    bias: Any
    use_bias: bool = True

    def __call__(self, x):
        y = x @ self.weight.T
        return y + self.bias if self.use_bias else y


@dataclasses.dataclass(frozen=True)
class MLP:
    layers: tuple
    activation: Callable = dataclasses.field(default=jnp.tanh)
    name: str = "mlp"

    def __call__(self, x):
        for layer in self.layers[:-1]:
            x = self.activation(layer(x))
        return self.layers[-1](x)


def _model(sizes, seed=0, make=lambda a: a):
    rng = np.random.default_rng(seed)
    layers = tuple(Linear(make((rng.standard_normal((o, i)) * i ** -0.5).astype(np.float32)),
                          make((rng.standard_normal(o) * 0.1).astype(np.float32))) for i, o in zip(sizes[:-1], sizes[1:]))
    return MLP(layers)


def test_pytree_round_trip_and_filtering():
    model = {"net": _model((3, 4, 2)), "step": 7, "tags": ["a", None], "scale": np.full((), 2.0, np.float32)}
    leaves, treedef = tree_flatten(model)
    rebuilt = tree_unflatten(treedef, leaves)
    assert rebuilt["step"] == 7 and rebuilt["tags"] == ["a", None] and rebuilt["net"].name == "mlp"
    assert rebuilt["net"].layers[1].weight is model["net"].layers[1].weight
    diff = [l for l in leaves if is_inexact_array(l)]
    assert len(diff) == 5                                    # 2 x (weight, bias) + scale; ints, strings, callables are static
    assert not is_inexact_array(np.ones(3, dtype=np.int32)) and not is_inexact_array(3.0)


def test_argnums_is_rejected_like_the_reference():
    with pytest.raises(ValueError, match="argnums"):
        filter_jacve(lambda m, x: m(x), order="rev", argnums=(0,))
    assert callable(filter_jacve(order="rev")(lambda m, x: m(x)))          # decorator form (fun=sentinel)


@pytest.mark.gpu
def test_small_model_gradients_shaped_like_the_model(built_library):
    """Vector-valued output, small matrices: the per-sample engine (every order works)."""
    import torch
    model = _model((3, 4, 2), make=lambda a: torch.from_numpy(a).cuda())
    x = torch.tensor([0.3, -0.2, 0.9], device="cuda")
    for order in ("rev", "fwd"):
        jac = filter_jacve(lambda m, v: m(v), order=order)(model, x)
        assert isinstance(jac, MLP) and jac.activation is None and jac.name is None     # non-differentiable leaves: None
        W1, b1, W2, b2 = [t.double().cpu().requires_grad_(True) for t in
                          (model.layers[0].weight, model.layers[0].bias, model.layers[1].weight, model.layers[1].bias)]
        out = torch.tanh(x.double().cpu() @ W1.T + b1) @ W2.T + b2
        for k in range(2):
            gs = torch.autograd.grad(out[k], (W1, b1, W2, b2), retain_graph=True)
            got = (jac.layers[0].weight[k], jac.layers[0].bias[k], jac.layers[1].weight[k], jac.layers[1].bias[k])
            for g, r in zip(got, gs):
                assert tuple(g.shape) == tuple(r.shape)
                assert torch.allclose(g.double().cpu(), r, rtol=1e-5, atol=1e-6)
    jac, counts = filter_jacve(lambda m, v: m(v), order="rev", count_ops=True)(model, x)
    assert counts["num_muls"] > 0 and isinstance(jac, MLP)


@pytest.mark.gpu
def test_loss_of_a_batched_mlp_like_the_reference_test(built_library):
    """tests/eqx_test.py::test_simple_NN_and_loss: mean squared error of an MLP over a batch, gradients w.r.t. the
    model.  Weight matrices of 64 x 64 and more take the tensor-valued path (tcgen05 contractions, float32 accuracy)."""
    import torch
    model = _model((10, 64, 64, 64, 5), seed=42, make=lambda a: torch.from_numpy(a).cuda())
    x = torch.ones((64, 10), device="cuda") * torch.linspace(0.5, 1.5, 64, device="cuda")[:, None]
    y = torch.ones((64, 5), device="cuda")

    def loss_fn(m, x, y):
        return jnp.mean((m(x) - y) ** 2)
    grads = filter_jacve(loss_fn, order="rev")(model, x, y)
    params = [t.double().cpu().requires_grad_(True) for l in model.layers for t in (l.weight, l.bias)]
    h = x.double().cpu()
    for i in range(0, len(params) - 2, 2):
        h = torch.tanh(h @ params[i].T + params[i + 1])
    loss = ((h @ params[-2].T + params[-1] - y.double().cpu()) ** 2).mean()
    ref = torch.autograd.grad(loss, params)
    got = [t for l in grads.layers for t in (l.weight, l.bias)]
    for g, r in zip(got, ref):
        assert tuple(g.shape) == tuple(r.shape)
        bound = 1e-5 * float(r.abs().max()) + 1e-4 * r.abs()          # the reference's tree_allclose scale
        assert bool(((g.double().cpu() - r).abs() <= bound).all())
