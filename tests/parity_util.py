"""Shared helpers of the GPU parity tests (test infrastructure).

Error measures: `rel` = whole-tensor 2-norm ratio; `maxabs` = largest element error scaled by the largest
reference element (a per-element bound that does not blow up on entries that are ~0 by cancellation).

Rounding-ambiguous activations.  leaky_relu'(x) jumps from 0.01 to 1 at x = 0.  A pre-activation whose exact
value lies within fp32 rounding of 0 gets its branch from the SUMMATION ORDER of the implementation (MKL's
blocking, a CUDA-core k-loop, the tensor core's 3xTF32 chunks): measured on the bench's `micro` shape (10 M
layer-0 activations), 4 entries with |x| ~ 1e-7 took the other branch on the tensor-core path than on the CUDA-core
path / CPU torch, and those 4 entries alone move conv_layers.0.Q.weight.grad by 1.2e-3 of its norm (one element's
weight in a heavily cancelling sum over 10^7 of them) -- every single GEMM of that step is within 1e-6 of an fp64
product of its own inputs.  No fp32 implementation reproduces another one's decisions there, the reference on a
different thread count included.  `ambiguous_activations` finds those entries with a float64 evaluation
(|x64| < tau * (sum_k |a_k||b_k| + |bias|), tau = 1e-5, ~100x the fp32 error bound of the products) and returns
the PRODUCT's branch for each; the oracle then differentiates the same function (oracle._leaky_with_signs).
The tests bound their number (they must stay a ~1e-6 fraction) and hold everything else to 1e-4.
"""
import numpy as np
import torch

TAU = 1e-5


def _t(x):
    return torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).detach().double().cpu()


def rel(a, b):
    a, b = _t(a), _t(b)
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def maxabs(a, b):
    a, b = _t(a), _t(b)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def assert_close(got, want, tol, what=""):
    r, m = rel(got, want), maxabs(got, want)
    assert r < tol and m < tol, f"{what}: norm-rel {r:.3e}, max-abs-scaled {m:.3e} (bar {tol:g})"


def ambiguous_activations(model, feats, plan, saved, tau=TAU):
    """{layer: {"Q": [(node, col, positive?)], "W": [...]}} and the number of activations examined, from the
    engine's own forward state (plan + saved of Engine.forward(keep=True))."""
    forced, examined = {}, 0
    in_dims = model.in_dim_per_layer
    for l, lp in enumerate(plan.layers):
        conv = model.conv_layers[l]
        h_in, z, cat, _, h, _, _ = saved[l]
        din = in_dims[l]
        if lp.zrows is not None:  # layer 0: z rows are gathered feature rows
            z_nodes = lp.zrows.long().cpu()
            X = h_in[lp.zrows.long(), :din]
        else:                     # layer l > 0: one z row per target of the layer below, in its order
            z_nodes = plan.layers[l - 1].nodes.cpu()
            X = h_in[:, :din]
        entries = {}
        for key, A, W, b, out, nodes in (("Q", X, conv.Q.weight, conv.Q.bias, z, z_nodes),
                                         ("W", cat, conv.W.weight, conv.W.bias, h, lp.nodes.cpu())):
            A64, W64, b64 = A.detach().double(), W.detach().double(), b.detach().double()
            pre = A64 @ W64.t() + b64
            bound = A64.abs() @ W64.abs().t() + b64.abs()
            amb = (pre.abs() < tau * bound).nonzero().cpu()
            examined += pre.numel()
            pos = (out > 0).cpu()
            entries[key] = [(int(nodes[r]), int(c), bool(pos[r, c])) for r, c in amb.tolist()]
        forced[l] = entries
    return forced, examined


def count_forced(forced):
    return sum(len(v) for d in forced.values() for v in d.values())
