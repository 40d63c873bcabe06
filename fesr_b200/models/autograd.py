"""torch.autograd bridge: forward/backward of the whole model are single C-ABI calls."""
from __future__ import annotations

import torch

from .. import ops


def _unflatten(model, flat):
    """inverse of _MeshModel._flat_params(): list of tensors -> param dict."""
    t = model.param_tensors()
    nm = len(t["mlp_w"])
    it = iter(flat)
    out = {"fc1_w": next(it), "fc1_b": next(it)}
    out["mlp_w"] = [next(it) for _ in range(nm)]
    out["mlp_b"] = [next(it) for _ in range(nm)]
    if t.get("lin_w") is not None:
        out["lin_w"], out["lin_b"] = next(it), next(it)
    else:
        out["lin_w"] = out["lin_b"] = None
    out["root"], out["bias"], out["fc2_w"], out["fc2_b"] = next(it), next(it), next(it), next(it)
    return out


def _flatten(model, d):
    flat = [d["fc1_w"], d["fc1_b"], *d["mlp_w"], *d["mlp_b"]]
    if d.get("lin_w") is not None:
        flat += [d["lin_w"], d["lin_b"]]
    flat += [d["root"], d["bias"], d["fc2_w"], d["fc2_b"]]
    return flat


class NNConvFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, csr, edge_attr, prec, x, *flat_params):
        tensors = _unflatten(model, [p.detach() for p in flat_params])
        y, ws = ops.nnconv_forward(model.dims, tensors, x.detach(), csr, edge_attr.detach(), prec,
                                   keep_for_backward=True)
        ctx.model, ctx.csr, ctx.prec, ctx.ws = model, csr, prec, ws
        ctx.save_for_backward(x, edge_attr, *flat_params)
        return y

    @staticmethod
    def backward(ctx, grad_y):
        x, edge_attr, *flat_params = ctx.saved_tensors
        tensors = _unflatten(ctx.model, [p.detach() for p in flat_params])
        grads, grad_x = ops.nnconv_backward(ctx.model.dims, tensors, x.detach(), ctx.csr, edge_attr.detach(), ctx.prec,
                                            grad_y.contiguous(), ctx.ws, need_grad_x=ctx.needs_input_grad[4])
        ctx.ws = None
        return (None, None, None, None, grad_x, *_flatten(ctx.model, grads))
