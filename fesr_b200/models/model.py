"""Drop-in mesh models: same class names, constructor arguments, ``forward(x, edge_index,
edge_attr)`` signature and ``state_dict`` keys as the reference's models/model.py, with the
whole forward (and backward) executed by libfesr.so on a B200.

  reference class         file:line                this module
  KernelNN                models/model.py:543-561  KernelNN
  NNConv_old              models/model.py:451-540  NNConv_old   (parameter container + layer call)
  DenseNet                models/model.py:289-315  DenseNet     (parameter container)
  TEECNet                 models/model.py:259-286  TEECNet
  KernelConv              models/model.py:365-448  KernelConv   (parameter container + layer call)

There is no PyTorch / CPU implementation of the message passing here: a CPU tensor or a
missing libfesr.so raises.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict

import torch
import torch.nn as nn

from .. import _lib, ops
from .._lib import FesrError


def _uniform(size, tensor):
    # torch_geometric.nn.inits.uniform, as used at models/model.py:417-419, 517-519
    bound = 1.0 / math.sqrt(size)
    with torch.no_grad():
        tensor.uniform_(-bound, bound)


class DenseNet(nn.Module):
    """Edge MLP parameter container; layer layout as models/model.py:297-310."""

    def __init__(self, layers, nonlinearity, out_nonlinearity=None, normalize=False):
        super().__init__()
        if normalize or out_nonlinearity is not None:
            raise NotImplementedError("fesr_b200 builds the DenseNet variants the mesh models use "
                                      "(no BatchNorm, no output nonlinearity)")
        self.n_layers = len(layers) - 1
        assert self.n_layers >= 1
        self.dims = list(layers)
        self.layers = nn.ModuleList()
        for j in range(self.n_layers):
            self.layers.append(nn.Linear(layers[j], layers[j + 1]))
            if j != self.n_layers - 1:
                self.layers.append(nonlinearity())

    def linears(self):
        return [l for l in self.layers if isinstance(l, nn.Linear)]

    def forward(self, x):
        raise FesrError("DenseNet is evaluated inside the fused CUDA forward of KernelNN / TEECNet; "
                        "it has no standalone forward in fesr_b200")


class _GraphCache:
    """Remembers the destination CSR of the last few edge_index tensors seen by a model."""

    def __init__(self, size=4):
        self.size = size
        self.entries = OrderedDict()

    def get(self, edge_index, n):
        if isinstance(edge_index, ops.Csr):
            return edge_index
        key = (edge_index.data_ptr(), tuple(edge_index.shape), int(n), edge_index._version, edge_index.device.index)
        hit = self.entries.get(key)
        if hit is not None:
            self.entries.move_to_end(key)
            return hit[0]
        csr = ops.csr_build(edge_index, n)
        # hold a reference to edge_index so its storage (and data_ptr) cannot be recycled
        self.entries[key] = (csr, edge_index)
        while len(self.entries) > self.size:
            self.entries.popitem(last=False)
        return csr


class _MeshModel(nn.Module):
    """Common host side of KernelNN / TEECNet."""

    kind = None

    def _init_common(self, width, in_ch, out_ch, layers):
        self._dims = _lib.model_dims(self.kind, width, in_ch, out_ch, layers) if os.path.exists(_lib.LIB_PATH) else None
        self._dims_args = (self.kind, width, in_ch, out_ch, layers)
        self._graphs = _GraphCache()
        self.precision = os.environ.get("FESR_PRECISION", "fp32")

    @property
    def dims(self):
        if self._dims is None:
            self._dims = _lib.model_dims(*self._dims_args)
        return self._dims

    def param_tensors(self):
        raise NotImplementedError

    _CACHE_KEYS = ("_plist", "_plan", "_plan_key")

    def __deepcopy__(self, memo):
        # (the scheduler deep-copies its model once per cluster) the predict-time launch plan holds device tensors and
        # ctypes structs that belong to THIS instance: the copy starts without one
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k not in self._CACHE_KEYS:
                new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def _predict_plan(self, csr, edge_attr, prec):
        """The validated launch plan of the predict-time forward, kept as long as the graph, the edge lengths, the
        precision and the parameter storages / versions are the ones it was made for."""
        plist = self.__dict__.get("_plist")
        if plist is None:
            plist = self.__dict__["_plist"] = list(self.parameters())
        key = (csr.rowptr.data_ptr(), csr.src.data_ptr(), csr.n, csr.E, edge_attr.data_ptr(), edge_attr._version, prec,
               getattr(self, "ws_tag", "fwd"), ops.weights_generation(), tuple((p.data_ptr(), p._version) for p in plist))
        if self.__dict__.get("_plan_key") != key:
            detached = {k: (None if v is None else v.detach() if torch.is_tensor(v) else [t.detach() for t in v])
                        for k, v in self.param_tensors().items()}
            self.__dict__["_plan"] = ops.make_forward_plan(self.dims, detached, csr, edge_attr.detach(), prec,
                                                           getattr(self, "ws_tag", "fwd"))
            self.__dict__["_plan_key"] = key
        return self.__dict__["_plan"]

    @torch.no_grad()
    def edge_phase(self, csr, edge_attr):
        """Extension (predict only): issues the part of the next forward on (csr, edge_attr) that does not read x --
        weight preparation + edge MLP -- so that a caller whose x is still on the host can start the GPU before it
        issues the copies.  The next forward on the same graph picks the edge features up from the workspace."""
        ops.run_forward_plan(self._predict_plan(csr, edge_attr, _lib.PRECISIONS[self.precision]), None, edge_only=True)

    def forward(self, x, edge_index, edge_attr, x_ready=None, out=None):
        """x_ready (extension, predict only): CUDA event after which `x` is valid -- see ops.run_forward_plan.
        out (extension, predict only): [n, out_ch] fp32 block the result is written into."""
        if not x.is_cuda:
            raise FesrError(f"{type(self).__name__}.forward needs CUDA tensors on a B200; fesr_b200 has no CPU path")
        csr = self._graphs.get(edge_index, x.shape[0])
        prec = _lib.PRECISIONS[self.precision]
        if not torch.is_grad_enabled():
            return ops.run_forward_plan(self._predict_plan(csr, edge_attr, prec), x.detach(), x_ready, out=out)
        if out is not None:
            raise FesrError("out= is a predict-time extension (torch.no_grad())")
        tensors = self.param_tensors()
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if needs_grad:
            if x_ready is not None:
                torch.cuda.current_stream(x.device).wait_event(x_ready)
            from .autograd import NNConvFunction
            return NNConvFunction.apply(self, csr, edge_attr, prec, x, *self._flat_params())
        detached = {k: (None if v is None else v.detach() if torch.is_tensor(v) else [t.detach() for t in v])
                    for k, v in tensors.items()}
        # `ws_tag`: the scheduler gives every per-cluster model its own workspace so that each keeps its prepared
        # weights between predict calls (one shared workspace would re-prepare on every model switch)
        return ops.nnconv_forward(self.dims, detached, x.detach(), csr, edge_attr.detach(), prec,
                                  ws_tag=getattr(self, "ws_tag", "fwd"), x_ready=x_ready)

    def _flat_params(self):
        t = self.param_tensors()
        flat = [t["fc1_w"], t["fc1_b"], *t["mlp_w"], *t["mlp_b"]]
        if t.get("lin_w") is not None:
            flat += [t["lin_w"], t["lin_b"]]
        flat += [t["root"], t["bias"], t["fc2_w"], t["fc2_b"]]
        return flat


class NNConv_old(nn.Module):
    """Edge-conditioned convolution parameters (models/model.py:488-519): ``nn`` (DenseNet),
    ``root`` [in, out], ``bias`` [out]; mean aggregation."""

    def __init__(self, in_channels, out_channels, nn, aggr='add', root_weight=True, bias=True, **kwargs):
        super().__init__()
        if aggr != 'mean' or not root_weight or not bias or in_channels != out_channels:
            raise NotImplementedError("fesr_b200 builds the configuration KernelNN uses: aggr='mean', "
                                      "root_weight=True, bias=True, in_channels == out_channels")
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.nn = nn
        self.aggr = aggr
        self.root = torch.nn.Parameter(torch.empty(in_channels, out_channels))
        self.bias = torch.nn.Parameter(torch.empty(out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        for l in self.nn.linears():
            l.reset_parameters()
        _uniform(self.in_channels, self.root)
        _uniform(self.in_channels, self.bias)

    def __repr__(self):
        return '{}({}, {})'.format(self.__class__.__name__, self.in_channels, self.out_channels)


class KernelNN(_MeshModel):
    """Graph neural operator (``--model=neuralop``), models/model.py:543-561."""

    kind = _lib.KERNELNN

    def __init__(self, width, ker_width, depth, ker_in=1, in_width=3, out_width=3):
        super().__init__()
        if ker_in != 1 or ker_width != width:
            raise NotImplementedError("fesr_b200 builds ker_in=1, ker_width == width (what utils.init_model passes)")
        self.depth = depth
        self.fc1 = torch.nn.Linear(in_width, width)
        kernel = DenseNet([ker_in, ker_width, ker_width, width ** 2], torch.nn.ReLU)
        self.conv1 = NNConv_old(width, width, kernel, aggr='mean')
        self.fc2 = torch.nn.Linear(width, out_width)
        self._init_common(width, in_width, out_width, depth)

    def param_tensors(self):
        lin = self.conv1.nn.linears()
        return {"fc1_w": self.fc1.weight, "fc1_b": self.fc1.bias,
                "mlp_w": [l.weight for l in lin], "mlp_b": [l.bias for l in lin],
                "lin_w": None, "lin_b": None, "root": self.conv1.root, "bias": self.conv1.bias,
                "fc2_w": self.fc2.weight, "fc2_b": self.fc2.bias}


class KernelConv(nn.Module):
    """TEECNet's convolution parameters (models/model.py:394-419): ``root_param``, ``bias``,
    ``linear``, ``operator_kernel`` = DenseNet([in_edge, 32, 64, 128, out^2], LeakyReLU)."""

    def __init__(self, in_channel, out_channel, kernel=None, in_edge=5, num_layers=3, **kwargs):
        super().__init__()
        if in_edge != 1 or in_channel != out_channel:
            raise NotImplementedError("fesr_b200 builds in_edge=1, in_channel == out_channel (what TEECNet passes)")
        self.in_channels = in_channel
        self.out_channels = out_channel
        self.in_edge = in_edge
        self.root_param = nn.Parameter(torch.empty(in_channel, out_channel))
        self.bias = nn.Parameter(torch.empty(out_channel))
        self.linear = nn.Linear(in_channel, out_channel)
        self.operator_kernel = DenseNet([in_edge, 32, 64, 128, out_channel ** 2], nn.LeakyReLU)
        self.retrieve_weights = bool(kwargs['retrieve_weight'])     # required kwarg, models/model.py:404
        if self.retrieve_weights:
            self.weight_k = None
            self.weight_op = None
        self.reset_parameters()

    def reset_parameters(self):
        self.linear.reset_parameters()
        for l in self.operator_kernel.linears():
            l.reset_parameters()
        _uniform(self.in_channels, self.root_param)
        _uniform(self.in_channels, self.bias)

    def __repr__(self):
        return '{}({}, {})'.format(self.__class__.__name__, self.in_channels, self.out_channels)


class PowerSeriesKernel:
    """Placeholder for the ctor argument KernelConv ignores (models/model.py:402 is commented out)."""


class TEECNet(_MeshModel):
    """Taylor-series Expansion Error Correction Network (``--model=teecnet``), models/model.py:259-286."""

    kind = _lib.TEECNET

    def __init__(self, in_channels, width, out_channels, num_layers=4, **kwargs):
        super().__init__()
        self.num_layers = num_layers
        self.fc1 = nn.Linear(in_channels, width)
        self.kernel = KernelConv(width, width, kernel=PowerSeriesKernel, in_edge=1, num_layers=3, **kwargs)
        self.fc_out = nn.Linear(width, out_channels)
        self._init_common(width, in_channels, out_channels, num_layers)

    def param_tensors(self):
        lin = self.kernel.operator_kernel.linears()
        return {"fc1_w": self.fc1.weight, "fc1_b": self.fc1.bias,
                "mlp_w": [l.weight for l in lin], "mlp_b": [l.bias for l in lin],
                "lin_w": self.kernel.linear.weight, "lin_b": self.kernel.linear.bias,
                "root": self.kernel.root_param, "bias": self.kernel.bias,
                "fc2_w": self.fc_out.weight, "fc2_b": self.fc_out.bias}
