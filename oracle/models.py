"""Oracle (CPU PyTorch): the two mesh models, node weight / loss, the train step.

TEST INFRASTRUCTURE -- see oracle/__init__.py.  Executes the operations in the
REFERENCE ORDER (edge MLP evaluated per edge in every layer, per-edge [w,w] matrix,
batched mat-vec, scatter-add mean), which is what the CUDA path -- that reorders the
contraction -- must match within tolerance.

Reference anchors (/root/reference):
  * DenseNet ............. models/model.py:289-315 (Linear, act, ..., last Linear bare)
  * NNConv_old ........... models/model.py:451-540 (message :527-529, update :531-536)
  * KernelNN ............. models/model.py:543-561
  * KernelConv ........... models/model.py:365-448 (message :426-441, update :444-445)
  * TEECNet .............. models/model.py:259-286
  * mean aggregation ..... torch_geometric==2.6.1 MessagePassing(aggr='mean'),
                           flow source_to_target: x_j = x[edge_index[0]], reduce over
                           edge_index[1], sum / clamp(count, min=1)
  * GradientbasedLoss .... models/scheduler_gnn.py:472-514
  * train step ........... models/scheduler_gnn.py:388-417 (MSELoss, Adam)
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def mean_aggregate(msg: torch.Tensor, dst: torch.Tensor, n: int) -> torch.Tensor:
    out = torch.zeros(n, msg.shape[1], dtype=msg.dtype)
    out.index_add_(0, dst, msg)
    cnt = torch.zeros(n, dtype=msg.dtype)
    cnt.index_add_(0, dst, torch.ones(dst.shape[0], dtype=msg.dtype))
    return out / cnt.clamp(min=1).unsqueeze(1)


class DenseNet(nn.Module):
    def __init__(self, layers, nonlinearity):
        super().__init__()
        self.layers = nn.ModuleList()
        nl = len(layers) - 1
        for j in range(nl):
            self.layers.append(nn.Linear(layers[j], layers[j + 1]))
            if j != nl - 1:
                self.layers.append(nonlinearity())

    def forward(self, x):
        for l in self.layers:
            x = l(x)
        return x


def _uniform(size, t):
    bound = 1.0 / math.sqrt(size)
    with torch.no_grad():
        t.uniform_(-bound, bound)


class NNConvOracle(nn.Module):
    def __init__(self, w, kernel):
        super().__init__()
        self.w = w
        self.nn = kernel
        self.root = nn.Parameter(torch.empty(w, w))
        self.bias = nn.Parameter(torch.empty(w))
        _uniform(w, self.root)
        _uniform(w, self.bias)

    def forward(self, x, edge_index, edge_attr):
        pseudo = edge_attr.unsqueeze(-1) if edge_attr.dim() == 1 else edge_attr
        src, dst = edge_index[0], edge_index[1]
        weight = self.nn(pseudo).view(-1, self.w, self.w)
        msg = torch.matmul(x[src].unsqueeze(1), weight).squeeze(1)
        aggr = mean_aggregate(msg, dst, x.shape[0])
        return aggr + torch.mm(x, self.root) + self.bias


class KernelNNOracle(nn.Module):
    """state_dict keys identical to the reference's KernelNN."""

    def __init__(self, width, ker_width, depth, ker_in=1, in_width=3, out_width=3):
        super().__init__()
        self.depth = depth
        self.fc1 = nn.Linear(in_width, width)
        self.conv1 = NNConvOracle(width, DenseNet([ker_in, ker_width, ker_width, width ** 2], nn.ReLU))
        self.fc2 = nn.Linear(width, out_width)

    def forward(self, x, edge_index, edge_attr):
        x = self.fc1(x)
        for _ in range(self.depth):
            x = F.relu(self.conv1(x, edge_index, edge_attr))
        return self.fc2(x)


class KernelConvOracle(nn.Module):
    def __init__(self, w):
        super().__init__()
        self.w = w
        self.root_param = nn.Parameter(torch.empty(w, w))
        self.bias = nn.Parameter(torch.empty(w))
        self.linear = nn.Linear(w, w)
        self.operator_kernel = DenseNet([1, 32, 64, 128, w ** 2], nn.LeakyReLU)
        _uniform(w, self.root_param)
        _uniform(w, self.bias)

    def forward(self, x, edge_index, edge_attr):
        pseudo = edge_attr.unsqueeze(-1) if edge_attr.dim() == 1 else edge_attr
        src, dst = edge_index[0], edge_index[1]
        weight_op = self.operator_kernel(pseudo).view(-1, self.w, self.w)
        x_j = self.linear(x[src])        # model.py:431 (x_i at :430 is computed and discarded)
        msg = torch.matmul(x_j.unsqueeze(1), weight_op).squeeze(1)
        aggr = mean_aggregate(msg, dst, x.shape[0])
        return aggr + torch.mm(x, self.root_param) + self.bias


class TEECNetOracle(nn.Module):
    """state_dict keys identical to the reference's TEECNet."""

    def __init__(self, in_channels, width, out_channels, num_layers=4, **kwargs):
        super().__init__()
        self.num_layers = num_layers
        self.fc1 = nn.Linear(in_channels, width)
        self.kernel = KernelConvOracle(width)
        self.fc_out = nn.Linear(width, out_channels)

    def forward(self, x, edge_index, edge_attr):
        x = self.fc1(x)
        for _ in range(self.num_layers):
            x = self.kernel(x, edge_index, edge_attr)
        return self.fc_out(x)


def make_model(kind: str, width=43, num_layers=5, in_channels=4, out_channels=4):
    if kind == "neuralop":
        return KernelNNOracle(width, width, num_layers, in_width=in_channels, out_width=out_channels)
    if kind == "teecnet":
        return TEECNetOracle(in_channels, width, out_channels, num_layers=num_layers)
    raise ValueError(kind)


# --------------------------------------------------------------------------------------
# a9: GradientbasedLoss
# --------------------------------------------------------------------------------------
def edge_weight(pred, data, edge_index, edge_attr):
    ea = edge_attr if edge_attr.dim() == 2 else edge_attr.unsqueeze(1)
    grad_pred = (pred[edge_index[0]] - pred[edge_index[1]]) / ea
    grad_data = (data[edge_index[0]] - data[edge_index[1]]) / ea
    return torch.max(grad_pred - grad_data, dim=1)[0]


def compute_node_weight(pred, data, edge_index, edge_attr, num_nodes):
    """scheduler_gnn.py:503-514: scatter_add by edge_index[0], global sum, broadcast."""
    ew = edge_weight(pred, data, edge_index, edge_attr)
    node_weight = torch.zeros(num_nodes, dtype=pred.dtype)
    node_weight.scatter_add_(0, edge_index[0], ew)
    return torch.sum(node_weight) * torch.ones(num_nodes, dtype=pred.dtype)


def gradient_loss(pred, data, edge_index, edge_attr, max_weight=1.0):
    """scheduler_gnn.py:481-501: scatter by edge_index[1], clamp(max), sum, * mse."""
    ew = edge_weight(pred, data, edge_index, edge_attr)
    node_weight = torch.zeros(pred.shape[0], dtype=pred.dtype)
    node_weight.scatter_add_(0, edge_index[1], ew)
    node_weight = torch.clamp(node_weight, max=max_weight)
    return (pred - data).pow(2).mean() * torch.sum(node_weight)


# --------------------------------------------------------------------------------------
# a11: one training step
# --------------------------------------------------------------------------------------
def train_step(model, optimizer, x, edge_index, edge_attr, y):
    """scheduler_gnn.py:398-408: zero_grad, forward, MSELoss, backward, Adam.step."""
    optimizer.zero_grad()
    out = model(x, edge_index, edge_attr)
    loss = F.mse_loss(out, y)
    loss.backward()
    optimizer.step()
    return loss.detach()
