"""Minimal stand-in for torch_geometric.data.Data (attribute bag of tensors with ``.to``)."""
from __future__ import annotations

import torch


class Data:
    def __init__(self, **kwargs):
        self.__dict__.update(kwargs)

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    def to(self, device, non_blocking=False):
        out = Data()
        for k, v in self.__dict__.items():
            out.__dict__[k] = v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v
        return out

    def __repr__(self):
        parts = [f"{k}={list(v.shape)}" if torch.is_tensor(v) else f"{k}={v!r}" for k, v in self.__dict__.items()
                 if not k.startswith("_")]
        return f"Data({', '.join(parts)})"
