"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the PyTorch-Geometric 2.6.1 pieces that
graphchem's MoleculeGCN hot path executes.

PARITY UNPINNED (numerics): `torch-geometric==2.6.1` (pinned in the reference's
`pyproject.toml:17-21`) is an un-vendored third-party dependency: it is not under
/root/reference, not installed in this image and not installable (no network).  The
reference's own tests hold only SHAPE assertions for this path (`tests/test_gcn.py:60-62,
99-101, 149-151`), no golden activations/gradients.  What follows is therefore a restatement
of PyG 2.6.1's *published* algorithm, anchored on the reference's call sites:

  * `graphchem/nn/gcn.py:84-86`   GeneralConv(D, D, D, aggr=aggr)
  * `graphchem/nn/gcn.py:89-91`   EdgeConv(nn.Sequential(nn.Linear(2D, D)))      (aggr='max')
  * `graphchem/nn/gcn.py:163,172` conv(x, edge_index[, edge_attr]) calls
  * `graphchem/nn/gcn.py:181`     global_add_pool(out_atom, batch)
  * `graphchem/data/structs.py:4,7,61-66` Data(x, edge_index, edge_attr, y)
  * `examples/predict_cn.ipynb:208-213`   torch_geometric.loader.DataLoader -> Batch

PyG 2.6.1 files restated (function by function, not in tree):
  utils/_scatter.py::scatter, nn/conv/message_passing.py::{propagate,_collect,_lift,
  _check_input}, nn/aggr/basic.py::{Sum,Mean,Max}Aggregation, nn/conv/edge_conv.py,
  nn/conv/general_conv.py, nn/dense/linear.py::Linear, nn/pool/glob.py::global_add_pool,
  data/{data,batch,collate}.py (the subset used for x/edge_index/edge_attr/y).

If `import torch_geometric` ever succeeds, `tests/test_oracle.py::test_oracle_matches_real_pyg`
compares these classes with the real ones bit for bit.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package (graphchem_b200/) never does.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
from torch import Tensor


# ----------------------------------------------------------------------------------------------
# utils/_scatter.py::scatter  (the torch_scatter-free branch: the reference does not depend on
# torch_scatter, pyproject.toml:17-21)
# ----------------------------------------------------------------------------------------------
def _broadcast(index: Tensor, ref: Tensor, dim: int) -> Tensor:
    dim = ref.dim() + dim if dim < 0 else dim
    size = [1] * dim + [-1] + [1] * (ref.dim() - dim - 1)
    return index.view(size).expand_as(ref)


def scatter(src: Tensor, index: Tensor, dim: int = 0, dim_size: Optional[int] = None,
            reduce: str = "sum") -> Tensor:
    if index.dim() != 1:
        raise ValueError("index must be one-dimensional")
    dim = src.dim() + dim if dim < 0 else dim
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    size = list(src.size())
    size[dim] = dim_size
    if reduce in ("sum", "add"):
        return src.new_zeros(size).scatter_add_(dim, _broadcast(index, src, dim), src)
    if reduce == "mean":
        count = src.new_zeros(dim_size)
        count.scatter_add_(0, index, src.new_ones(src.size(dim)))
        count = count.clamp(min=1)
        out = src.new_zeros(size).scatter_add_(dim, _broadcast(index, src, dim), src)
        return out / _broadcast(count, out, dim)
    if reduce in ("max", "min", "amax", "amin"):
        # zeros-initialised, include_self=False: rows without contributions stay exactly 0
        return src.new_zeros(size).scatter_reduce_(
            dim, _broadcast(index, src, dim), src, reduce=f"a{reduce[-3:]}", include_self=False)
    raise ValueError(f"unsupported reduce {reduce!r}")


# ----------------------------------------------------------------------------------------------
# nn/dense/linear.py::Linear  (kaiming_uniform(a=sqrt(5)) weight, U(+-1/sqrt(fan_in)) bias)
# ----------------------------------------------------------------------------------------------
class Linear(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, bias: bool = True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self) -> None:
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1.0 / math.sqrt(self.in_channels) if self.in_channels > 0 else 0.0
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x: Tensor) -> Tensor:
        return nn.functional.linear(x, self.weight, self.bias)


# ----------------------------------------------------------------------------------------------
# nn/conv/message_passing.py  (flow='source_to_target': j = edge_index[0], i = edge_index[1])
# ----------------------------------------------------------------------------------------------
class MessagePassing(nn.Module):
    def __init__(self, aggr: str = "add", node_dim: int = -2):
        super().__init__()
        self.aggr = aggr
        self.node_dim = node_dim

    @staticmethod
    def _check_input(edge_index: Tensor) -> None:
        if not isinstance(edge_index, Tensor):
            raise ValueError("`MessagePassing.propagate` only supports integer tensors of "
                             "shape `[2, num_messages]`")
        if edge_index.dtype not in (torch.int64, torch.int32, torch.int16, torch.int8, torch.uint8):
            raise ValueError(f"Expected 'edge_index' to be of integer type (got '{edge_index.dtype}')")
        if edge_index.dim() != 2:
            raise ValueError(f"Expected 'edge_index' to be two-dimensional (got {edge_index.dim()} dimensions)")
        if edge_index.size(0) != 2:
            raise ValueError(f"Expected 'edge_index' to have size '2' in the first dimension (got '{edge_index.size(0)}')")

    def _lift(self, src: Tensor, index: Tensor) -> Tensor:
        try:
            return src.index_select(self.node_dim, index)
        except (IndexError, RuntimeError) as e:
            if index.numel() > 0 and (int(index.min()) < 0 or int(index.max()) >= src.size(self.node_dim)):
                raise IndexError(
                    f"Encountered an index error. Please ensure that all indices in 'edge_index' "
                    f"point to valid indices in the interval [0, {src.size(self.node_dim)}) in your "
                    f"node feature matrix and try again.") from e
            raise

    def _aggregate(self, inputs: Tensor, index: Tensor, dim_size: int) -> Tensor:
        reduce = {"add": "sum", "sum": "sum", "mean": "mean", "max": "max"}[self.aggr]
        return scatter(inputs, index, dim=self.node_dim, dim_size=dim_size, reduce=reduce)


class EdgeConv(MessagePassing):
    """nn/conv/edge_conv.py: out_i = max_{j->i} nn([x_i, x_j - x_i]); output sized by x.size(0)."""

    def __init__(self, nn_module: nn.Module, aggr: str = "max"):
        super().__init__(aggr=aggr)
        self.nn = nn_module

    def forward(self, x: Tensor, edge_index: Tensor) -> Tensor:
        self._check_input(edge_index)
        x_j = self._lift(x, edge_index[0])
        x_i = self._lift(x, edge_index[1])
        msg = self.nn(torch.cat([x_i, x_j - x_i], dim=-1))
        return self._aggregate(msg, edge_index[1], dim_size=x.size(self.node_dim))


class GeneralConv(MessagePassing):
    """nn/conv/general_conv.py with the defaults graphchem uses: directed_msg=True, heads=1,
    attention=False, skip_linear=False, l2_normalize=False, bias=True.  Registration order of
    the sub-modules (lin_msg, lin_self, lin_edge) follows PyG."""

    def __init__(self, in_channels: int, out_channels: int, in_edge_channels: Optional[int] = None,
                 aggr: str = "add"):
        super().__init__(aggr=aggr, node_dim=0)
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, 1
        self.lin_msg = Linear(in_channels, out_channels * self.heads, bias=True)
        if in_channels != out_channels:
            self.lin_self = Linear(in_channels, out_channels, bias=True)
        else:
            self.lin_self = nn.Identity()
        if in_edge_channels is not None:
            self.lin_edge = Linear(in_edge_channels, out_channels * self.heads, bias=True)

    def forward(self, x: Tensor, edge_index: Tensor, edge_attr: Optional[Tensor] = None) -> Tensor:
        self._check_input(edge_index)
        x_j = self._lift(x, edge_index[0])
        _x_i = self._lift(x, edge_index[1])  # gathered by PyG, unused when directed_msg=True
        msg = self.lin_msg(x_j)
        if edge_attr is not None:
            msg = msg + self.lin_edge(edge_attr)
        msg = msg.view(-1, self.heads, self.out_channels)
        out = self._aggregate(msg, edge_index[1], dim_size=x.size(0))
        out = out.mean(dim=1)
        return out + self.lin_self(x)


# nn/pool/glob.py::global_add_pool
def global_add_pool(x: Tensor, batch: Optional[Tensor], size: Optional[int] = None) -> Tensor:
    dim = -1 if x.dim() == 1 else -2
    if batch is None:
        return x.sum(dim=dim, keepdim=x.dim() <= 2)
    return scatter(x, batch, dim=dim, dim_size=size, reduce="sum")


# nn/pool/glob.py::global_mean_pool (not called by graphchem; the product offers it as `pool="mean"`)
def global_mean_pool(x: Tensor, batch: Optional[Tensor], size: Optional[int] = None) -> Tensor:
    dim = -1 if x.dim() == 1 else -2
    if batch is None:
        return x.mean(dim=dim, keepdim=x.dim() <= 2)
    return scatter(x, batch, dim=dim, dim_size=size, reduce="mean")


# ----------------------------------------------------------------------------------------------
# data/data.py, data/batch.py, data/collate.py -- the subset MoleculeGraph uses
# ----------------------------------------------------------------------------------------------
class Data:
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, batch=None):
        self.x, self.edge_index, self.edge_attr, self.y, self.batch = x, edge_index, edge_attr, y, batch

    @property
    def num_nodes(self) -> Optional[int]:
        if self.x is not None:
            return self.x.size(0)
        if self.edge_index is not None and self.edge_index.numel() > 0:
            return int(self.edge_index.max()) + 1
        return None

    @property
    def num_node_features(self) -> int:
        if self.x is None:
            return 0
        return 1 if self.x.dim() == 1 else self.x.size(-1)


def collate(graphs: Sequence[Data]) -> Data:
    """Batch.from_data_list for the keys x, edge_index, edge_attr, y (insertion order of
    structs.py:61-66): cat on dim 0, except keys containing 'index' which are cat'ed on dim -1
    and incremented by the cumulative node count; adds `batch`, `ptr`, `num_graphs`."""
    num_nodes = [g.x.size(0) for g in graphs]
    offsets = [0]
    for n in num_nodes:
        offsets.append(offsets[-1] + n)
    out = Data()
    out.x = torch.cat([g.x for g in graphs], dim=0)
    out.edge_index = torch.cat([g.edge_index + off for g, off in zip(graphs, offsets[:-1])], dim=-1)
    out.edge_attr = torch.cat([g.edge_attr for g in graphs], dim=0)
    if all(g.y is not None for g in graphs):
        out.y = torch.cat([g.y for g in graphs], dim=0)
    out.batch = torch.repeat_interleave(torch.arange(len(graphs)), torch.tensor(num_nodes, dtype=torch.long))
    out.ptr = torch.tensor(offsets, dtype=torch.long)
    out.num_graphs = len(graphs)
    return out
