"""Data contract of the hot path: MoleculeGraph / MoleculeDataset plus PyG-free stand-ins for
`torch_geometric.data.Data`, `Batch` and `torch_geometric.loader.DataLoader`.

Mirrors reference `graphchem/data/structs.py:7-117` (field names, target normalisation rules and
the ValueError of :52-59) and the PyG 2.6.1 collate the notebooks rely on
(`examples/predict_cn.ipynb:212-213,245`).  PyTorch-Geometric itself is not
required; real PyG `Data`/`Batch` objects are accepted everywhere by duck typing.
"""
from __future__ import annotations

import numbers
from typing import Any, Iterable, Iterator, List, Optional, Sequence, Union

import numpy as np

import torch
from torch import Tensor


class Data:
    """Minimal `torch_geometric.data.Data`: attribute bag with the PyG-derived properties that
    `MoleculeGCN.forward` reads (reference gcn.py:144-149)."""

    def __init__(self, x: Optional[Tensor] = None, edge_index: Optional[Tensor] = None,
                 edge_attr: Optional[Tensor] = None, y: Optional[Tensor] = None, **kwargs: Any):
        self.x = x
        self.edge_index = edge_index
        self.edge_attr = edge_attr
        self.y = y
        for k, v in kwargs.items():
            setattr(self, k, v)

    # `batch` is absent on a bare graph (-> None), exactly what gcn.py:144-145 + :181 rely on.
    def __getattr__(self, name: str) -> Any:
        if name == "batch":
            return None
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '{name}'")

    @property
    def num_nodes(self) -> Optional[int]:
        if self.x is not None:
            return int(self.x.size(0))
        if self.edge_index is not None and self.edge_index.numel() > 0:
            return int(self.edge_index.max()) + 1
        return None

    @property
    def num_edges(self) -> int:
        return 0 if self.edge_index is None else int(self.edge_index.size(-1))

    @property
    def num_node_features(self) -> int:
        if self.x is None:
            return 0
        return 1 if self.x.dim() == 1 else int(self.x.size(-1))

    @property
    def num_edge_features(self) -> int:
        if self.edge_attr is None:
            return 0
        return 1 if self.edge_attr.dim() == 1 else int(self.edge_attr.size(-1))

    def keys(self) -> List[str]:
        return [k for k, v in self.__dict__.items() if not k.startswith("_") and v is not None]

    def to(self, device, non_blocking: bool = False) -> "Data":
        for k, v in list(self.__dict__.items()):
            if isinstance(v, Tensor):
                self.__dict__[k] = v.to(device, non_blocking=non_blocking)
        self.__dict__.pop("_gcb_packed", None)
        return self

    def cuda(self) -> "Data":
        return self.to("cuda")

    def cpu(self) -> "Data":
        return self.to("cpu")

    def __repr__(self) -> str:
        parts = []
        for k in self.keys():
            v = self.__dict__[k]
            parts.append(f"{k}={list(v.shape)}" if isinstance(v, Tensor) else f"{k}={v}")
        return f"{type(self).__name__}({', '.join(parts)})"


class Batch(Data):
    """Minimal `torch_geometric.data.Batch` (adds `batch`, `ptr`, `num_graphs`)."""

    @classmethod
    def from_data_list(cls, data_list: Sequence[Data]) -> "Batch":
        """PyG collate restated: keys in insertion order (x, edge_index, edge_attr, y;
        structs.py:61-66) are concatenated on dim 0, except `*index*` keys which are concatenated
        on dim -1 after adding the cumulative node count of the preceding graphs."""
        data_list = list(data_list)
        if len(data_list) == 0:
            raise ValueError("cannot collate an empty list of graphs")
        num_nodes = [int(d.x.size(0)) for d in data_list]
        offsets = [0]
        for n in num_nodes:
            offsets.append(offsets[-1] + n)
        out = cls()
        out.x = torch.cat([d.x for d in data_list], dim=0)
        out.edge_index = torch.cat([d.edge_index + off for d, off in zip(data_list, offsets[:-1])], dim=-1)
        out.edge_attr = torch.cat([d.edge_attr for d in data_list], dim=0)
        if all(getattr(d, "y", None) is not None for d in data_list):
            out.y = torch.cat([d.y for d in data_list], dim=0)
        dev = out.x.device
        out.batch = torch.repeat_interleave(torch.arange(len(data_list), device=dev),
                                            torch.tensor(num_nodes, dtype=torch.long, device=dev))
        out.ptr = torch.tensor(offsets, dtype=torch.long, device=dev)
        out.num_graphs = len(data_list)
        return out


class MoleculeGraph(Data):
    """Reference `graphchem/data/structs.py:7-66`: (atom_attr, bond_attr, connectivity, target) ->
    Data(x, edge_index, edge_attr, y) with y normalised to [1, n_targets]."""

    def __init__(self, atom_attr: Tensor, bond_attr: Tensor, connectivity: Tensor,
                 target: Optional[Tensor] = None):
        if target is None:
            target = torch.zeros(1, 1, dtype=torch.float32)          # structs.py:52-54
        elif target.dim() == 1:
            target = target.reshape(1, -1)                           # structs.py:55-57
        if target.shape[0] != 1:
            raise ValueError("Target tensor must have shape (1, num_targets)")  # structs.py:58-59
        super().__init__(x=atom_attr, edge_index=connectivity, edge_attr=bond_attr, y=target)


class MoleculeDataset(torch.utils.data.Dataset):
    """Reference `graphchem/data/structs.py:69-117`: list-backed dataset with PyG's `len()`/`get()`
    protocol; `dataset[i]` returns the stored MoleculeGraph, out-of-range raises IndexError."""

    def __init__(self, graphs: Iterable[MoleculeGraph]):
        super().__init__()
        self._graphs = list(graphs)

    def len(self) -> int:
        return len(self._graphs)

    def get(self, idx: int) -> MoleculeGraph:
        return self._graphs[idx]

    def __len__(self) -> int:
        return self.len()

    def __getitem__(self, idx: Union[int, slice, Sequence[int], Tensor]):
        if isinstance(idx, numbers.Integral) or (isinstance(idx, (Tensor, np.ndarray)) and idx.ndim == 0):
            return self.get(int(idx))
        if isinstance(idx, slice):
            return MoleculeDataset(self._graphs[idx])
        if isinstance(idx, np.ndarray):
            idx = torch.from_numpy(idx)
        if isinstance(idx, Tensor):
            idx = idx.nonzero().flatten().tolist() if idx.dtype == torch.bool else idx.tolist()
        return MoleculeDataset([self._graphs[int(i)] for i in idx])

    def __iter__(self) -> Iterator[MoleculeGraph]:
        return iter(self._graphs)


class DataLoader(torch.utils.data.DataLoader):
    """`torch_geometric.loader.DataLoader` stand-in: a torch DataLoader whose collate function is
    `Batch.from_data_list` (examples/predict_cn.ipynb:212-213)."""

    def __init__(self, dataset, batch_size: int = 1, shuffle: bool = False, **kwargs: Any):
        kwargs.pop("collate_fn", None)
        super().__init__(dataset, batch_size=batch_size, shuffle=shuffle, collate_fn=Batch.from_data_list, **kwargs)
