"""Pre-packed batches: the host-side half of the hot path.

A `PackedBatch` is ONE int32 buffer (tokens, edge list, destination- and source-sorted CSR, graph
segments, token buckets; layout in include/graphchem_b200.h) plus the float targets.  It is built
on host cores by the C-ABI packer (`gcb_pack_host` / `gcb_collate_pack_host`), normally into pinned
memory, and shipped to HBM with a single copy.  This is the "RDKit featurisation stays on host
cores, is prepacked into pinned CSR buffers" step of the north star; its index semantics restate
PyG's collate and scatter order (reference call site
examples/predict_cn.ipynb:212,245).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch
from torch import Tensor

from .. import _lib


def live_bond_rows(n_nodes: int, n_edges: int, n_messages: int) -> int:
    """Rows of the bond matrix that carry information: min(N, E) once EdgeConv has run on the bond
    matrix indexed by atom ids (reference gcn.py:163), all E rows otherwise."""
    return min(n_nodes, n_edges) if n_messages >= 1 else n_edges


class PackedBatch:
    """Packed batch; `ints` lives on the host (optionally pinned) or on a CUDA device."""

    __slots__ = ("ints", "header", "y", "N", "E", "G", "M")

    def __init__(self, ints: Tensor, header: np.ndarray, y: Optional[Tensor]):
        self.ints, self.header, self.y = ints, header, y
        self.N, self.E, self.G, self.M = (int(header[_lib.H_N]), int(header[_lib.H_E]), int(header[_lib.H_G]),
                                          int(header[_lib.H_M]))

    @property
    def n_tiles(self) -> int:
        """Closed row tiles of at most 128 rows (no edge leaves a tile); 0 = the batch has a component
        wider than a tile and runs on the unfused kernels."""
        return int(self.header[_lib.H_NTILES])

    @property
    def device(self) -> torch.device:
        return self.ints.device

    @property
    def nbytes(self) -> int:
        return self.ints.numel() * 4 + (0 if self.y is None else self.y.numel() * self.y.element_size())

    @property
    def compact_ints(self) -> int:
        """Leading ints that have to cross PCIe: the ELL-4 views behind them are derived on the device."""
        return int(self.header[_lib.H_COMPACT_INTS])

    def upload_into(self, dst: Tensor, non_blocking: bool = True) -> None:
        """Host -> device into a preallocated int32 buffer of at least `ints.numel()` elements, on the current stream:
        copies the compact prefix and lets `gcb_pack_expand_ell` rebuild the ELL-4 views (bit-identical to the host's)."""
        n, total = self.compact_ints, self.ints.numel()
        # small batches (the notebooks' 16 molecules: ~40 kB) go over in one copy: the rebuild kernel would cost more
        # than the bytes it saves
        if dst.device.type != "cuda" or self.ints.device.type != "cpu" or n <= 0 or n >= total or total < (1 << 18):
            dst[:total].copy_(self.ints, non_blocking=non_blocking)
            return
        dst[:n].copy_(self.ints[:n], non_blocking=non_blocking)
        with torch.cuda.device(dst.device):
            _lib.check(_lib.lib().gcb_pack_expand_ell(dst.data_ptr(), self.header.ctypes.data,
                                                      torch.cuda.current_stream(dst.device).cuda_stream), "gcb_pack_expand_ell")

    def to(self, device, non_blocking: bool = False) -> "PackedBatch":
        device = torch.device(device)
        y = None if self.y is None else self.y.to(device, non_blocking=non_blocking)
        if device.type == "cuda" and self.ints.device.type == "cpu":
            ints = torch.empty(self.ints.numel(), dtype=torch.int32, device=device)
            self.upload_into(ints, non_blocking=non_blocking)
            return PackedBatch(ints, self.header, y)
        return PackedBatch(self.ints.to(device, non_blocking=non_blocking), self.header, y)

    def section(self, field: int, count: int) -> Tensor:
        off = int(self.header[field])
        return self.ints[off:off + count]

    def sections(self) -> dict:
        """All index arrays as int32 tensors (views), for tests and debugging."""
        N, E, G, M = self.N, self.E, self.G, self.M
        av, bv = int(self.header[_lib.H_AV]), int(self.header[_lib.H_BV])
        L = _lib
        return dict(
            x_tok=self.section(L.H_X_TOK, N), e_tok=self.section(L.H_E_TOK, E), src=self.section(L.H_SRC, E),
            dst=self.section(L.H_DST, E), in_ptr=self.section(L.H_IN_PTR, N + 1), in_eid=self.section(L.H_IN_EID, E),
            in_src=self.section(L.H_IN_SRC, E), out_ptr=self.section(L.H_OUT_PTR, N + 1),
            out_eid=self.section(L.H_OUT_EID, E), out_dst=self.section(L.H_OUT_DST, E),
            graph_ptr=self.section(L.H_GRAPH_PTR, G + 1), graph_node=self.section(L.H_GRAPH_NODE, N),
            atok_ptr=self.section(L.H_ATOK_PTR, av + 1), atok_node=self.section(L.H_ATOK_NODE, N),
            btok_ptr=self.section(L.H_BTOK_PTR, bv + 1), btok_row=self.section(L.H_BTOK_ROW, M),
            node_graph=self.section(L.H_NODE_GRAPH, N), in_src4=self.section(L.H_IN_SRC4, 4 * N),
            in_eid4=self.section(L.H_IN_EID4, 4 * N), out_dst4=self.section(L.H_OUT_DST4, 4 * N),
            out_inpos4=self.section(L.H_OUT_INPOS4, 4 * N), out_inpos=self.section(L.H_OUT_INPOS, E),
            tile_ptr=self.section(L.H_TILE_PTR, self.n_tiles + 1),
            **({"bond_pos": self.section(L.H_BOND_POS, E)} if int(self.header[L.H_FLAGS]) & L.PACK_EACH else {}))


class RawBatch:
    """Collated but not yet packed batch (include/graphchem_b200.h "device-side packer"): molecule offsets,
    tokens, edge list and tile table in ONE int32 buffer of ~0.7 kB per molecule, plus the header the packed
    batch will have.  `pack()` derives the packed batch on the device (`gcb_pack_device`), bit-identical
    to what `MoleculeStore.collate_pack` builds on host cores."""

    __slots__ = ("ints", "raw_header", "header", "y")

    def __init__(self, ints: Tensor, raw_header: np.ndarray, header: np.ndarray, y: Optional[Tensor]):
        self.ints, self.raw_header, self.header, self.y = ints, raw_header, header, y

    @property
    def device(self) -> torch.device:
        return self.ints.device

    @property
    def nbytes(self) -> int:
        return self.ints.numel() * 4 + (0 if self.y is None else self.y.numel() * self.y.element_size())

    def to(self, device, non_blocking: bool = False) -> "RawBatch":
        y = None if self.y is None else self.y.to(device, non_blocking=non_blocking)
        return RawBatch(self.ints.to(device, non_blocking=non_blocking), self.raw_header, self.header, y)

    def pack(self, out: Optional[Tensor] = None, scratch: Optional[Tensor] = None) -> PackedBatch:
        """Packed batch on this batch's CUDA device, built by kernels on torch's current stream.  `out` /
        `scratch`: device int32 buffers to reuse (large enough), else fresh ones.  No CPU fallback."""
        if self.ints.device.type != "cuda":
            raise RuntimeError("RawBatch.pack runs on the GPU: move the batch with .to('cuda') first "
                               "(on host cores use MoleculeStore.collate_pack)")
        dev = self.ints.device
        total = int(self.header[_lib.H_TOTAL_INTS])
        n_scr = int(_lib.check(int(_lib.lib().gcb_pack_device_scratch_ints(self.header.ctypes.data)),
                               "gcb_pack_device_scratch_ints"))
        with torch.cuda.device(dev):
            if out is None:
                out = torch.empty(total, dtype=torch.int32, device=dev)
            if scratch is None:
                scratch = torch.empty(n_scr, dtype=torch.int32, device=dev)
            if out.numel() < total or scratch.numel() < n_scr or out.dtype != torch.int32 or scratch.dtype != torch.int32:
                raise ValueError("RawBatch.pack: `out` / `scratch` must be int32 device buffers with room for the batch")
            rc = _lib.lib().gcb_pack_device(self.ints.data_ptr(), self.raw_header.ctypes.data, self.header.ctypes.data,
                                            out.data_ptr(), out.numel(), scratch.data_ptr(), scratch.numel(),
                                            torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(rc, "gcb_pack_device")
        return PackedBatch(out[:total], self.header.copy(), self.y)


def _alloc_ints(n: int, pin: bool) -> Tensor:
    pin = pin and torch.cuda.is_available()
    return torch.empty(n, dtype=torch.int32, pin_memory=pin)


def _header(N: int, E: int, G: int, M: int, av: int, bv: int, flags: int = 0):
    header = np.zeros(_lib.PACK_HEADER_INTS, dtype=np.int32)
    total = _lib.lib().gcb_pack_ints_flags(N, E, G, M, av, bv, flags, header.ctypes.data)
    _lib.check(int(total), "gcb_pack_ints")
    return header, int(total)


def _host_index_tensor(t: Tensor, name: str) -> Tensor:
    if t.dtype not in (torch.int64, torch.int32):
        if t.dtype in (torch.int16, torch.int8, torch.uint8):
            t = t.to(torch.int64)
        else:
            # nn.Embedding / PyG _check_input reject non-integer indices
            raise (ValueError if name == "edge_index" else RuntimeError)(
                f"Expected '{name}' to be of integer type (got '{t.dtype}')")
    return t.detach().to("cpu").contiguous()


def pack_batch(data, n_messages: int, atom_vocab: int, bond_vocab: int, pin: bool = True) -> PackedBatch:
    """Pack a collated PyG-style Batch (or a single graph with `batch=None`)."""
    x, edge_attr, edge_index = data.x, data.edge_attr, data.edge_index
    batch = getattr(data, "batch", None)
    if x is None:
        # reference gcn.py:148-149 would feed float ones to nn.Embedding, which raises
        raise RuntimeError("MoleculeGCN needs integer atom tokens in data.x (nn.Embedding indices)")
    if edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise ValueError(f"Expected 'edge_index' to be of shape [2, num_edges] (got {list(edge_index.shape)})")
    if x.dim() != 1 or edge_attr.dim() != 1:
        raise ValueError("x and edge_attr must be 1-D integer token tensors")
    N, E = int(x.size(0)), int(edge_index.size(1))
    if int(edge_attr.size(0)) != E:
        raise ValueError("edge_attr and edge_index disagree on the number of edges")
    x_h = _host_index_tensor(x, "x").to(torch.int64)
    ea_h = _host_index_tensor(edge_attr, "edge_attr").to(torch.int64)
    ei_h = _host_index_tensor(edge_index, "edge_index").to(torch.int64)
    b_h = None
    if batch is not None:
        b_h = _host_index_tensor(batch, "batch").to(torch.int64)
        if b_h.numel() != N:
            raise ValueError("batch must have one entry per node")
        G = getattr(data, "num_graphs", None)
        G = int(G) if isinstance(G, int) else (int(b_h.max()) + 1 if N > 0 else 0)
    else:
        G = 1
    M = live_bond_rows(N, E, n_messages)
    header, total = _header(N, E, G, M, atom_vocab, bond_vocab)
    ints = _alloc_ints(total, pin)
    rc = _lib.lib().gcb_pack_host(x_h.data_ptr(), ea_h.data_ptr(), ei_h.data_ptr(),
                                  b_h.data_ptr() if b_h is not None else None, 8, N, E, G, M, atom_vocab, bond_vocab,
                                  1 if n_messages >= 1 else 0, ints.data_ptr(), total)
    _lib.check(rc, "gcb_pack_host")
    header[:] = ints[:_lib.PACK_HEADER_INTS].numpy()  # the packer fills the data-dependent fields (tile count)
    y = getattr(data, "y", None)
    y = None if y is None else y.detach().to("cpu", torch.float32).contiguous()
    return PackedBatch(ints, header, y)


class MoleculeStore:
    """Flat host storage of many encoded molecules (int32 atoms / int32 bonds / int64 [2,e]
    connectivity, i.e. the MoleculeEncoder.encode layout of reference features.py:263-321), laid
    out so that the C-ABI collate can address any subset by pointer arithmetic."""

    def __init__(self, atoms: Sequence[Tensor], bonds: Sequence[Tensor], conns: Sequence[Tensor],
                 targets: Optional[Tensor] = None):
        n = len(atoms)
        self.n_atoms = np.array([int(a.numel()) for a in atoms], dtype=np.int32)
        self.n_edges = np.array([int(b.numel()) for b in bonds], dtype=np.int32)
        self.atoms = torch.cat([a.reshape(-1).to(torch.int32) for a in atoms]) if n else torch.zeros(0, dtype=torch.int32)
        self.bonds = torch.cat([b.reshape(-1).to(torch.int32) for b in bonds]) if n else torch.zeros(0, dtype=torch.int32)
        self.conn = torch.cat([c.to(torch.int64).contiguous().reshape(-1) for c in conns]) if n else torch.zeros(0, dtype=torch.int64)
        self.atom_off = np.concatenate([[0], np.cumsum(self.n_atoms, dtype=np.int64)])
        self.edge_off = np.concatenate([[0], np.cumsum(self.n_edges, dtype=np.int64)])
        self.targets = None if targets is None else targets.to(torch.float32).reshape(n, -1).contiguous()

    @classmethod
    def from_flat(cls, atoms: Tensor, bonds: Tensor, conn: Tensor, n_atoms: np.ndarray, n_edges: np.ndarray,
                  targets: Optional[Tensor] = None) -> "MoleculeStore":
        """Store over already-concatenated arrays: `atoms` int32 [sum n_atoms], `bonds` int32 [sum n_edges],
        `conn` int64 [2 * sum n_edges] (per molecule: its e sources, then its e destinations, molecule-local ids,
        i.e. each molecule's [2, e] connectivity flattened).  No per-molecule Python objects: the form a
        10 M-molecule shard is built in."""
        self = cls.__new__(cls)
        self.n_atoms = np.ascontiguousarray(n_atoms, dtype=np.int32)
        self.n_edges = np.ascontiguousarray(n_edges, dtype=np.int32)
        n = int(self.n_atoms.shape[0])
        self.atoms = atoms.reshape(-1).to(torch.int32).contiguous()
        self.bonds = bonds.reshape(-1).to(torch.int32).contiguous()
        self.conn = conn.reshape(-1).to(torch.int64).contiguous()
        self.atom_off = np.concatenate([[0], np.cumsum(self.n_atoms, dtype=np.int64)])
        self.edge_off = np.concatenate([[0], np.cumsum(self.n_edges, dtype=np.int64)])
        if (int(self.atom_off[-1]) != self.atoms.numel() or int(self.edge_off[-1]) != self.bonds.numel()
                or 2 * int(self.edge_off[-1]) != self.conn.numel() or self.n_edges.shape[0] != n):
            raise ValueError("MoleculeStore.from_flat: array lengths disagree with the per-molecule counts")
        self.targets = None if targets is None else targets.to(torch.float32).reshape(n, -1).contiguous()
        return self

    @classmethod
    def from_graphs(cls, graphs: Sequence) -> "MoleculeStore":
        ys = [g.y for g in graphs]
        targets = torch.cat(ys, dim=0) if all(y is not None for y in ys) and len(ys) else None
        return cls([g.x for g in graphs], [g.edge_attr for g in graphs], [g.edge_index for g in graphs], targets)

    def __len__(self) -> int:
        return int(self.n_atoms.shape[0])

    def _pointer_tables(self, idx: np.ndarray):
        atoms_p = (self.atoms.data_ptr() + 4 * self.atom_off[idx]).astype(np.uint64)
        bonds_p = (self.bonds.data_ptr() + 4 * self.edge_off[idx]).astype(np.uint64)
        conn_p = (self.conn.data_ptr() + 16 * self.edge_off[idx]).astype(np.uint64)
        return atoms_p, bonds_p, conn_p

    def collate_raw(self, idx: Sequence[int], n_messages: int, atom_vocab: int, bond_vocab: int,
                    pin: bool = True, out: Optional[Tensor] = None, y_out: Optional[Tensor] = None) -> RawBatch:
        """Collate molecules `idx` into a `RawBatch` (one C call: validation, gather, tile table); the device
        derives the packed batch from it (`RawBatch.to(device).pack()`).  A fifth of the host work and of the
        H2D bytes of `collate_pack`.  `out` / `y_out` as in `collate_pack`."""
        idx = np.asarray(idx, dtype=np.int64)
        G = int(idx.shape[0])
        na, ne = np.ascontiguousarray(self.n_atoms[idx]), np.ascontiguousarray(self.n_edges[idx])
        N, E = int(na.sum()), int(ne.sum())
        total = int(_lib.check(int(_lib.lib().gcb_raw_ints(N, E, G)), "gcb_raw_ints"))
        if out is not None and (out.dtype != torch.int32 or out.device.type != "cpu" or out.numel() < total):
            raise ValueError("collate_raw: `out` must be a host int32 tensor with room for the raw batch")
        ints = out[:total] if out is not None else _alloc_ints(total, pin)
        atoms_p, bonds_p, conn_p = self._pointer_tables(idx)
        nt = 0 if self.targets is None else int(self.targets.size(1))
        y, y_p, yo_p, y_ptrs = None, None, None, None
        if nt:
            if y_out is not None and y_out.numel() >= G * nt and y_out.dtype == torch.float32:
                y = y_out.reshape(-1)[:G * nt].view(G, nt)
            else:
                y = torch.empty(G, nt, dtype=torch.float32, pin_memory=pin and torch.cuda.is_available())
            y_ptrs = (self.targets.data_ptr() + 4 * nt * idx).astype(np.uint64)  # kept alive across the call
            y_p, yo_p = y_ptrs.ctypes.data, y.data_ptr()
        header = np.zeros(_lib.PACK_HEADER_INTS, dtype=np.int32)
        rc = _lib.lib().gcb_collate_raw_host(atoms_p.ctypes.data, na.ctypes.data, bonds_p.ctypes.data, conn_p.ctypes.data,
                                             ne.ctypes.data, y_p, nt, G, n_messages, atom_vocab, bond_vocab,
                                             ints.data_ptr(), total, yo_p, header.ctypes.data)
        _lib.check(rc, "gcb_collate_raw_host")
        raw_header = ints[:_lib.RAW_HEADER_INTS].numpy().copy()
        return RawBatch(ints, raw_header, header, y)

    def collate_pack_epoch(self, batches: Sequence[Sequence[int]], n_messages: int, atom_vocab: int, bond_vocab: int,
                           device) -> List[PackedBatch]:
        """All batches of an epoch (index sets, in order) packed into ONE pinned buffer and shipped with ONE host-to-device
        copy each for the indices and the targets; returns device-resident PackedBatches that are views of one device
        buffer (valid until the next-but-one call: two pinned / device buffer sets alternate, so the CPU can pack the
        next epoch while the GPU still trains on this one).  For small data sets trained in small batches -- the
        notebooks' 368 molecules in batches of 16, ~1 MB per epoch -- together with `TrainStep.run_epoch`."""
        device = torch.device(device)
        if device.type != "cuda":
            raise ValueError("collate_pack_epoch ships to a CUDA device")
        nt = 0 if self.targets is None else int(self.targets.size(1))
        sizes, offs, yoffs, total, ytotal = [], [], [], 0, 0
        for idx in batches:
            idx = np.asarray(idx, dtype=np.int64)
            N, E = int(self.n_atoms[idx].sum()), int(self.n_edges[idx].sum())
            n_ints = _header(N, E, len(idx), live_bond_rows(N, E, n_messages), atom_vocab, bond_vocab, 0)[1]
            offs.append(total); sizes.append(n_ints)
            total += (n_ints + 63) & ~63          # every batch starts 256-byte aligned (the kernels' 16-byte vector loads)
            yoffs.append(ytotal)
            ytotal += (len(idx) * nt + 3) & ~3
        st = getattr(self, "_epoch_state", None)
        if st is None or st["device"] != device:
            st = self._epoch_state = {"device": device, "flip": 0, "sets": [None, None]}
        st["flip"] ^= 1
        cur = st["sets"][st["flip"]]
        if cur is None or cur["ints"].numel() < total or cur["y"].numel() < max(1, ytotal):
            cap, ycap = int(total * 1.25) + 1024, int(max(1, ytotal) * 1.25) + 16
            cur = st["sets"][st["flip"]] = {
                "ints": torch.empty(cap, dtype=torch.int32, pin_memory=True), "y": torch.empty(ycap, dtype=torch.float32, pin_memory=True),
                "d_ints": torch.empty(cap, dtype=torch.int32, device=device), "d_y": torch.empty(ycap, dtype=torch.float32, device=device),
                "ev": None}
        if cur["ev"] is not None:
            cur["ev"].synchronize()  # the copy that last read this pinned pair (two epochs ago) has left the host
        hosts = []
        for idx, off, n_ints, yo in zip(batches, offs, sizes, yoffs):
            hosts.append(self.collate_pack(idx, n_messages, atom_vocab, bond_vocab, pin=False, out=cur["ints"][off:off + n_ints],
                                           y_out=cur["y"][yo:] if nt else None))
        with torch.cuda.device(device):
            cur["d_ints"][:total].copy_(cur["ints"][:total], non_blocking=True)
            if nt:
                cur["d_y"][:ytotal].copy_(cur["y"][:ytotal], non_blocking=True)
            cur["ev"] = torch.cuda.Event()
            cur["ev"].record(torch.cuda.current_stream(device))
        out = []
        for hb, off, n_ints, yo in zip(hosts, offs, sizes, yoffs):
            y = cur["d_y"][yo:yo + hb.y.numel()].view(hb.y.shape) if hb.y is not None else None
            out.append(PackedBatch(cur["d_ints"][off:off + n_ints], hb.header, y))
        return out

    def collate_pack(self, idx: Sequence[int], n_messages: int, atom_vocab: int, bond_vocab: int,
                     pin: bool = True, each: bool = False, out: Optional[Tensor] = None,
                     y_out: Optional[Tensor] = None) -> PackedBatch:
        """Collate molecules `idx` (in that order) and pack them: one C call, no Python per-graph loop.
        `each=True` lays the batch out so that every molecule is computed as if it were alone in the call
        (include/graphchem_b200.h GCB_PACK_EACH; prediction only).  `out` / `y_out`: host buffers to pack into
        (int32 / float32, large enough; a loader recycles pinned buffers through these) instead of fresh ones."""
        idx = np.asarray(idx, dtype=np.int64)
        G = int(idx.shape[0])
        na, ne = self.n_atoms[idx], self.n_edges[idx]
        N, E = int(na.sum()), int(ne.sum())
        if each and (n_messages < 1 or E < N):
            raise NotImplementedError("each=True needs n_messages >= 1 and at least as many edges as atoms in the batch")
        M = N if each else live_bond_rows(N, E, n_messages)
        header, total = _header(N, E, G, M, atom_vocab, bond_vocab, _lib.PACK_EACH if each else 0)
        if out is not None and (out.dtype != torch.int32 or out.device.type != "cpu" or out.numel() < total):
            raise ValueError("collate_pack: `out` must be a host int32 tensor with room for the packed batch")
        ints = out[:total] if out is not None else _alloc_ints(total, pin)
        atoms_p = (self.atoms.data_ptr() + 4 * self.atom_off[idx]).astype(np.uint64)
        bonds_p = (self.bonds.data_ptr() + 4 * self.edge_off[idx]).astype(np.uint64)
        conn_p = (self.conn.data_ptr() + 16 * self.edge_off[idx]).astype(np.uint64)
        nt = 0 if self.targets is None else int(self.targets.size(1))
        y = None
        y_p, yo_p, y_ptrs = None, None, None
        if nt:
            if y_out is not None and y_out.numel() >= G * nt and y_out.dtype == torch.float32:
                y = y_out.reshape(-1)[:G * nt].view(G, nt)
            else:
                y = torch.empty(G, nt, dtype=torch.float32, pin_memory=pin and torch.cuda.is_available())
            y_ptrs = (self.targets.data_ptr() + 4 * nt * idx).astype(np.uint64)  # kept alive across the call
            y_p = y_ptrs.ctypes.data
            yo_p = y.data_ptr()
        na = np.ascontiguousarray(na)
        ne = np.ascontiguousarray(ne)
        fn = _lib.lib().gcb_collate_pack_each_host if each else _lib.lib().gcb_collate_pack_host
        rc = fn(atoms_p.ctypes.data, na.ctypes.data, bonds_p.ctypes.data, conn_p.ctypes.data, ne.ctypes.data, y_p, nt, G,
                n_messages, atom_vocab, bond_vocab, ints.data_ptr(), total, yo_p)
        _lib.check(rc, "gcb_collate_pack_host")
        header[:] = ints[:_lib.PACK_HEADER_INTS].numpy()
        return PackedBatch(ints, header, y)
