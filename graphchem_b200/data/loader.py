"""Packed-batch loader: the host side of training at scale.

The reference iterates a `torch_geometric.loader.DataLoader` whose collate is a Python loop per key per
graph (examples/predict_cn.ipynb:212, 245).  At millions of molecules per second on the GPU that loop -- and
even the one-C-call-per-batch packer, ~0.4 M molecules/s per host core -- is the bottleneck, so the loader

* packs batches on a pool of host threads (the C-ABI packer releases the GIL) into RECYCLED pinned buffers
  (a fresh 32 MB allocation per batch costs more in page faults / cudaHostAlloc than the packing itself),
* ships them on a copy stream, `prefetch` batches ahead of the consumer,
* keeps the pipeline full ACROSS epochs (`lookahead`, default on): the worker pool is persistent and, once the current
  epoch's last batch has been handed to a worker, the first batches of the next epoch (whose order is a function of
  `seed + epoch`) are packed ahead, so an epoch does not start with an empty pipeline (short epochs lost ~2.5 ms each),
* and, with `cache=True`, keeps the packed batches resident (in HBM when `device` is CUDA) after the first
  epoch: later epochs replay them (optionally in shuffled ORDER; batch composition is then fixed), which is
  what lets `TrainStep` reuse one CUDA graph per batch.  10 M molecules of the bench shape are 39 GB packed.

`indices` restricts the loader to a shard (data parallel: `shard_indices` in graphchem_b200.distributed).
"""
from __future__ import annotations

import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Iterator, List, Optional, Sequence

import numpy as np
import torch

from .packed import MoleculeStore, PackedBatch, RawBatch


class PackedLoader:
    def __init__(self, store: MoleculeStore, batch_size: int, n_messages: int, atom_vocab: int, bond_vocab: int,
                 device="cuda", shuffle: bool = False, drop_last: bool = False, seed: int = 0, workers: int = 8,
                 prefetch: int = 4, cache: bool = False, indices: Optional[Sequence[int]] = None, each: bool = False,
                 device_slots: int = 0, device_pack: bool = False, buffers: Optional[dict] = None, lookahead: bool = True):
        self.store, self.batch_size = store, int(batch_size)
        self.n_messages, self.av, self.bv = int(n_messages), int(atom_vocab), int(bond_vocab)
        self.device = torch.device(device)
        self.shuffle, self.drop_last, self.seed = shuffle, drop_last, seed
        self.workers, self.prefetch, self.cache, self.each = max(1, workers), max(1, prefetch), cache, each
        # device_slots > 0: batches land in a fixed ring of device buffers (instead of a fresh allocation each), so
        # consumers that key CUDA graphs on buffer addresses (TrainStep) replay instead of re-capturing; a yielded
        # batch is valid until `device_slots - 1` further batches have been requested.  Not combinable with cache.
        self.device_slots = 0 if cache else max(0, int(device_slots))
        self._slots: List[tuple] = []
        # device_pack: host threads only gather each batch into a raw buffer (`MoleculeStore.collate_raw`, a fifth
        # of the packing work and of the H2D bytes); the packed batch is derived on the GPU by `gcb_pack_device`
        # on the consumer's stream when the batch is handed out, bit-identical to the host packer's.  CUDA devices
        # only, not with `each`.
        self.device_pack = bool(device_pack)
        if self.device_pack and (each or torch.device(device).type != "cuda"):
            raise ValueError("device_pack needs a CUDA device and is not available for each=True batches")
        self.indices = np.arange(len(store), dtype=np.int64) if indices is None else np.asarray(indices, dtype=np.int64)
        self.epoch = 0
        self.lookahead = bool(lookahead) and not cache
        self._executor: Optional[ThreadPoolExecutor] = None
        self._ahead: List = []                       # futures of the NEXT epoch's first batches, in order
        self._ahead_batches: List[np.ndarray] = []   # the index sets they pack
        self._ahead_epoch = -1
        self._cached: Optional[List[PackedBatch]] = None
        self._pin = self.device.type == "cuda"
        self._pool: List[tuple] = []  # free (int32 buffer, float32 target buffer) pairs, pinned; CUDA devices only
        self._pool_lock = threading.Lock()
        self._pool_ready = False
        self._scratch: Optional[torch.Tensor] = None   # device scratch of the pack kernels, reused across batches
        if buffers:  # pinned pool / device slots of an earlier loader of the same geometry (`export_buffers`)
            self.adopt_buffers(buffers)

    def export_buffers(self) -> dict:
        """Pinned host buffers, device slots and pack scratch of this loader, for a later loader over another store
        with the same batch geometry (`buffers=`): cudaHostAlloc of the pool costs more than a whole epoch."""
        with self._pool_lock:
            return {"pool": list(self._pool), "slots": list(self._slots), "scratch": self._scratch,
                    "host_cap": self._capacity_host() if self._pool else 0,
                    "slot_cap": self._capacity() if self._slots else 0, "device": self.device,
                    "device_pack": self.device_pack, "n_pool": len(self._pool)}

    def adopt_buffers(self, b: dict) -> None:
        if b.get("device") != self.device or b.get("device_pack") != self.device_pack:
            return
        nt = 0 if self.store.targets is None else int(self.store.targets.size(1))
        y_cap = max(1, self.batch_size * nt)
        if b.get("pool") and b["host_cap"] >= self._capacity_host() and b["pool"][0][1].numel() >= y_cap:
            self._pool = list(b["pool"])
        if (b.get("slots") and self.device_slots and len(b["slots"]) == self.device_slots
                and b["slot_cap"] >= self._capacity() and b["slots"][0][1].numel() >= y_cap):
            self._slots = list(b["slots"])
        self._scratch = b.get("scratch")

    def __len__(self) -> int:
        n = len(self.indices)
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def _get_executor(self) -> ThreadPoolExecutor:
        if self._executor is None:
            self._executor = ThreadPoolExecutor(self.workers, thread_name_prefix="gcb-pack")
        return self._executor

    def close(self) -> None:
        """Stop the worker pool (batches packed ahead are dropped).  The loader can be iterated again afterwards."""
        ex, self._executor = self._executor, None
        self._ahead, self._ahead_batches, self._ahead_epoch = [], [], -1
        if ex is not None:
            ex.shutdown(wait=False, cancel_futures=True)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _recycle(self, futures) -> None:
        """Pinned buffers of batches that were packed but will not be consumed go back to the pool."""
        for f in futures:
            if f.cancel():
                continue
            try:
                _, bufs = f.result()
            except Exception:
                continue
            if bufs is not None:
                with self._pool_lock:
                    self._pool.append(bufs)

    def _take_ahead(self, batches: List[np.ndarray]) -> list:
        """Futures started during the previous epoch for the head of THIS epoch (checked against the index sets: a
        loader whose shuffle / indices changed in between packs them again)."""
        ahead, ab, ep = self._ahead, self._ahead_batches, self._ahead_epoch
        self._ahead, self._ahead_batches, self._ahead_epoch = [], [], -1
        if not ahead:
            return []
        ok = ep == self.epoch and len(ahead) <= len(batches) and all(np.array_equal(a, b) for a, b in zip(ab, batches))
        if ok:
            return list(ahead)
        self._recycle(ahead)
        return []

    def _batches(self, epoch: Optional[int] = None) -> List[np.ndarray]:
        idx = self.indices
        if self.shuffle:
            idx = idx[np.random.default_rng(self.seed + (self.epoch if epoch is None else epoch)).permutation(len(idx))]
        out = [idx[i:i + self.batch_size] for i in range(0, len(idx), self.batch_size)]
        if self.drop_last and out and len(out[-1]) < self.batch_size:
            out.pop()
        return out

    def _pack(self, idx: np.ndarray):
        """-> (host PackedBatch, pooled buffers or None).  On CUDA devices the batch is packed into a recycled
        pinned buffer pair sized for this loader's largest batch."""
        if not self._pin:
            return self.store.collate_pack(idx, self.n_messages, self.av, self.bv, pin=False, each=self.each), None
        with self._pool_lock:
            bufs = self._pool.pop() if self._pool else None
        if bufs is None:
            nt = 0 if self.store.targets is None else int(self.store.targets.size(1))
            bufs = (torch.empty(self._capacity_host(), dtype=torch.int32, pin_memory=True),
                    torch.empty(max(1, self.batch_size * nt), dtype=torch.float32, pin_memory=True))
        if self.device_pack:
            hb = self.store.collate_raw(idx, self.n_messages, self.av, self.bv, pin=True, out=bufs[0], y_out=bufs[1])
        else:
            hb = self.store.collate_pack(idx, self.n_messages, self.av, self.bv, pin=True, each=self.each, out=bufs[0],
                                         y_out=bufs[1])
        return hb, bufs

    def _capacity(self) -> int:
        """int32 elements that hold any batch of this loader (packed size grows with atoms and edges)."""
        from .. import _lib
        from .packed import _header
        na = np.sort(self.store.n_atoms[self.indices])[::-1][:self.batch_size].sum()
        ne = np.sort(self.store.n_edges[self.indices])[::-1][:self.batch_size].sum()
        N, E = int(na), int(max(ne, na if self.each else 0))
        M = min(N, E) if self.n_messages >= 1 else E
        return _header(N, E, self.batch_size, M, self.av, self.bv, _lib.PACK_EACH if self.each else 0)[1]

    def _capacity_host(self) -> int:
        """int32 elements of a pinned host buffer: the raw batch when the device packs (a sixth of the packed size)."""
        if not self.device_pack:
            return self._capacity()
        from .. import _lib
        na = np.sort(self.store.n_atoms[self.indices])[::-1][:self.batch_size].sum()
        ne = np.sort(self.store.n_edges[self.indices])[::-1][:self.batch_size].sum()
        return int(_lib.check(int(_lib.lib().gcb_raw_ints(int(na), int(ne), self.batch_size)), "gcb_raw_ints"))

    def __iter__(self) -> Iterator[PackedBatch]:
        self.epoch += 1
        if self._cached is not None:  # resident batches: replay, in shuffled order if asked
            order = np.arange(len(self._cached))
            if self.shuffle:
                order = np.random.default_rng(self.seed + self.epoch).permutation(len(order))
            for i in order:
                yield self._cached[int(i)]
            return
        batches = self._batches()
        keep: List[PackedBatch] = []
        cuda = self.device.type == "cuda"
        copy_stream = torch.cuda.Stream(self.device) if cuda else None
        pending = self._take_ahead(batches)  # futures of host batches, in order (the head may have been packed ahead)
        if cuda:
            # all pinned buffers up front (topping up an adopted pool): cudaHostAlloc is slow and synchronises the
            # device, it must not happen in the middle of an epoch
            nt = 0 if self.store.targets is None else int(self.store.targets.size(1))
            with self._pool_lock:
                # (batches packed ahead during the previous epoch hold one pinned pair each; with lookahead the pipeline
                # stays `prefetch + workers` deep across the epoch boundary, whatever the epoch length -- a worker that
                # finds the pool empty would call cudaHostAlloc, which synchronises the device and INVALIDATES a CUDA graph
                # capture running on the consumer's thread)
                cap = self.prefetch + self.workers
                missing = (cap if self.lookahead else min(len(batches), cap)) - len(self._pool) - len(pending)
                if missing > 0:
                    cap = self._capacity_host()
                    for _ in range(missing):
                        self._pool.append((torch.empty(cap, dtype=torch.int32, pin_memory=True),
                                           torch.empty(max(1, self.batch_size * nt), dtype=torch.float32, pin_memory=True)))
            self._pool_ready = True
        if cuda and self.device_slots and not self._slots:
            nt = 0 if self.store.targets is None else int(self.store.targets.size(1))
            cap = self._capacity()
            self._slots = [(torch.empty(cap, dtype=torch.int32, device=self.device),
                            torch.empty(max(1, self.batch_size * nt), dtype=torch.float32, device=self.device),
                            torch.cuda.Event()) for _ in range(self.device_slots)]
            for _, _, ev0 in self._slots:
                ev0.record(torch.cuda.current_stream(self.device))
        slot_i = 0
        pool = self._get_executor()
        inflight = []  # (device batch, copy-done event)
        nxt = len(pending)
        try:
            def submit():
                nonlocal nxt
                cap = self.prefetch + self.workers
                while nxt < len(batches) and len(pending) + len(inflight) < cap:
                    pending.append(pool.submit(self._pack, batches[nxt]))
                    nxt += 1
                if self.lookahead and nxt >= len(batches):  # this epoch is fully submitted: start on the next one
                    if self._ahead_epoch != self.epoch + 1:
                        self._ahead_batches, self._ahead_epoch = self._batches(self.epoch + 1), self.epoch + 1
                    while (len(self._ahead) < len(self._ahead_batches)
                           and len(pending) + len(inflight) + len(self._ahead) < cap):
                        self._ahead.append(pool.submit(self._pack, self._ahead_batches[len(self._ahead)]))

            submit()
            while pending or inflight:
                # with a slot ring, a slot may only be refilled after its previous batch was consumed: at most
                # device_slots - 1 copies ahead of the consumer
                ahead = min(self.prefetch, len(self._slots) - 1) if self._slots else self.prefetch
                while pending and len(inflight) < max(1, ahead) and (pending[0].done() or not inflight):
                    hb, bufs = pending.pop(0).result()
                    if cuda:
                        with torch.cuda.stream(copy_stream):
                            if self._slots:
                                d_ints, d_y, free_ev = self._slots[slot_i % len(self._slots)]
                                slot_i += 1
                                copy_stream.wait_event(free_ev)  # the consumer is done with this slot's previous batch
                                y = None
                                if hb.y is not None:
                                    y = d_y[:hb.y.numel()].view(hb.y.shape)
                                    y.copy_(hb.y, non_blocking=True)
                                if self.device_pack:  # raw H2D here; packed into the slot on the consumer's stream
                                    db = (RawBatch(hb.ints.to(self.device, non_blocking=True), hb.raw_header, hb.header,
                                                   y), d_ints)
                                else:
                                    n = hb.ints.numel()
                                    hb.upload_into(d_ints)  # compact prefix over PCIe, ELL-4 views rebuilt on the device
                                    db = PackedBatch(d_ints[:n], hb.header, y)
                            else:
                                db = hb.to(self.device, non_blocking=True)
                                if self.device_pack:
                                    db = (db, None)
                            ev = torch.cuda.Event()
                            ev.record(copy_stream)
                        inflight.append((db, ev, bufs))
                    else:
                        inflight.append((hb, None, None))
                    submit()
                db, ev, bufs = inflight.pop(0)
                if ev is not None:
                    consumer = torch.cuda.current_stream(self.device)
                    consumer.wait_event(ev)
                    if self.device_pack:
                        # The pack kernels run on the CONSUMER's stream, in front of its step, not on the copy stream:
                        # running next to a train step they cost more than they hide (measured: 2.68 ms per batch with
                        # the packer on the copy stream vs 2.12 ms with the host packer, profiles/micro/stream_modes.py;
                        # presumably their blocks hold SM resources that the persistent one-CTA-per-SM GEMMs need, and
                        # a GEMM waits for its slowest SM).
                        raw, slot_ints = db
                        raw.ints.record_stream(consumer)  # allocated on the copy stream, read by the pack kernels here
                        if self._scratch is None:
                            from .. import _lib
                            n_scr = int(_lib.check(int(_lib.lib().gcb_pack_device_scratch_ints(raw.header.ctypes.data)),
                                                   "gcb_pack_device_scratch_ints"))
                            self._scratch = torch.empty(int(n_scr * 1.25) + 1024, dtype=torch.int32, device=self.device)
                        try:
                            db = raw.pack(out=slot_ints, scratch=self._scratch)
                        except ValueError:  # a larger batch than the scratch was sized for
                            self._scratch = None
                            db = raw.pack(out=slot_ints)
                    if not self._slots:
                        # the batch was allocated on the copy stream but is read on the consumer's stream: without
                        # this the caching allocator hands its memory to a later copy as soon as the consumer drops
                        # the batch, while the consumer's kernels may still be reading it (garbage indices)
                        db.ints.record_stream(consumer)
                        if db.y is not None:
                            db.y.record_stream(consumer)
                    if bufs is not None:  # the pinned pair is free again once its copy has left the host
                        ev.synchronize()
                        with self._pool_lock:
                            self._pool.append(bufs)
                if self.cache:
                    keep.append(db)
                yield db
                if self._slots:  # whatever the consumer enqueued on its stream for this batch frees the slot
                    k = (slot_i - len(inflight) - 1) % len(self._slots)
                    self._slots[k][2].record(torch.cuda.current_stream(self.device))
                submit()
        finally:
            if pending:  # the consumer left the epoch early
                self._recycle(pending)
        self._ahead_batches = self._ahead_batches[:len(self._ahead)]
        if self.cache:
            self._cached = keep
