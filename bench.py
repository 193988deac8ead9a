#!/usr/bin/env python
"""bench.py -- MoleculeGCN molecules/s on B200 (BASELINE.json metric), with the HBM roofline of the dominant
kernel, the end-to-end number through the public API with host buffers, and the reference's CPU path timed
beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config train|infer] [--batch B] [--molecules M]
                    [--impl ours|reference]

N > 1 is launched by the driver with torchrun (one rank per GPU, NCCL); RANK / LOCAL_RANK / WORLD_SIZE /
MASTER_* come from the environment.  Rank 0 prints ONE JSON line.

--config train (default; BASELINE.json configs[3] at 1..8 GPUs, configs[2] shape with --batch 4096): synthetic
  ZINC-scale molecules, 23 atoms / 50 directed edges each, 8192 molecules per GPU per step (weak scaling),
  embedding_dim 128, n_messages 4, n_readout 2, readout_dim 64, bf16 storage / fp32 accumulate, MSE + Adam.
  By default 8 distinct pre-packed batches per GPU are resident in HBM and cycled (per-step working set
  ~1.3 GB, far above the 126 MB L2).  --molecules M makes the resident set configs[3]'s own: a shard of
  M / n_gpus molecules per GPU (10 000 000 -> 1220 batches, 39 GB packed on one GPU), collated raw on the host,
  packed by the device packer at set-up, one CUDA graph per batch, K >= 200 steps.
--config infer (BASELINE.json configs[4]): forward only, embedding_dim 256, n_messages 6, 32768 molecules per
  step per GPU, replicas only (no collective); `value` = resident batches through `InferStep`, `e2e` = host
  molecules through `graphchem_b200.train.screen` (raw collate on host threads -> H2D -> device pack -> forward
  -> predictions back to the host).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

ATOMS, EDGES = 23, 50
CFGS = {
    "train": dict(atom_vocab=64, bond_vocab=32, out_dim=1, D=128, L=4, n_readout=2, R=64, batch=8192),
    "infer": dict(atom_vocab=64, bond_vocab=32, out_dim=1, D=256, L=6, n_readout=2, R=64, batch=32768),
}
N_RESIDENT = 8
CPU_SAMPLE_BATCH = 4096     # one 4096-molecule batch, >= 3 warm-up + >= 5 timed steps


# ------------------------------------------------------------------------------------------------
# algorithmic bytes
def fwd_bytes(N, E, D, L, s):
    return s * D * (5 * N * L + E - 2 * N) + L * (8 * N + 16 * E) + 8 * N


def train_step_bytes(N, E, D, L, s):
    return 3 * fwd_bytes(N, E, D, L, s)


def kernel_alg_bytes(name, N, E, M, G, D, s):
    """Algorithmic HBM bytes of ONE launch of a kernel (operands read once + results written once; weights,
    embedding tables and the indices of neighbouring rows are cache resident and not counted).  Every timing
    name is one kernel template instance (DESIGN.md section 4 lists the same figures per molecule)."""
    i4 = 4
    table = {
        "embed_rows_kernel": N * i4 + N * D * s,
        "bond_update_kernel": M * 2 * D * s + 4 * M * i4 + M * D * s + M * D + E * 16,
        "atom_gather_kernel": N * D * s + 4 * N * i4 + N * D * s,
        # tile-staged row kernels (closed row tiles; indices = ELL-4 views + row pointers)
        "tile_bond_update_kernel": M * 2 * D * s + M * D * s + M * D + E * 16 + 5 * M * i4,   # PQ in; B', tie counts, tie masks out
        "tile_bond_update_tok_kernel": M * D * s + M * D + E * 16 + 6 * M * i4,                   # tokens in (table in smem); B', tie data out
        "tile_atom_gather_kernel": N * D * s + M * D * s + N * 2 * D * s + 9 * N * i4,           # A tile + B' rows (by edge id) in; S out
        "tile_atom_bwd_kernel": 3 * N * D * s + N * D * s + 5 * N * i4,                           # dS half, dZA, A_prev in; dZA_prev out
        "tile_bond_bwd_kernel": 3 * M * D * s + M * D + E * 16 + M * 2 * D * s + 12 * M * i4,   # dS half, dBnext, B_l, counts, masks in; [dP|dQ] out
        "tile_bond_scatter_kernel": M * D * s + E * 16 + 9 * M * i4 + M * D * s,
        "fused_bond_fwd_kernel": M * D * s + M * D * s + E * 16 + 5 * M * i4,
        "fused_atom_bwd_kernel": 2 * N * D * s + 2 * N * D * s + 5 * N * i4,
        "fused_atom_fwd_kernel": 2 * N * D * s + N * 2 * D * s + N * D * s + 9 * N * i4,
        "pool_kernel": N * D * s + (G + N) * i4 + G * D * 4,
        "pool_bwd_kernel": G * D * 4 + N * D * s + N * i4 + N * D * s,
        "atom_bwd_gather_kernel": 3 * N * D * s + 4 * N * i4 + N * D * s,
        "bond_bwd_tie_kernel": 3 * M * D * s + M * D + 2 * M * i4 + 2 * M * D * s,
        "bond_bwd_scatter_kernel": M * D * s + E * 16 + 8 * M * i4 + M * D * s,
        # tcgen05 GEMMs, one entry per template instance (A read once, C written once, weights smem/L2 resident)
        "gemm_nt_tc_k128n256": M * D * s + M * 2 * D * s,                  # fwd [P|Q] = B W^T and bwd dS = dZA W (rows ~ N = M here)
        "gemm_nt_tc_k256n128_res": N * 2 * D * s + N * D * s + N * D * s,  # fwd atoms: S in, residual in, A' out
        "gemm_nt_tc_k256n128": M * 2 * D * s + M * D * s,                  # bwd dB = [dP|dQ] W
        "gemm_nt_tcs_kernel": N * 2 * D * s + 2 * N * D * s,               # wide models (streamed weights), atom shape
        "gemm_tn_tc_n128k256": N * D * s + N * 2 * D * s,                  # dW_at = dZA^T S
        "gemm_tn_tc_n256k128": M * 2 * D * s + M * D * s,                  # dW_pq = [dP|dQ]^T B
    }
    return table.get(name)


def measured_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this
    workload at this round's kernels (profiles/r2_traffic.json, made by profiles/extract_traffic.py), or None."""
    for f in ("r2_traffic.json",):
        try:
            t = json.load(open(os.path.join(ROOT, "profiles", f)))
            v = t.get(name, {}).get("dram_bytes_per_launch")
            if v is not None:
                return v
        except Exception:
            pass
    return None


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        rows = [ln for (t, ln) in self.lines if t0 <= t <= t1 + 0.2] or [ln for _, ln in self.lines]
        for ln in rows:
            f = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference CPU path (the oracle: op-for-op restatement of graphchem + PyG 2.6.1), all host threads
def oracle_throughput(config: str, batch: int, steps: int, warmup: int, seed: int = 1234):
    """(molecules/s from the MEDIAN step, seconds per step, threads) of the oracle on one `batch`-molecule batch
    of the bench's synthetic shape: a train step (fwd + MSE + bwd + Adam, predict_cn.ipynb:247-251) for
    `train`, an eval-mode forward for `infer`."""
    from graphchem_b200.synthetic import make_molecules
    from oracle.gcn_oracle import OracleMoleculeGCN, train_step
    from oracle.pyg_restated import Data as OData, collate
    cfg = CFGS[config]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    store = make_molecules(batch, seed=seed, atom_vocab=cfg["atom_vocab"], bond_vocab=cfg["bond_vocab"])
    graphs = []
    for i in range(batch):
        a0, a1 = store.atom_off[i], store.atom_off[i + 1]
        e0, e1 = store.edge_off[i], store.edge_off[i + 1]
        graphs.append(OData(x=store.atoms[a0:a1], edge_attr=store.bonds[e0:e1],
                            edge_index=store.conn[2 * e0:2 * e1].view(2, -1), y=store.targets[i:i + 1]))
    data = collate(graphs)
    torch.manual_seed(0)
    model = OracleMoleculeGCN(cfg["atom_vocab"], cfg["bond_vocab"], cfg["out_dim"], embedding_dim=cfg["D"],
                              n_messages=cfg["L"], n_readout=cfg["n_readout"], readout_dim=cfg["R"])
    if config == "train":
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        model.train()
        one = lambda: train_step(model, opt, data)  # noqa: E731
    else:
        model.eval()

        def one():
            with torch.no_grad():
                model(data)
    for _ in range(warmup):
        one()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        one()
        times.append(time.perf_counter() - t0)
    med = float(np.median(times))
    return batch / med, med, cores


def cpu_sample_text(config: str, steps: int, warmup: int, cores: int, sec: float) -> str:
    what = "train steps (fwd + MSE + bwd + Adam)" if config == "train" else "eval-mode forward passes"
    return (f"{steps} {what} after {warmup} warm-up on ONE {CPU_SAMPLE_BATCH}-molecule batch of the bench's synthetic "
            f"shape and model (fp32 torch CPU oracle, {cores} threads, median {sec:.2f} s/step; per-molecule rate)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = args.steps, args.warmup  # ~1 s per step on 16 host threads: the driver's K = 20, W = 5 take half a minute
    value, sec, cores = oracle_throughput(args.config, CPU_SAMPLE_BATCH, steps, warmup)
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    line = {
        "impl": "reference", "metric": metric_name(args.config), "value": value, "unit": "molecules/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, ranks=world),
        "cpu_baseline": {"value": value, "unit": "molecules/s", "cores": cores, "kind": "port",
                         "sample": cpu_sample_text(args.config, steps, warmup, cores, sec)},
        "e2e": {"value": value, "unit": "molecules/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = CPU oracle (op-for-op restatement of graphchem 2.3.3 on PyG 2.6.1; PyG/RDKit are not "
                "installable offline, DESIGN.md section 2); it runs on rank 0's host cores only, whatever --gpus says",
    }
    print(json.dumps(line))


def metric_name(config):
    return "MoleculeGCN train molecules/sec" if config == "train" else "MoleculeGCN inference (screening) molecules/sec"


def workload_config(args, ranks):
    cfg = CFGS[args.config]
    B = args.batch
    resident = (f"a resident shard of {args.molecules // ranks} molecules per GPU ({args.molecules // ranks // B} batches)"
                if args.molecules else f"{N_RESIDENT} distinct resident batches per GPU, cycled")
    if args.config == "train":
        work = (f"synthetic ZINC-scale molecules ({ATOMS} atoms / {EDGES} directed edges), batch {B} per GPU, MoleculeGCN "
                f"D={cfg['D']} n_messages={cfg['L']} n_readout={cfg['n_readout']} R={cfg['R']}, train step = fwd + MSE + bwd + "
                "Adam (BASELINE.json configs[3]; configs[2] shape)")
    else:
        work = (f"inference-only screening of synthetic ZINC-scale molecules ({ATOMS} atoms / {EDGES} directed edges), batch {B} "
                f"per GPU, MoleculeGCN D={cfg['D']} n_messages={cfg['L']} n_readout={cfg['n_readout']} R={cfg['R']}, forward only, "
                "replicas (BASELINE.json configs[4])")
    return {
        "workload": work, "global_batch": B * ranks, "atoms_per_molecule": ATOMS, "edges_per_molecule": EDGES,
        "parallelism": (f"dp{ranks}" if args.config == "train" else f"replicas{ranks}"), "resident": resident,
        "l2": "no flush needed: the per-step working set (> 1 GB of activations) exceeds the 126 MB L2 and distinct "
              "resident batches are cycled",
    }


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from graphchem_b200 import _lib
    from graphchem_b200.nn import MoleculeGCN
    from graphchem_b200.synthetic import make_molecules
    from graphchem_b200.train import InferStep, Screener, TrainStep

    cfg = CFGS[args.config]
    train = args.config == "train"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL's own log (NCCL_DEBUG=INFO: rank / channel / NVLS lines) stays on; it goes to stderr so that stdout
        # carries the one JSON line
        if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
            os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()

    B, K, W = args.batch, args.steps, args.warmup
    N, E = ATOMS * B, EDGES * B
    D, L = cfg["D"], cfg["L"]
    s_bytes = 2
    torch.manual_seed(0)
    model = MoleculeGCN(cfg["atom_vocab"], cfg["bond_vocab"], cfg["out_dim"], embedding_dim=D, n_messages=L,
                        n_readout=cfg["n_readout"], readout_dim=cfg["R"], compute_dtype=torch.bfloat16).to(dev)
    if world > 1 and train:  # identical replicas
        dist.broadcast(model.flat_parameters(), src=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def note(msg):
        print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- host side: synthetic shard (not timed as GPU work) -----------------------------------------
    n_batches = (args.molecules // world) // B if args.molecules else N_RESIDENT
    if n_batches < 1:
        raise SystemExit("--molecules is smaller than one batch per GPU")
    t_gen0 = time.perf_counter()
    store = make_molecules(n_batches * B, seed=1234 + rank, atom_vocab=cfg["atom_vocab"], bond_vocab=cfg["bond_vocab"])
    note(f"generated {n_batches * B} molecules in {time.perf_counter() - t_gen0:.1f} s")
    t_pack0 = time.perf_counter()
    host_batches = [store.collate_pack(np.arange(i * B, (i + 1) * B), L, cfg["atom_vocab"], cfg["bond_vocab"], pin=True)
                    for i in range(min(n_batches, N_RESIDENT))]
    pack_ms = (time.perf_counter() - t_pack0) * 1e3 / len(host_batches)
    if n_batches <= N_RESIDENT:
        dev_batches = [hb.to(dev) for hb in host_batches]
    else:
        # the resident shard: raw collate on the host, one 0.7 kB/molecule copy, packed on the device (bit-identical
        # to the host packer, tests/test_pack_device.py)
        dev_batches, scratch = [], None
        for i in range(n_batches):
            raw = store.collate_raw(np.arange(i * B, (i + 1) * B), L, cfg["atom_vocab"], cfg["bond_vocab"], pin=False).to(dev)
            if scratch is None:
                scratch = torch.empty(int(lib.gcb_pack_device_scratch_ints(raw.header.ctypes.data)), dtype=torch.int32, device=dev)
            dev_batches.append(raw.pack(scratch=scratch))
        torch.cuda.synchronize()
        note(f"resident shard: {n_batches} batches, {sum(b.nbytes for b in dev_batches) / 2**30:.1f} GiB packed in HBM, "
             f"{time.perf_counter() - t_pack0:.1f} s")
    torch.cuda.synchronize()

    step = (TrainStep(model, lr=1e-3, capture_after=1, max_graphs=n_batches + 8) if train
            else InferStep(model))  # resident batches repeat: capture on first sight
    if not train:
        step.max_graphs = n_batches + 8
        model.eval()

    # ---- (1) device-resident throughput -------------------------------------------------------
    note("graph capture + warm-up")
    for i in range(n_batches):  # one CUDA graph per resident batch (set-up, not a timed or warm-up step)
        step(dev_batches[i])
    for i in range(W):
        step(dev_batches[i % n_batches])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record()
    for i in range(K):
        step(dev_batches[(W + i) % n_batches])
    ev1.record()
    barrier()
    t_wall1 = time.time()
    ms = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    loss_resident = float(step.loss.item()) if train else None
    value = B * world * K / (ms * 1e-3)
    note(f"resident: {value:.0f} mol/s")

    # ---- (2) end to end through the public API with HOST buffers ---------------------------------
    if train:
        # per step: pinned packed batch -> HBM (copy stream, double buffered), step, loss -> pinned host
        copy_stream = torch.cuda.Stream(dev)
        n_ints = host_batches[0].ints.numel()
        stage = [(torch.empty(n_ints, dtype=torch.int32, device=dev), torch.empty(B, 1, dtype=torch.float32, device=dev))
                 for _ in range(2)]
        from graphchem_b200.data.packed import PackedBatch
        staged = [PackedBatch(si, host_batches[0].header, sy) for si, sy in stage]
        loss_host = torch.zeros(K + W + 4, dtype=torch.float32).pin_memory()
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        n_host = len(host_batches)

        def upload(i):
            hb, slot = host_batches[i % n_host], i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[slot])
                hb.upload_into(stage[slot][0])  # compact prefix over PCIe; gcb_pack_expand_ell rebuilds the ELL-4 views
                stage[slot][1].copy_(hb.y, non_blocking=True)
                ready[slot].record(copy_stream)

        def e2e_loop(n, base):
            cur = torch.cuda.current_stream(dev)
            upload(base)
            for i in range(n):
                slot = (base + i) % 2
                if i + 1 < n:
                    upload(base + i + 1)
                cur.wait_event(ready[slot])
                loss = step(staged[slot])
                consumed[slot].record(cur)
                loss_host[base + i:base + i + 1].copy_(loss, non_blocking=True)

        for ev in consumed:
            ev.record(torch.cuda.current_stream(dev))
        e2e_loop(max(W, 2) + (max(W, 2) % 2), 0)  # warm-up (captures the two staging graphs); even count keeps slots aligned
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_loop(K, 0)
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))
        e2e_steps = K
        h2d = host_batches[0].compact_ints * 4 + B * 4
        d2h = 4
        e2e_api = ("pinned pre-packed PackedBatch -> HBM (PackedBatch.upload_into on a copy stream, double buffered: compact prefix over "
                   "PCIe, ELL-4 views rebuilt by gcb_pack_expand_ell) -> TrainStep -> loss to pinned host, every step")
    else:
        # host molecules -> screen(): raw collate on host threads -> H2D -> gcb_pack_device -> forward -> predictions
        # gathered on the device and copied back once (B * out floats per step)
        n_stream = max(8, min(K, 24))
        sub = make_molecules(n_stream * B, seed=4321 + rank, atom_vocab=cfg["atom_vocab"], bond_vocab=cfg["bond_vocab"])
        # half of the host threads: with one worker per core the main thread (H2D, graph launches) gets descheduled for
        # milliseconds at a time and the call time doubles at random (profiles/micro/screener_diag.py: 4-8 workers are
        # steady at 8.1 ms per 32768-molecule batch, 16 workers swing between 8.2 and 13-18 ms on a 16-thread box)
        workers = max(2, min(8, (os.cpu_count() or 2) // (2 * max(1, world))))
        screener = Screener(model, batch_size=B, workers=workers)
        screener(sub)  # warm-up: pinned pool, device slot ring, one CUDA graph per slot (kept by the Screener)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        preds = screener(sub)
        e1.record()
        barrier()
        assert preds.shape[0] == n_stream * B and bool(torch.isfinite(preds).all())
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))
        e2e_steps = n_stream
        h2d = int(lib.gcb_raw_ints(N, E, B)) * 4 + B * 4
        d2h = B * 4
        e2e_api = (f"graphchem_b200.train.Screener over a host MoleculeStore: raw collate on {workers} host threads -> H2D -> "
                   "gcb_pack_device -> forward -> predictions D2H; CUDA events around the whole call")
    e2e_value = B * world * e2e_steps / (ms_e2e * 1e-3)
    note(f"e2e: {e2e_value:.0f} mol/s")

    # ---- (3) per-kernel timing with CUDA events (un-graphed pass, rank 0, N == 1 only) -------------
    roofline, kernels, cpu_baseline, stream, stream_dev = None, None, None, None, None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_hbm = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    step_bytes = train_step_bytes(N, E, D, L, s_bytes) if train else fwd_bytes(N, E, D, L, s_bytes)
    step_roofline = {"alg_bytes": step_bytes, "achieved": step_bytes * world / (ms / K * 1e-3) / 1e9,
                     "peak": peak_hbm * world, "frac": step_bytes / (ms / K * 1e-3) / 1e9 / peak_hbm,
                     "roofline_molecules_per_s": peak_hbm * 1e9 * world / (step_bytes / B)}
    # every rank: one un-graphed step to count kernel launches (collective-safe: all ranks take part)
    if train:
        plain = TrainStep(model, lr=1e-3, use_cuda_graph=False)
        plain.exp_avg, plain.exp_avg_sq, plain.step_count = step.exp_avg, step.exp_avg_sq, step.step_count
    else:
        plain = InferStep(model, use_cuda_graph=False)
    plain(dev_batches[0])
    torch.cuda.synchronize()
    l0 = lib.gcb_launch_count()
    plain(dev_batches[1 % n_batches])
    launches_per_step = int(lib.gcb_launch_count() - l0)
    barrier()
    note(f"launches/step {launches_per_step}")
    if world > 1:
        roofline = {"bound": "hbm", "kernel": "whole step (the per-kernel breakdown is reported at N=1)",
                    "achieved": step_roofline["achieved"], "peak": step_roofline["peak"], "unit": "GB/s",
                    "frac": step_roofline["frac"], "traffic": None, "peak_source": peak_src, "step": step_roofline}
    if rank == 0 and world == 1:
        for i in range(2):
            plain(dev_batches[i % n_batches])
        torch.cuda.synchronize()
        lib.gcb_timing_enable(1)
        n_prof = 8
        for i in range(n_prof):
            plain(dev_batches[i % n_batches])
        torch.cuda.synchronize()
        buf = C.create_string_buffer(1 << 16)
        lib.gcb_timing_report(buf, len(buf))
        lib.gcb_timing_enable(0)
        rep = json.loads(buf.value.decode())
        total_ms = sum(r["ms"] for r in rep) or 1.0
        kernels = []
        M = min(N, E)
        for r in sorted(rep, key=lambda r: -r["ms"]):
            per = r["ms"] / r["launches"]
            alg = kernel_alg_bytes(r["name"], N, E, M, B, D, s_bytes)
            kernels.append({"name": r["name"], "launches_per_step": r["launches"] / n_prof, "ms_per_launch": per,
                            "share": r["ms"] / total_ms,
                            "alg_GBps": (alg / (per * 1e-3) / 1e9) if alg else None})
        # the dominant kernel = largest share among the kernels of the message-passing path (those with an algorithmic
        # byte count); small reductions / the readout MLP have none, and side-stream kernels are timed while they wait
        top = next((k for k in kernels if k["alg_GBps"] is not None and not k["name"].startswith("gemm_tn")), kernels[0])
        alg_top = kernel_alg_bytes(top["name"], N, E, M, B, D, s_bytes)
        roofline = {
            "bound": "hbm", "kernel": top["name"], "achieved": top["alg_GBps"], "peak": peak_hbm, "unit": "GB/s",
            "frac": (top["alg_GBps"] / peak_hbm) if top["alg_GBps"] else None, "traffic": measured_traffic(top["name"]),
            "alg_bytes_per_launch": alg_top, "ms_per_launch": top["ms_per_launch"], "share_of_step": top["share"],
            "launches_per_step": top["launches_per_step"], "peak_source": peak_src,
            "timing": "CUDA events around every launch of an un-graphed step on the launching stream (gcb_timing_enable); "
                      "kernels of the side stream overlap the main stream, so per-kernel times sum to more than the step",
            "step": step_roofline,
        }
        if train and not args.no_extra:
            # ---- (3b) streaming: host collate + pack on a thread pool INSIDE the timed region ------
            def stream_leg(device_pack):
                from graphchem_b200.data import PackedLoader
                # half of the host threads (see the screening leg: one worker per core starves the launching thread)
                workers = max(2, min(8, (os.cpu_count() or 2) // 2))
                loader = PackedLoader(store, B, L, cfg["atom_vocab"], cfg["bond_vocab"], device=dev, workers=workers,
                                      prefetch=4, device_slots=3, device_pack=device_pack)
                n_stream, done = max(8, min(K, 24)), 0
                for pb in loader:  # one warm-up epoch: pinned pool, slot ring, three graph captures
                    step(pb)
                torch.cuda.synchronize()
                t_s0 = time.perf_counter()
                while done < n_stream:
                    for pb in loader:
                        step(pb)
                        done += 1
                        if done >= n_stream:
                            break
                float(step.loss.item())
                t_s = time.perf_counter() - t_s0
                return {"value": B * n_stream / t_s, "unit": "molecules/s", "steps": n_stream, "pack_threads": workers,
                        "note": ("raw collate on host threads, H2D of the raw batch, pack kernels on the GPU (gcb_pack_device) and "
                                 "the train step per batch, wall clock") if device_pack else
                                "collate + pack on host threads, H2D and the train step per batch, wall clock"}

            try:
                stream = stream_leg(False)
            except Exception as e:  # informative only
                stream = {"error": repr(e)}
            try:
                stream_dev = stream_leg(True)
            except Exception as e:  # informative only
                stream_dev = {"error": repr(e)}
        # ---- (4) reference CPU path on this box's host cores: the SAME bounded sample as `--impl reference` --------
        if not args.no_cpu_baseline:
            cb, cb_sec, cores = oracle_throughput(args.config, CPU_SAMPLE_BATCH, 5, 3)
            cpu_baseline = {"value": cb, "unit": "molecules/s", "cores": cores, "kind": "port",
                            "sample": cpu_sample_text(args.config, 5, 3, cores, cb_sec)}

    if rank == 0:
        line = {
            "metric": metric_name(args.config), "value": value, "unit": "molecules/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": "molecules/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps, "api": e2e_api},
            "gpu_launches": (launches_per_step or 0) * K,
            "launches_per_step": launches_per_step,
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline, "kernels": kernels,
            "loss": loss_resident, "host_pack_ms_per_batch": pack_ms, "resident_batches": n_batches,
            "e2e_stream": stream, "e2e_stream_device_pack": stream_dev,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # tear down in order: CUDA graphs that captured NCCL work first, then the communicator; a
        # watchdog guarantees the process exits even if NCCL teardown stalls
        sys.stdout.flush()
        threading.Timer(30.0, lambda: os._exit(0)).start()
        step._graphs.clear()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--config", default="train", choices=["train", "infer"])
    ap.add_argument("--batch", type=int, default=None, help="molecules per GPU per step (train 8192, infer 32768)")
    ap.add_argument("--molecules", type=int, default=0,
                    help="train: total molecules of the resident shard over all GPUs (BASELINE configs[3]: 10000000)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the informative streaming legs")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = CFGS[args.config]["batch"]
    if args.steps is None:
        args.steps = 200 if args.molecules else 40
    args.warmup = max(3, args.warmup)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
