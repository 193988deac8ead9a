"""Where does the tcgen05 NT GEMM's time go?  One B200: the [188416 x K] x [N x K]^T GEMM of the bench through the
debug entry with the epilogue pieces switched on one by one (bias, degree-scaled bias, residual, activation),
several distinct operand sets cycled so that inputs come from HBM; plus the memory ceilings of this GPU for the
read : write mixes those kernels have (torch fill / copy / add on 1 GiB buffers).  One JSON line per case."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from graphchem_b200 import _lib

lib = _lib.lib()
rows = 8192 * 23
dev = "cuda"


def timed(fn, n=20, warm=5):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us


# ---- memory ceilings ----
n = 1 << 29  # bf16 elements = 1 GiB
a = torch.empty(n, dtype=torch.bfloat16, device=dev).normal_()
b = torch.empty_like(a).normal_()
c = torch.empty_like(a)
for name, fn, nbytes in [("fill (0R:1W)", lambda i: c.fill_(1.0), 2 * n), ("copy (1R:1W)", lambda i: c.copy_(a), 4 * n),
                         ("add (2R:1W)", lambda i: torch.add(a, b, out=c), 6 * n), ("sum (1R:0W)", lambda i: a.sum(), 2 * n)]:
    us = timed(fn, 10, 3)
    print(json.dumps({"case": name, "us": round(us, 1), "GBps": round(nbytes / us / 1e3, 1)}), flush=True)
del a, b, c

# ---- GEMM epilogue pieces ----
NSET = 4
g = torch.Generator().manual_seed(0)
for K, N in [(128, 256), (256, 128)]:
    A = [(torch.randn(rows, K, generator=g) * 0.5).bfloat16().to(dev) for _ in range(NSET)]
    W = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16().to(dev)
    Cm = [torch.empty(rows, N, dtype=torch.bfloat16, device=dev) for _ in range(NSET)]
    bias = torch.randn(N, generator=g).to(dev)
    d = torch.randint(0, 4, (rows,), generator=g)
    rp = torch.cat([torch.zeros(1, dtype=torch.long), d.cumsum(0)]).int().to(dev)
    R = [torch.randn(rows, N, generator=g).bfloat16().to(dev) for _ in range(NSET)]
    s = torch.cuda.current_stream().cuda_stream
    for label, use_b, use_rp, use_r, act in [("plain", 0, 0, 0, -1), ("bias", 1, 0, 0, -1), ("deg*bias", 1, 1, 0, -1),
                                             ("act", 0, 0, 0, 0), ("res", 0, 0, 1, -1), ("res+act", 0, 0, 1, 0),
                                             ("deg*bias+res+act", 1, 1, 1, 0)]:
        def run(i):
            k = i % NSET
            rc = lib.gcb_debug_gemm_nt_bf16(A[k].data_ptr(), W.data_ptr(), Cm[k].data_ptr(), rows, N, K,
                                            bias.data_ptr() if use_b else None, rp.data_ptr() if use_rp else None, 0,
                                            R[k].data_ptr() if use_r else None, act, s)
            assert rc == 0
        us = timed(run)
        nbytes = rows * (K + N + (N if use_r else 0)) * 2
        print(json.dumps({"case": f"gemm K={K} N={N} {label}", "us": round(us, 1), "alg_GBps": round(nbytes / us / 1e3, 1)}), flush=True)
    del A, Cm, R
