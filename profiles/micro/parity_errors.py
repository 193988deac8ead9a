"""Error table of the CUDA path against the CPU oracle (one B200): bf16 forward error per depth and width at the
benchmarked shapes, gradient errors, fp32 errors against the float64 oracle.  One JSON line per case; the
committed copy is profiles/r2_parity_errors.jsonl and the tolerances of tests/test_parity_bench_shape_gpu.py /
tests/test_parity_gpu.py are set from it.
    python profiles/micro/parity_errors.py > gpurun_out/r2_parity_errors.jsonl"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from tests.test_parity_bench_shape_gpu import forward_errors, gradient_errors, synthetic_case

torch.set_num_threads(os.cpu_count() or 1)
for n_mol, D, L, grads in [(4096, 128, 1, False), (4096, 128, 2, False), (4096, 128, 3, False), (4096, 128, 4, True),
                           (8192, 128, 4, True), (4096, 128, 6, False), (4096, 256, 4, False), (4096, 256, 6, False),
                           (2048, 256, 6, True)]:
    t0 = time.time()
    data, pb, oracle, model = synthetic_case(n_mol, D, L)
    rec = {"molecules": n_mol, "D": D, "L": L, "dtype": "bf16", "n_tiles": int(pb.n_tiles),
           "fwd_err_out_atom_bond": [float(f"{e:.3e}") for e in forward_errors(data, pb, oracle, model)]}
    if grads:
        ge, le = gradient_errors(data, pb, oracle, model)
        rec["grad_err_max"] = float(f"{max(ge.values()):.3e}")
        rec["grad_err"] = {k: float(f"{v:.2e}") for k, v in ge.items()}
        rec["loss_rel_err"] = float(f"{le:.2e}")
    rec["sec"] = round(time.time() - t0, 1)
    print(json.dumps(rec), flush=True)
