"""Is the CPU oracle run-to-run deterministic on this host?  (a retry in the parity tests was blamed on
torch's CPU kernels.)  Evaluates the oracle forward + backward N times on the same inputs, per thread count, and
reports every run whose outputs / gradients are not bit-identical to the first, with the first layer at which the
activations diverge.  CPU only; prints JSON lines."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F

from oracle.gcn_oracle import OracleMoleculeGCN
from tests.util import random_batch

N = int(sys.argv[1]) if len(sys.argv) > 1 else 300
print(json.dumps({"cpu_count": os.cpu_count(), "threads": torch.get_num_threads(), "torch": torch.__version__,
                  "mkldnn": torch.backends.mkldnn.is_available()}), flush=True)
for threads in (os.cpu_count() or 1, 1):
    torch.set_num_threads(threads)
    for (ng, D, L, act) in [(8, 64, 3, F.tanh), (16, 128, 2, F.softplus), (64, 128, 4, F.softplus)]:
        data, _ = random_batch(seed=11, n_graphs=ng)
        torch.manual_seed(0)
        m = OracleMoleculeGCN(12, 7, 2, embedding_dim=D, n_messages=L, n_readout=2, readout_dim=48, act_fn=act)

        def run():
            m.zero_grad()
            out = m(data, return_layers=True)
            loss = F.mse_loss(out[0], torch.zeros_like(out[0])) + out[1].mean() + out[2].mean()
            loss.backward()
            return out, [p.grad.clone() for p in m.parameters()]

        ref, gref = run()
        bad, t0 = [], time.time()
        for i in range(N):
            out, g = run()
            same_out = all(torch.equal(a, b) for a, b in zip(out[:3], ref[:3]))
            same_g = all(torch.equal(a, b) for a, b in zip(g, gref))
            if not (same_out and same_g):
                first = None
                for l, ((a1, b1), (a0, b0)) in enumerate(zip(out[3], ref[3])):
                    if not torch.equal(b1, b0):
                        first = ("bond", l, float((b1 - b0).abs().max() / b0.abs().max())); break
                    if not torch.equal(a1, a0):
                        first = ("atom", l, float((a1 - a0).abs().max() / a0.abs().max())); break
                bad.append({"run": i, "fwd_same": same_out, "grad_same": same_g, "first_diverging_layer": first,
                            "fwd_rel": [float((a - b).abs().max() / b.abs().max()) for a, b in zip(out[:3], ref[:3])],
                            "grad_rel": max(float((a - b).abs().max() / max(float(b.abs().max()), 1e-30)) for a, b in zip(g, gref))})
        print(json.dumps({"threads": threads, "graphs": ng, "D": D, "L": L, "act": act.__name__, "runs": N,
                          "nondeterministic_runs": len(bad), "examples": bad[:5], "sec": round(time.time() - t0, 1)}), flush=True)
