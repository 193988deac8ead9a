import sys, gc, torch, torch.nn.functional as F
sys.path.insert(0, '/root/repo')
from tests.util import make_pair, random_batch, nerr
from tests.test_parity_gpu import _to_dev
data, _ = random_batch(seed=11, n_graphs=8)
junk = []
for it in range(400):
    oracle, model = make_pair(12, 7, 2, embedding_dim=64, n_messages=3, n_readout=2, readout_dim=48, act_fn=F.tanh)
    model.eval(); oracle.eval()
    with torch.no_grad():
        out = model(_to_dev(data)); ref = oracle(data)
    errs = [nerr(a.float(), b) for a, b in zip(out, ref)]
    if max(errs) > 1e-5:
        diff = (out[1].float().cpu() - ref[1]).abs()
        bad = (diff > 1e-5).nonzero()
        print("iter", it, errs, "bad rows", sorted(set(bad[:, 0].tolist()))[:20], "bad cols", sorted(set(bad[:, 1].tolist()))[:20], "n", bad.shape[0])
    # perturb the allocator state
    junk.append(torch.empty(int(torch.randint(1, 1 << 20, (1,))), device="cuda"))
    if len(junk) > 8: junk.pop(0)
print("done")
