import sys, torch, torch.nn.functional as F
sys.path.insert(0, '/root/repo')
from tests.util import make_pair, random_batch
from tests.test_parity_gpu import _to_dev
for act in (F.tanh, F.softplus):
    data, _ = random_batch(seed=11, n_graphs=8)
    oracle, model = make_pair(12, 7, 2, embedding_dim=64, n_messages=3, n_readout=2, readout_dim=48, act_fn=act)
    model.eval()
    d = _to_dev(data)
    with torch.no_grad():
        ref = [t.clone() for t in model(d)]
        bad = 0
        for it in range(300):
            out = model(d)
            for k, (a, b) in enumerate(zip(out, ref)):
                if not torch.equal(a, b):
                    bad += 1
                    diff = (a.float() - b.float()).abs()
                    idx = (diff > 0).nonzero()
                    print(act.__name__, "iter", it, "output", k, "maxdiff", float(diff.max()), "n_diff", idx.shape[0], "rows", sorted(set(idx[:, 0].tolist()))[:10])
                    break
            if bad > 5: break
    print(act.__name__, "mismatching iterations:", bad)
