import sys, torch, torch.nn.functional as F
sys.path.insert(0, '/root/repo')
from tests.util import make_pair, random_batch, nerr
from tests.test_parity_gpu import _to_dev
data, _ = random_batch(seed=11, n_graphs=8)
oracle, model = make_pair(12, 7, 2, embedding_dim=64, n_messages=3, n_readout=2, readout_dim=48, act_fn=F.tanh)
model.eval(); oracle.eval()
with torch.no_grad():
    out = model(_to_dev(data)); ref = oracle(data)
torch.cuda.synchronize()
print([nerr(a.float(), b) for a, b in zip(out, ref)])
model.train(); oracle.train()
o = model(_to_dev(data)); (o[0].sum() + o[1].float().sum() * 0.01 + o[2].float().sum() * 0.01).backward()
torch.cuda.synchronize()
print("bwd ok")
