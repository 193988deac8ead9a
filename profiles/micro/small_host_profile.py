import cProfile, pstats, sys, os, io, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from graphchem_b200.data import MoleculeStore, PackedLoader
from graphchem_b200.nn import MoleculeGCN
from graphchem_b200.train import TrainStep
from tests.util import encoded_fuel_set
enc, train, test = encoded_fuel_set("cn")
av, bv = enc.vocab_sizes
KW = dict(embedding_dim=128, n_messages=3, n_readout=3, readout_dim=128)
model = MoleculeGCN(av, bv, 1, **KW).cuda().train()
store = MoleculeStore.from_graphs(train)
loader = PackedLoader(store, 16, 3, av, bv, device="cuda", shuffle=True, seed=0, workers=2, prefetch=4)
step = TrainStep(model, lr=1e-3, use_cuda_graph=False)
for _ in range(3):
    for pb in loader: step(pb)
torch.cuda.synchronize()
def run(n):
    for _ in range(n):
        losses = [step(pb).clone() for pb in loader]
    torch.cuda.synchronize()
t0=time.perf_counter(); run(20); t=time.perf_counter()-t0
print("ms/step", 1e3*t/(20*len(loader)))
# loader alone
t0=time.perf_counter()
for _ in range(20):
    for pb in loader: pass
torch.cuda.synchronize(); print("loader alone ms/batch", 1e3*(time.perf_counter()-t0)/(20*len(loader)))
pbs=[pb for pb in PackedLoader(store, 16, 3, av, bv, device="cuda", shuffle=True, seed=0, workers=2, prefetch=4, cache=True)]
t0=time.perf_counter()
for _ in range(20):
    for pb in pbs: step(pb)
torch.cuda.synchronize(); print("step alone (resident batches, eager) ms/step", 1e3*(time.perf_counter()-t0)/(20*len(pbs)))
pr=cProfile.Profile(); pr.enable(); run(10); pr.disable()
s=io.StringIO(); pstats.Stats(pr,stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:5000])
