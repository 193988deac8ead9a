import sys, time, os
sys.path.insert(0,'/root/repo')
import numpy as np, torch
from graphchem_b200.synthetic import make_molecules
from graphchem_b200.data import PackedLoader
print("cpus", os.cpu_count(), open('/proc/cpuinfo').read().count('processor\t:'))
store = make_molecules(8192*8, seed=1)
for w in (1,2,4,8,16):
    ld = PackedLoader(store, 8192, 4, 64, 32, device="cuda", workers=w, prefetch=4)
    for _ in ld: pass
    torch.cuda.synchronize()
    t0=time.perf_counter(); n=0
    for _ in range(3):
        for b in ld: n+=1
    torch.cuda.synchronize()
    dt=time.perf_counter()-t0
    print(w,'workers', round(dt/n*1e3,2),'ms/batch', round(8192*n/dt/1e6,2),'M mol/s', flush=True)
