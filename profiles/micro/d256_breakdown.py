import sys, json, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from graphchem_b200 import _lib
from graphchem_b200.nn import MoleculeGCN
from graphchem_b200.synthetic import make_molecules
dev=torch.device('cuda',0)
model = MoleculeGCN(64, 32, 1, embedding_dim=256, n_messages=6, n_readout=2, readout_dim=64, compute_dtype=torch.bfloat16).to(dev).eval()
store = make_molecules(2*8192, seed=7)
bs=[store.collate_pack(np.arange(i*8192,(i+1)*8192),6,64,32).to(dev) for i in range(2)]
lib=_lib.lib()
with torch.no_grad():
    for i in range(3): model(bs[i%2])
    torch.cuda.synchronize()
    lib.gcb_timing_enable(1)
    for i in range(4): model(bs[i%2])
    torch.cuda.synchronize()
buf=C.create_string_buffer(1<<16); lib.gcb_timing_report(buf,len(buf)); lib.gcb_timing_enable(0)
rep=json.loads(buf.value.decode())
tot=sum(r['ms'] for r in rep)
for r in sorted(rep,key=lambda r:-r['ms']): print(f"{r['name']:30s} {r['launches']/4:5.1f} {r['ms']/r['launches']*1e3:8.1f} us  {r['ms']/tot*100:5.1f}%")
print('total per fwd', tot/4, 'ms')
