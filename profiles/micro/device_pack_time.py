"""Device-side packer timing (one B200): CUDA-event time of gcb_pack_device per 8192-molecule batch of the bench
shape, per kernel (gcb_timing facility), H2D bytes, and the host halves (raw collate vs full host pack) on one core.
Prints one JSON line."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from graphchem_b200 import _lib
from graphchem_b200.synthetic import make_molecules

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
store = make_molecules(4 * B, seed=1234)
lib = _lib.lib()
idx = [np.arange(i * B, (i + 1) * B) for i in range(4)]
ref = store.collate_pack(idx[0], 4, 64, 32)
raws = [store.collate_raw(i, 4, 64, 32) for i in idx]
out_h = torch.empty(ref.ints.numel() + 64, dtype=torch.int32).pin_memory()
y_h = torch.empty(B, dtype=torch.float32).pin_memory()


def host_ms(fn):
    fn(idx[0])
    t0 = time.perf_counter()
    for i in idx:
        fn(i)
    return (time.perf_counter() - t0) / len(idx) * 1e3


ms_pack = host_ms(lambda i: store.collate_pack(i, 4, 64, 32, out=out_h, y_out=y_h))
ms_raw = host_ms(lambda i: store.collate_raw(i, 4, 64, 32, out=out_h, y_out=y_h))
dev = [r.to("cuda") for r in raws]
total = int(raws[0].header[_lib.H_TOTAL_INTS])
out = torch.empty(total, dtype=torch.int32, device="cuda")
scr = torch.empty(int(lib.gcb_pack_device_scratch_ints(raws[0].header.ctypes.data)), dtype=torch.int32, device="cuda")
pb = dev[0].pack(out=out, scratch=scr)
torch.cuda.synchronize()
assert torch.equal(pb.ints.cpu(), ref.ints), "device pack differs from host pack"
for k in range(5):
    dev[k % 4].pack(out=out, scratch=scr)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 40
e0.record()
for k in range(n):
    dev[k % 4].pack(out=out, scratch=scr)
e1.record()
torch.cuda.synchronize()
ms_dev = e0.elapsed_time(e1) / n
lib.gcb_timing_enable(1)
for k in range(8):
    dev[k % 4].pack(out=out, scratch=scr)
torch.cuda.synchronize()
buf = C.create_string_buffer(1 << 14)
lib.gcb_timing_report(buf, len(buf))
lib.gcb_timing_enable(0)
kern = {r["name"]: round(r["ms"] / r["launches"] * 1e3, 2) for r in json.loads(buf.value.decode())}
print(json.dumps({"batch": B, "device_pack_ms": ms_dev, "device_pack_M_mol_per_s": B / ms_dev / 1e3,
                  "packed_bytes": total * 4, "raw_bytes": raws[0].ints.numel() * 4,
                  "host_full_pack_ms_1core": ms_pack, "host_raw_collate_ms_1core": ms_raw, "kernel_us_per_launch": kern}))
