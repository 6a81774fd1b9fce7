#!/usr/bin/env python
"""One Netlib LP (or n copies) through one forced kernel configuration, device-resident: the ncu target for
latency-mode profiling.   python scripts/one_lp.py KLEIN1 1 256 8 [n] [reps]   (path 1=smem 2=gmem)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import yalps_b200
from conftest import load_netlib
name, path, threads, rows = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
n = int(sys.argv[5]) if len(sys.argv) > 5 else 1
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
eng = yalps_b200.Engine(0)
g = load_netlib().get(name); H, W = g["height"], g["width"]
d = torch.from_numpy(np.tile(np.asarray(g["matrix"], np.float64).reshape(-1), n)).cuda()
work = torch.empty_like(d)
st = torch.empty(n, dtype=torch.int32, device="cuda"); piv = torch.empty(n, 2, dtype=torch.int64, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
eng.set_tuning(path, threads, rows)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(reps):
    if path != 1: work.copy_(d)
    e0.record()
    eng.solve_batch_device(n, H, W, d.data_ptr(), d_work=work.data_ptr(), d_status=st.data_ptr(), d_pivots=piv.data_ptr(), stream=stream)
    e1.record(); torch.cuda.synchronize()
    p = int(piv[0].sum().item())
    print(f"{name} {H}x{W} n={n} path={path} threads={threads} rows={rows}: {e0.elapsed_time(e1)*1e3:.1f} us, {p} pivots, {e0.elapsed_time(e1)*1e3/max(p,1):.3f} us/pivot")
eng.close()
