#!/usr/bin/env python
"""Latency mode: few dense LPs (n <= 2 x SMs).  K1t (one warp per LP, no CTA barrier) against the automatic choice
(row-split K1s) -- where should the automatic policy switch to K1t?"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import yalps_b200
from yalps_b200 import engine as E
eng = yalps_b200.Engine(0)
stream = torch.cuda.current_stream().cuda_stream
for (m, nv) in ((32, 64), (16, 32), (8, 16)):
    H, W = m + 1, nv + 1
    for n in (1, 8, 32, 74, 148, 296, 592, 1184):
        d = torch.empty(n * H * W, dtype=torch.float64, device="cuda")
        eng.generate_synthetic_device(0, n, m, nv, d.data_ptr())
        st = torch.empty(n, dtype=torch.int32, device="cuda"); piv = torch.empty(n, 2, dtype=torch.int64, device="cuda")
        row = {"shape": [H, W], "n": n}
        for path, name in ((E.PATH_AUTO, "auto"), (E.PATH_TMEM, "K1t"), (E.PATH_SMEM, "K1")):
            eng.set_tuning(path, 0)
            run = lambda: eng.solve_batch_device(n, H, W, d.data_ptr(), d_status=st.data_ptr(), d_pivots=piv.data_ptr(), stream=stream)
            for _ in range(3): run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): run()
            e1.record(); torch.cuda.synchronize()
            row[name + "_us"] = round(e0.elapsed_time(e1) / 20 * 1e3, 1)
        row["pivots_max"] = int(piv.sum(1).max().item())
        print(json.dumps(row), flush=True)
# sparse tableaus (60 % zeros outside the RHS column): the row-split kernels compact the active rows, K1t only skips
# the arithmetic of inactive rows
import numpy as np
sys.path.insert(0, os.path.join(ROOT))
from oracle import lib as O
rng = np.random.default_rng(3)
for (m, nv, zeros) in ((32, 64, 0.6), (16, 32, 0.6), (32, 64, 0.9), (32, 32, 0.9), (16, 32, 0.9)):
    H, W = m + 1, nv + 1
    for n in (1, 32, 148, 296):
        t = O.generate_synthetic(9, n, m, nv, 4).reshape(n, H, W).copy()
        mask = rng.random(t.shape) < zeros
        mask[:, :, 0] = False
        t[mask] = 0.0
        d = torch.from_numpy(t.reshape(-1)).cuda()
        st = torch.empty(n, dtype=torch.int32, device="cuda"); piv = torch.empty(n, 2, dtype=torch.int64, device="cuda")
        row = {"shape": [H, W], "n": n, "sparse": zeros}
        cells = H * W
        cfg = (128, 4) if cells < 2500 else (256, 8) if cells < 4000 else (256, 4)
        for args, name in (((E.PATH_SMEM,) + cfg, "K1s"), ((E.PATH_TMEM, 0, 0), "K1t")):
            eng.set_tuning(*args)
            run = lambda: eng.solve_batch_device(n, H, W, d.data_ptr(), d_status=st.data_ptr(), d_pivots=piv.data_ptr(), stream=stream)
            for _ in range(3): run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): run()
            e1.record(); torch.cuda.synchronize()
            row[name + "_us"] = round(e0.elapsed_time(e1) / 20 * 1e3, 1)
        row["pivots_max"] = int(piv.sum(1).max().item())
        print(json.dumps(row), flush=True)
eng.close()
