#!/usr/bin/env python
"""Tableaus of 34..65 rows: the 256-column K1t shape against the row-split latency kernel the policy used before,
for 1..296 LPs (Netlib AFIRO / KLEIN1 replicas and dense synthetic LPs)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import yalps_b200
from yalps_b200 import engine as E
from conftest import load_netlib
eng = yalps_b200.Engine(0)
stream = torch.cuda.current_stream().cuda_stream
NL = load_netlib()
def k1s_cfg(cells):
    return (128, 4) if cells < 2500 else (256, 8) if cells < 4000 else (256, 4) if cells < 12000 else (512, 8)
cases = []
for name in ("AFIRO", "KLEIN1"):
    g = NL.get(name)
    cases.append((name, g["height"], g["width"], np.asarray(g["matrix"], np.float64).reshape(-1), None))
for (m, nv) in ((48, 64), (64, 64), (40, 32)):
    cases.append((f"dense {m}x{nv}", m + 1, nv + 1, None, (m, nv)))
for label, H, W, mat, gen in cases:
    for n in (1, 2, 4, 8, 16, 32, 74, 148, 296, 1184, 16384):
        if mat is not None:
            d = torch.from_numpy(np.tile(mat, n)).cuda()
        else:
            d = torch.empty(n * H * W, dtype=torch.float64, device="cuda")
            eng.generate_synthetic_device(0, n, gen[0], gen[1], d.data_ptr(), neg_rows=4)
        st = torch.empty(n, dtype=torch.int32, device="cuda"); piv = torch.empty(n, 2, dtype=torch.int64, device="cuda")
        row = {"case": label, "shape": [H, W], "n": n}
        th, rw = k1s_cfg(H * W)
        for name, args in (("K1t", (E.PATH_TMEM, 0, 0)), ("K1s", (E.PATH_SMEM, th, rw)), ("K1", (E.PATH_SMEM, 0, 0))):
            if name == "K1s" and n > 296: continue
            eng.set_tuning(*args)
            run = lambda: eng.solve_batch_device(n, H, W, d.data_ptr(), d_status=st.data_ptr(), d_pivots=piv.data_ptr(), stream=stream)
            for _ in range(3): run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): run()
            e1.record(); torch.cuda.synchronize()
            row[name + "_us"] = round(e0.elapsed_time(e1) / 10 * 1e3, 1)
        row["pivots_max"] = int(piv.sum(1).max().item())
        print(json.dumps(row), flush=True)
eng.close()
