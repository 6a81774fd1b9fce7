#!/usr/bin/env python
"""Sparse Netlib replicas, batch sizes between latency mode and the big batches the sparse -> K2 rule was tuned on:
what the automatic policy picks against the forced paths."""
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
for name in ("AFIRO", "KLEIN1", "ADLITTLE", "SC105"):
    g = NL.get(name); H, W = g["height"], g["width"]
    mat = np.asarray(g["matrix"], np.float64).reshape(-1)
    for n in (296, 592, 1184, 4096, 16384):
        d = torch.from_numpy(np.tile(mat, n)).cuda()
        work = torch.empty_like(d)
        st = torch.empty(n, dtype=torch.int32, device="cuda"); piv = torch.empty(n, 2, dtype=torch.int64, device="cuda")
        row = {"model": name, "shape": [H, W], "n": n}
        for label, args in (("auto", (E.PATH_AUTO, 0, 0)), ("K1", (E.PATH_SMEM, 0, 0)), ("K1_128x2", (E.PATH_SMEM, 128, 2)),
                            ("K2", (E.PATH_GMEM, 0, 0)), ("K1t", (E.PATH_TMEM, 0, 0))):
            eng.set_tuning(*args)
            def run():
                work.copy_(d)
                eng.solve_batch_device(n, H, W, work.data_ptr(), d_work=work.data_ptr(), d_status=st.data_ptr(), d_pivots=piv.data_ptr(), stream=stream)
            try:
                for _ in range(2): run()
                torch.cuda.synchronize()
            except Exception:
                continue
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): run()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            e0.record()
            for _ in range(5): work.copy_(d)
            e1.record(); torch.cuda.synchronize()
            row[label + "_us"] = round((ms - e0.elapsed_time(e1) / 5) * 1e3, 1)
        print(json.dumps(row), flush=True)
eng.close()
