#!/usr/bin/env python
"""Per-pivot latency of ONE LP (and of one LP per SM) for every kernel path and CTA width, device-resident
inputs, CUDA events: the data behind the latency-mode policy (single solve() calls, branch-and-cut nodes)."""
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
names = sys.argv[1:] or ["AFIRO", "KLEIN1", "ADLITTLE", "BLEND", "SC105", "SC205"]
# warm the clocks
x = torch.empty(1 << 28, dtype=torch.float32, device="cuda")
for _ in range(20): x.add_(1.0)
torch.cuda.synchronize()
for name in names:
    g = NL.get(name); H, W = g["height"], g["width"]
    for n in ((1,) if os.environ.get('ONLY_ONE') else (1, 148)):
        d = torch.from_numpy(np.tile(np.asarray(g["matrix"], np.float64).reshape(-1), n)).cuda()
        work = torch.empty_like(d)
        st = torch.empty(n, dtype=torch.int32, device="cuda"); piv = torch.empty(n, 2, dtype=torch.int64, device="cuda")
        for path, pn in ((E.PATH_SMEM, "K1"), (E.PATH_GMEM, "K2"), (E.PATH_GRID, "K4"), (E.PATH_CLUSTER, "KC"), (E.PATH_AUTO, "auto")):
            for threads, rows in (((0, 0),) if path in (E.PATH_GRID, E.PATH_CLUSTER, E.PATH_AUTO) else
                                  ((32, 1), (64, 1), (128, 1), (256, 1), (128, 2), (128, 4), (256, 2), (256, 4), (256, 8),
                                   (512, 4), (512, 8), (512, 16))):
                if path == E.PATH_GRID and n > 1: continue
                eng.set_tuning(path, threads, rows)
                def run():
                    if path != E.PATH_SMEM: work.copy_(d)
                    eng.solve_batch_device(n, H, W, d.data_ptr(), d_work=work.data_ptr(), d_status=st.data_ptr(),
                                           d_pivots=piv.data_ptr(), stream=stream)
                try:
                    run(); torch.cuda.synchronize()
                except Exception as e:
                    continue
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 5
                e0.record()
                for _ in range(reps): run()
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                p = int(piv[0].sum().item())
                print(json.dumps({"model": name, "shape": [H, W], "n": n, "kernel": pn, "threads": threads, "rows": rows,
                                  "ms": round(ms, 4), "pivots": p, "us_per_pivot": round(1e3 * ms / max(p, 1), 3)}), flush=True)
eng.close()
