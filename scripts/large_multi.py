"""Config 5 over several GPUs: yalps_multi_solve_large (K4m) against the single-GPU grid kernel (K4) on the same
tableau, bit for bit, with device-timed microseconds per pivot.  Usage: python scripts/large_multi.py [m nv cap] ..."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import yalps_b200  # noqa: E402
from yalps_b200 import engine as E  # noqa: E402

shapes = [(4096, 8192, 48), (1024, 2048, 200)]
if len(sys.argv) >= 4:
    shapes = [tuple(int(x) for x in sys.argv[1:4])]
ndev = torch.cuda.device_count()
worlds = [w for w in (2, 4, 8) if w <= ndev] or [2]
eng = yalps_b200.Engine(0)
for m, nv, cap in shapes:
    H, W = m + 1, nv + 1
    d = torch.empty(H * W, dtype=torch.float64, device="cuda:0")
    eng.generate_synthetic_device(0, 1, m, nv, d.data_ptr())
    torch.cuda.synchronize()
    mats = d.cpu().numpy().reshape(1, -1)
    opt = E.make_options(max_pivots=cap)
    work = torch.empty_like(d)
    one = eng.solve_batch(mats, H, W, opt, want_matrices=True)
    # device time of K4 alone: the device entry on a resident copy
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = None
    for _ in range(3):
        work.copy_(d)
        torch.cuda.synchronize()
        ev0.record()
        eng.solve_batch_device(1, H, W, work.data_ptr(), opt, d_work=work.data_ptr())
        ev1.record()
        torch.cuda.synchronize()
        t = ev0.elapsed_time(ev1)
        best = t if best is None else min(best, t)
    piv = int(one["pivots"][0].sum())
    print(json.dumps({"shape": [H, W], "pivots": piv, "ranks": 1, "kernel": "K4", "ms": round(best, 3),
                      "us_per_pivot": round(1e3 * best / max(piv, 1), 2)}), flush=True)
    for world in worlds:
        devs = [i % ndev for i in range(world)]
        with yalps_b200.MultiEngine(devs) as me:
            ms = None
            for _ in range(3):
                got = me.solve_large(mats, H, W, opt, want_matrix=True)
                ms = got["kernel_ms"] if ms is None else min(ms, got["kernel_ms"])
            same = (np.array_equal(got["matrix"].view(np.uint64), one["matrices"][0].view(np.uint64))
                    and got["pivots"] == tuple(int(x) for x in one["pivots"][0]) and got["status"] == int(one["status"][0]))
            print(json.dumps({"shape": [H, W], "pivots": piv, "ranks": world, "gpus": len(set(devs)), "kernel": "K4m",
                              "ms": round(ms, 3), "us_per_pivot": round(1e3 * ms / max(piv, 1), 2),
                              "speedup_vs_K4": round(best / ms, 2), "bit_identical": bool(same)}), flush=True)
