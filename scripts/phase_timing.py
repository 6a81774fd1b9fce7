#!/usr/bin/env python
"""Debug: per-phase cycle counters of the row-split kernel (library built with -DYALPS_TIMING, see
simplex_split.cuh YT_MARK).  python scripts/phase_timing.py NAME threads rows"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from yalps_b200 import _ffi
_ffi.LIB_PATH = os.path.join(ROOT, "yalps_b200", "libyalps_timing.so")
import yalps_b200
from conftest import load_netlib
NL = load_netlib()
eng = yalps_b200.Engine(0)
stream = torch.cuda.current_stream().cuda_stream
labels = ["sel1", "sel2", "normalise", "column+compact", "barrier1", "rhs", "update", "barrier2"]
for spec in sys.argv[1:]:
    name, threads, rows = spec.split(":")
    if name.startswith("DENSE"):  # DENSEmxn: synthetic dense LP (config 5 family), 200 pivots
        m_, nv_ = (int(x) for x in name[5:].split("x"))
        H, W = m_ + 1, nv_ + 1
        d = torch.empty(H * W, dtype=torch.float64, device="cuda")
        eng.generate_synthetic_device(0, 1, m_, nv_, d.data_ptr())
        work = torch.empty_like(d)
    elif name.startswith("MILP="):  # MILP=<case name>: the model's initial tableau on the HBM/L2-resident split kernel
        import bench_workloads as BW
        tb = yalps_b200.tableau_model(BW.milp_case(name[5:].replace("_", " "))["model"]).tableau
        H, W = tb.height, tb.width
        d = torch.from_numpy(tb.matrix.copy()).cuda()
        work = torch.empty_like(d)
    else:
        g = NL.get(name); H, W = g["height"], g["width"]
        d = torch.from_numpy(np.asarray(g["matrix"], np.float64).reshape(-1).copy()).cuda()
        work = None
    rhs = torch.zeros(H, dtype=torch.float64, device="cuda")
    if int(threads) <= 0:  # 0: KC (one cluster), -1: KG (the whole grid, shared-memory resident)
        eng.set_tuning(5 if int(threads) == 0 else 7, 0, 0)
        labels = ["sel1", "sel2(+exchange)", "dsmem+normalise", "column+B1", "rhs+obj", "update", "cluster wait (KG: publish+record)",
                  "writeback+B2 (KG: + row staging)"]
    else:
        eng.set_tuning(2 if name.startswith("MILP=") else 1, int(threads), int(rows))
    for _ in range(2):
        if work is not None:
            work.copy_(d)
            eng.solve_batch_device(1, H, W, work.data_ptr(), yalps_b200.engine.make_options(max_pivots=200),
                                   d_work=work.data_ptr(), d_rhs=rhs.data_ptr(), stream=stream)
        else:
            eng.solve_batch_device(1, H, W, d.data_ptr(), d_rhs=rhs.data_ptr(), stream=stream)
        torch.cuda.synchronize()
    r = rhs.cpu().numpy()
    for who, off in ((("thread 0", 0), ("last thread", 10)) if int(threads) <= 0 else (("thread 0", 0), ("thread 32", 9), ("last thread", 18))):
        piv = r[off + 8]
        print(f"{name} {H}x{W} threads={threads} rows={rows} {who}: pivots={int(piv)} cycles/pivot: " +
              ", ".join(f"{l}={r[off + k] / max(piv, 1):.0f}" for k, l in enumerate(labels)) +
              (f", record wait={r[off + 9] / max(piv, 1):.0f}" if int(threads) < 0 else "") +
              f"  total={(sum(r[off:off + 8]) + (r[off + 9] if int(threads) < 0 else 0)) / max(piv, 1):.0f}")
eng.close()
