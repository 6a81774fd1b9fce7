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
    g = NL.get(name); H, W = g["height"], g["width"]
    d = torch.from_numpy(np.asarray(g["matrix"], np.float64).reshape(-1).copy()).cuda()
    rhs = torch.zeros(H, dtype=torch.float64, device="cuda")
    if int(threads) == 0:
        eng.set_tuning(5, 0, 0)
        labels = ["sel1", "sel2(+exchange)", "dsmem+normalise", "column+B1", "rhs+obj", "update", "cluster wait", "writeback+B2"]
    else:
        eng.set_tuning(1, int(threads), int(rows))
    for _ in range(2):
        eng.solve_batch_device(1, H, W, d.data_ptr(), d_rhs=rhs.data_ptr(), stream=stream)
        torch.cuda.synchronize()
    r = rhs.cpu().numpy()
    for who, off in ((("thread 0", 0), ("last thread", 9)) if int(threads) == 0 else (("thread 0", 0), ("thread 32", 9), ("last thread", 18))):
        piv = r[off + 8]
        print(f"{name} {H}x{W} threads={threads} rows={rows} {who}: pivots={int(piv)} cycles/pivot: " +
              ", ".join(f"{l}={r[off + k] / max(piv, 1):.0f}" for k, l in enumerate(labels)) +
              f"  total={sum(r[off:off + 8]) / max(piv, 1):.0f}")
eng.close()
