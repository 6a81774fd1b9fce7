#!/usr/bin/env python
"""K1 (shared memory) against K1t (tensor memory) on dense synthetic batches, device-resident inputs, CUDA events."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import yalps_b200
from yalps_b200 import engine as E
eng = yalps_b200.Engine(0)
stream = torch.cuda.current_stream().cuda_stream
shapes = [(32, 64, 65536), (32, 64, 262144), (24, 48, 65536), (16, 32, 65536), (32, 32, 65536), (8, 64, 65536), (8, 16, 65536), (16, 64, 65536),
          (32, 16, 65536), (24, 64, 65536), (32, 48, 65536)]
if len(sys.argv) > 1: shapes = shapes[int(sys.argv[2]) if len(sys.argv) > 2 else 0: int(sys.argv[1])]
for (m, nv, n) in shapes:
    H, W = m + 1, nv + 1
    d = torch.empty(n * H * W, dtype=torch.float64, device="cuda")
    eng.generate_synthetic_device(0, n, m, nv, d.data_ptr())
    st = torch.empty(n, dtype=torch.int32, device="cuda"); piv = torch.empty(n, 2, dtype=torch.int64, device="cuda")
    val = torch.empty(n, dtype=torch.float64, device="cuda")
    ref = None
    for path, name in ((E.PATH_SMEM, "K1"), (E.PATH_TMEM, "K1t"), (E.PATH_AUTO, "auto")):
        eng.set_tuning(path, 0)
        run = lambda: eng.solve_batch_device(n, H, W, d.data_ptr(), d_status=st.data_ptr(), d_pivots=piv.data_ptr(), d_value=val.data_ptr(), stream=stream)
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        p = int(piv.sum().item())
        sig = (p, int(st.sum().item()), float(val.nan_to_num().sum().item()))
        if ref is None: ref = sig
        print(json.dumps({"shape": [H, W], "n": n, "kernel": name, "ms": round(ms, 4), "Mpivots_per_s": round(p / ms / 1e3, 1),
                          "MLPs_per_s": round(n / ms / 1e3, 2), "agrees_with_K1": sig == ref,
                          "smem_roofline_frac_36.6TBs": round(p * 16.0 * H * W / (ms * 1e-3) / 36.6e12, 3)}), flush=True)
eng.close()
