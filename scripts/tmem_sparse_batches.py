#!/usr/bin/env python
"""Big batches of SPARSE small LPs (density probe < 0.35 sends them to K2 today): K2 against K1 and K1t."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import yalps_b200
from yalps_b200 import engine as E
eng = yalps_b200.Engine(0)
stream = torch.cuda.current_stream().cuda_stream
torch.manual_seed(1)
for (m, nv, n, zero) in ((32, 64, 65536, 0.6), (32, 64, 65536, 0.8), (16, 32, 65536, 0.7), (24, 48, 65536, 0.9)):
    H, W = m + 1, nv + 1
    d = torch.empty(n * H * W, dtype=torch.float64, device="cuda")
    eng.generate_synthetic_device(0, n, m, nv, d.data_ptr())
    t = d.view(n, H, W)
    mask = torch.rand(n, H, W, device="cuda") < zero
    mask[:, :, 0] = False
    t[mask] = 0.0
    work = torch.empty_like(d)
    st = torch.empty(n, dtype=torch.int32, device="cuda"); piv = torch.empty(n, 2, dtype=torch.int64, device="cuda")
    row = {"shape": [H, W], "n": n, "zeros": zero}
    ref = None
    for path, name in ((E.PATH_AUTO, "auto"), (E.PATH_GMEM, "K2"), (E.PATH_SMEM, "K1"), (E.PATH_TMEM, "K1t")):
        eng.set_tuning(path, 0)
        def run():
            work.copy_(d)
            eng.solve_batch_device(n, H, W, work.data_ptr(), d_work=work.data_ptr(), d_status=st.data_ptr(), d_pivots=piv.data_ptr(), stream=stream)
        for _ in range(2): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        e0.record()
        for _ in range(5): work.copy_(d)
        e1.record(); torch.cuda.synchronize()
        ms -= e0.elapsed_time(e1) / 5
        sig = (int(piv.sum().item()), int(st.sum().item()))
        ref = ref or sig
        row[name + "_ms"] = round(ms, 3)
        row[name + "_ok"] = sig == ref
    row["Mpivots"] = round(ref[0] / 1e6, 2)
    print(json.dumps(row), flush=True)
eng.close()
