"""One big sparse LP (the root tableau of a benchmarks/json MILP) on K4 (whole grid) against the HBM/L2-resident
row-split kernel K2s (one CTA): per-pivot latency, device-resident, CUDA events.
    python scripts/big_sparse_paths.py ["Monster 2" ...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import yalps_b200
from yalps_b200 import engine as E
import bench_workloads as BW
eng = yalps_b200.Engine(0)
stream = torch.cuda.current_stream().cuda_stream
for name in sys.argv[1:] or ["Monster 2", "Vendor Selection", "Monster Problem"]:
    tm = yalps_b200.tableau_model(BW.milp_case(name)["model"])
    t = tm.tableau
    H, W = t.height, t.width
    for n in (6,):
        d = torch.from_numpy(np.tile(t.matrix, n)).cuda()
        work = torch.empty_like(d)
        st = torch.empty(n, dtype=torch.int32, device="cuda"); piv = torch.empty(n, 2, dtype=torch.int64, device="cuda")
        pos = torch.empty(n, W + H, dtype=torch.int32, device="cuda"); var = torch.empty(n, W + H, dtype=torch.int32, device="cuda")
        configs = [(E.PATH_GRID, 0, 0, "K4")] + [(E.PATH_GMEM, t, r, f"K2s {t}x{r}") for t in (128, 256, 512, 1024) for r in (1, 2, 4, 8) if t // r >= 32]
        for path, threads, rows, label in configs:
            eng.set_tuning(path, threads, rows)
            def run():
                work.copy_(d)
                eng.solve_batch_device(n, H, W, work.data_ptr(), d_work=work.data_ptr(), d_status=st.data_ptr(), d_pivots=piv.data_ptr(),
                                       d_pos=pos.data_ptr(), d_var=var.data_ptr(), stream=stream)
            try:
                run(); torch.cuda.synchronize()
            except Exception as e:
                print(json.dumps({"model": name, "kernel": label, "error": str(e)[:80]})); continue
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record(); run(); e1.record(); torch.cuda.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(); work.copy_(d); c1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) - c0.elapsed_time(c1)
            p = int(piv[0].sum().item())
            print(json.dumps({"model": name, "shape": [H, W], "n": n, "kernel": label, "ms": round(ms, 3), "pivots": p,
                              "us_per_pivot": round(1e3 * ms / max(p, 1), 2), "status": int(st[0].item())}), flush=True)
        eng.set_tuning(0, 0, 0)
eng.close()
