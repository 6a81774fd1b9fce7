"""Root LP of a benchmarks/json MILP model through KG (YALPS_CASE_PATH=7) or the automatic path, device-resident, CUDA events:
    python scripts/kg_case_milp.py "Monster 2" "Vendor Selection\""""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import yalps_b200
from yalps_b200 import engine as E
import bench_workloads as BW
eng = yalps_b200.Engine(0)
stream = torch.cuda.current_stream().cuda_stream
for name in sys.argv[1:] or ["Monster 2", "Vendor Selection"]:
    c = BW.milp_case(name)
    tm = yalps_b200.tableau_model(c["model"])
    t = tm.tableau
    H, W = t.height, t.width
    d = torch.from_numpy(np.asarray(t.matrix, np.float64).copy()).cuda()
    work = torch.empty_like(d)
    piv = torch.empty(1, 2, dtype=torch.int64, device="cuda")
    eng.set_tuning(int(os.environ.get("YALPS_CASE_PATH", E.PATH_AUTO)), 0)
    best = 1e9
    for _ in range(4):
        work.copy_(d)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.solve_batch_device(1, H, W, work.data_ptr(), E.make_options(), d_work=work.data_ptr(), d_pivots=piv.data_ptr(), stream=stream)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    p = int(piv.sum().item())
    print(f"{os.environ.get('YALPS_B200_LIB', 'lib')[-14:]} {name} {H}x{W}: {p} pivots, {best:.3f} ms, {best * 1e3 / max(p, 1):.2f} us/pivot")
eng.close()
