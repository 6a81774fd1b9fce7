"""One dense LP through the grid kernel (K4) or, with YALPS_CASE_PATH=7, the grid-resident kernel (KG), in place on the
device, for profiling and A/B runs:
    python scripts/k4_case.py [m] [n] [max_pivots] [netlib name]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import yalps_b200
from yalps_b200 import engine as E
import bench_workloads as BW
m = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
nv = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
cap = float(sys.argv[3]) if len(sys.argv) > 3 else 200.0
eng = yalps_b200.Engine(0)
if len(sys.argv) > 4:
    g = BW.netlib_base(sys.argv[4])
    H, W = g["height"], g["width"]
    d = torch.from_numpy(g["matrix"]).cuda()
else:
    H, W = m + 1, nv + 1
    d = torch.empty(H * W, dtype=torch.float64, device="cuda")
    eng.generate_synthetic_device(0, 1, m, nv, d.data_ptr())
work = torch.empty_like(d)
piv = torch.empty(1, 2, dtype=torch.int64, device="cuda")
opt = E.make_options(max_pivots=cap)
stream = torch.cuda.current_stream().cuda_stream
path = int(os.environ.get('YALPS_CASE_PATH', E.PATH_GRID))
eng.set_tuning(path, 0)
for _ in range(3):
    work.copy_(d)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.solve_batch_device(1, H, W, work.data_ptr(), opt, d_work=work.data_ptr(), d_pivots=piv.data_ptr(), stream=stream)
    e1.record()
    torch.cuda.synchronize()
    p = int(piv.sum().item())
    print(f"{os.environ.get('YALPS_B200_LIB', 'lib')[-12:]} path {path} {H}x{W}: {p} pivots, {e0.elapsed_time(e1) * 1e3 / max(p, 1):.2f} us/pivot")
eng.close()
