"""One launch of the config-3 path (automatic policy: the row-split HBM/L2-resident kernel K2s) over n RHS-perturbed
replicas of a Netlib model, for profiling:   python scripts/config3_case.py [SC105|ADLITTLE] [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import yalps_b200
from yalps_b200 import engine as E
import bench_workloads as BW
name = sys.argv[1] if len(sys.argv) > 1 else "SC105"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
g = BW.netlib_base(name)
H, W = g["height"], g["width"]
eng = yalps_b200.Engine(0)
d = torch.empty(n * H * W, dtype=torch.float64, device="cuda")
eng.generate_replicas_device(0, n, g["matrix"], H, W, g["row_groups"], d.data_ptr())
work = torch.empty_like(d)
piv = torch.empty(n, 2, dtype=torch.int64, device="cuda")
st = torch.empty(n, dtype=torch.int32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
rows = torch.zeros(1, dtype=torch.int64, device="cuda")
for it in range(3):
    if it == 2:
        eng.set_row_counter(rows.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.solve_batch_device(n, H, W, d.data_ptr(), E.make_options(), d_work=work.data_ptr(), d_status=st.data_ptr(),
                           d_pivots=piv.data_ptr(), stream=stream)
    e1.record()
    torch.cuda.synchronize()
    p = int(piv.sum().item())
    print(f"{name} {H}x{W} x {n}: {p} pivots, {e0.elapsed_time(e1):.3f} ms, {p / e0.elapsed_time(e1) / 1e3:.2f} M pivots/s, "
          f"optimal {int((st == 0).sum().item())}, rows {int(rows.item())}")
print("algorithmic bytes per launch:", BW.pivot_bytes(H, W, p, int(rows.item())))
eng.close()
