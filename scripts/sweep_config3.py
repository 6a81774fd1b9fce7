#!/usr/bin/env python
"""Config 3 (RHS-perturbed Netlib replicas): throughput of every kernel path / CTA shape, device-resident."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import yalps_b200
from yalps_b200 import engine as E
from conftest import load_netlib
eng = yalps_b200.Engine(0)
stream = torch.cuda.current_stream().cuda_stream
NL = load_netlib()
combos = [(0, 0), (128, 1), (256, 1), (64, 2), (128, 2), (128, 4), (256, 2), (256, 4), (256, 8), (512, 8), (512, 16)]
if os.environ.get('SWEEP_THIN'):
    combos = [(0, 0), (32, 1), (64, 1), (64, 2)]
if os.environ.get('SWEEP_K1_ONLY'):
    combos = [(0, 0), (64, 2), (128, 2), (128, 4), (256, 4), (256, 8), (512, 8)]
if os.environ.get('SWEEP_K2_ONLY'):
    combos = [c for c in combos if c[1] != 1 or c == (0, 0)]
for name, n in (("SC105", 16384), ("ADLITTLE", 32768), ("AFIRO", 65536), ("BLEND", 16384)):
    if len(sys.argv) > 1 and name not in sys.argv[1:]: continue
    g = NL.get(name); H, W = g["height"], g["width"]
    d = torch.empty(n * H * W, dtype=torch.float64, device="cuda")
    eng.generate_replicas_device(0, n, g["matrix"], H, W, g["row_groups"], d.data_ptr())
    work = torch.empty_like(d)
    st = torch.empty(n, dtype=torch.int32, device="cuda"); piv = torch.empty(n, 2, dtype=torch.int64, device="cuda")
    for path, pn in (((E.PATH_AUTO, "auto"), (E.PATH_SMEM, "K1")) if os.environ.get("SWEEP_K1_ONLY") else ((E.PATH_AUTO, "auto"), (E.PATH_GMEM, "K2")) if (os.environ.get("SWEEP_K2_ONLY") or os.environ.get("SWEEP_THIN")) else ((E.PATH_AUTO, "auto"), (E.PATH_SMEM, "K1"), (E.PATH_GMEM, "K2"))):
        for threads, rows in (combos[:1] if path == E.PATH_AUTO else combos[1:]):
            eng.set_tuning(path, threads, rows)
            def run():
                work.copy_(d)
                eng.solve_batch_device(n, H, W, d.data_ptr(), d_work=work.data_ptr(), d_status=st.data_ptr(), d_pivots=piv.data_ptr(), stream=stream)
            try:
                run(); torch.cuda.synchronize()
            except Exception as e:
                continue
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); run(); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 2
            e0.record(); work.copy_(d); work.copy_(d); e1.record(); torch.cuda.synchronize(); ms -= e0.elapsed_time(e1) / 2
            p = int(piv.sum().item())
            print(json.dumps({"case": name, "shape": [H, W], "n": n, "kernel": pn, "threads": threads, "rows": rows, "ms": round(ms, 3),
                              "Mpivots_per_s": round(p / ms / 1e3, 2), "kLPs_per_s": round(n / ms, 1), "optimal": int((st == 0).sum().item())}), flush=True)
eng.close()
