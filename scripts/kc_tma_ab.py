#!/usr/bin/env python
"""A/B of the winner-row staging in the cluster kernel KC (north_star (b): "pivot row staged by TMA and multicast across
a thread-block cluster"): YALPS_KC_TMA = 0 (ld.global.cg loop), 1 (cp.async.bulk per CTA + mbarrier), 2 (one
cp.async.bulk.multicast::cluster by the winner's owner).  One LP per launch, device-resident input, CUDA events, the
three modes interleaved on the same box; results must be bit-identical (status, pivots, RHS, basis)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import yalps_b200
from yalps_b200 import engine as E
from conftest import load_netlib
NL = load_netlib()
names = sys.argv[1:] or ["SC205", "SCFXM1", "DEGEN2", "E226", "BANDM"]
engines = {}
for mode in (0, 1, 2):
    os.environ["YALPS_KC_TMA"] = str(mode)
    engines[mode] = yalps_b200.Engine(0)
    engines[mode].set_tuning(E.PATH_CLUSTER, 0)
stream = torch.cuda.current_stream().cuda_stream
x = torch.empty(1 << 28, dtype=torch.float32, device="cuda")
for _ in range(20): x.add_(1.0)
torch.cuda.synchronize()
for name in names:
    g = NL.get(name); H, W = g["height"], g["width"]
    d = torch.from_numpy(np.asarray(g["matrix"], np.float64).reshape(-1)).cuda()
    work = torch.empty_like(d)
    st = torch.empty(1, dtype=torch.int32, device="cuda"); piv = torch.empty(1, 2, dtype=torch.int64, device="cuda")
    rhs = torch.empty(H, dtype=torch.float64, device="cuda"); pos = torch.empty(W + H, dtype=torch.int32, device="cuda")
    opt = E.make_options(check_cycles=g["check_cycles"])
    ref, row = None, {"model": name, "shape": [H, W]}
    times = {0: [], 1: [], 2: []}
    for rep in range(4):
        for mode in (0, 1, 2):
            eng = engines[mode]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.solve_batch_device(1, H, W, d.data_ptr(), opt, d_work=work.data_ptr(), d_status=st.data_ptr(),
                                   d_pivots=piv.data_ptr(), d_rhs=rhs.data_ptr(), d_pos=pos.data_ptr(), stream=stream)
            e1.record(); torch.cuda.synchronize()
            if rep: times[mode].append(e0.elapsed_time(e1))
            out = (int(st.item()), piv.cpu().numpy().tolist(), rhs.cpu().numpy().view(np.uint64).tolist(), pos.cpu().numpy().tolist())
            if ref is None: ref = out
            assert out == ref, (name, mode, "results differ")
    p = sum(ref[1][0])
    row["pivots"] = p
    row["matches_golden"] = bool(ref[0] == g["status"] and tuple(ref[1][0]) == g["pivots"])
    for mode, label in ((0, "ldcg"), (1, "bulk"), (2, "multicast")):
        ms = min(times[mode])
        row[label] = {"ms": round(ms, 4), "us_per_pivot": round(1e3 * ms / max(p, 1), 3)}
    print(json.dumps(row), flush=True)
for e in engines.values(): e.close()
