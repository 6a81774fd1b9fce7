#!/usr/bin/env python
"""Secondary workloads of BASELINE.json (configs 1, 3, 4, 5) on one B200; prints one JSON line per workload.
The headline number (config 2) is bench.py; these lines go to profiles/ as supporting measurements.
    python scripts/bench_extra.py [config1] [config3] [config4] [config5] [netlib]
CPU columns are the oracle port (tests/golden timings or timed here), never part of the product path."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

import yalps_b200
from yalps_b200 import engine as E
from conftest import load_cases, load_netlib

HBM_PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
eng = yalps_b200.Engine(0)
NL = load_netlib()
what = sys.argv[1:] or ["config1", "config3", "config4", "config5", "netlib"]


def emit(**kw):
    print(json.dumps(kw), flush=True)


def ev_time(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


if "config1" in what:  # AFIRO single-LP latency through the host API
    g = NL.get("AFIRO")
    for _ in range(5):
        r = eng.solve_batch(g["matrix"], g["height"], g["width"])
    t0 = time.perf_counter()
    for _ in range(200):
        r = eng.solve_batch(g["matrix"], g["height"], g["width"])
    dt = (time.perf_counter() - t0) / 200
    emit(workload="config1: Netlib AFIRO 36x33, one LP through yalps_solve_batch (host call, H2D+kernel+D2H)",
         status=int(r["status"][0]), value=float(r["value"][0]), pivots=[int(x) for x in r["pivots"][0]],
         latency_us=dt * 1e6, oracle_cpu_us=g["oracle_seconds"] * 1e6 or None)

if "config3" in what:  # RHS-perturbed replicas, SMEM-resident (K1) vs HBM-resident (K2)
    for name, n in (("SC105", 32768), ("ADLITTLE", 65536)):
        g = NL.get(name)
        H, W = g["height"], g["width"]
        d = torch.empty(n * H * W, dtype=torch.float64, device="cuda")
        eng.generate_replicas_device(0, n, g["matrix"], H, W, g["row_groups"], d.data_ptr())
        work = torch.empty_like(d)
        st = torch.empty(n, dtype=torch.int32, device="cuda")
        piv = torch.empty(n, 2, dtype=torch.int64, device="cuda")
        val = torch.empty(n, dtype=torch.float64, device="cuda")
        stream = torch.cuda.current_stream().cuda_stream
        for path, pname in ((E.PATH_AUTO, "auto policy"), (E.PATH_SMEM, "K1 smem-resident"), (E.PATH_GMEM, "K2 hbm-resident")):
            for threads in ((0,) if path == E.PATH_AUTO else (64, 128, 256) if path == E.PATH_SMEM else (64, 128, 256)):
                eng.set_tuning(path, threads)

                def run():
                    if path != E.PATH_SMEM:
                        work.copy_(d)
                    eng.solve_batch_device(n, H, W, d.data_ptr(), d_work=work.data_ptr(), d_status=st.data_ptr(),
                                           d_value=val.data_ptr(), d_pivots=piv.data_ptr(), stream=stream)
                ms = ev_time(run, reps=2)
                if path != E.PATH_SMEM:
                    ms -= ev_time(lambda: work.copy_(d), reps=2)
                p = int(piv.sum().item())
                emit(workload=f"config3: {n} RHS-perturbed (eps=1e-2) replicas of Netlib {name} {H}x{W}", kernel=pname,
                     threads_per_lp=threads, ms=ms, lps_per_s=n / ms * 1e3, pivots_per_s=p / ms * 1e3,
                     pivots_per_lp=p / n, optimal=int((st == 0).sum().item()),
                     dense_bytes_per_pivot=16 * H * W,
                     hbm_roofline_dense_frac=(p * 16 * H * W / (ms * 1e-3) / 1e9) / HBM_PEAK,
                     value_range=[float(val.min().item()), float(val.max().item())], base_value=g["value"])
        eng.set_tuning(0, 0)
        del d, work

if "config5" in what:  # one large dense LP across the grid (K4)
    for (m, nv, cap) in ((1024, 2048, 1e18), (4096, 8192, 300)):
        H, W = m + 1, nv + 1
        d = torch.empty(H * W, dtype=torch.float64, device="cuda")
        work = torch.empty_like(d)
        piv = torch.empty(1, 2, dtype=torch.int64, device="cuda")
        st = torch.empty(1, dtype=torch.int32, device="cuda")
        eng.generate_synthetic_device(0, 1, m, nv, d.data_ptr())
        opt = E.make_options(max_pivots=cap if cap < 1e17 else float("inf"))
        stream = torch.cuda.current_stream().cuda_stream

        def run():
            work.copy_(d)
            eng.solve_batch_device(1, H, W, d.data_ptr(), opt, d_work=work.data_ptr(), d_status=st.data_ptr(),
                                   d_pivots=piv.data_ptr(), stream=stream)
        ms = ev_time(run, reps=1, warm=1) - ev_time(lambda: work.copy_(d), reps=2)
        p = int(piv.sum().item())
        emit(workload=f"config5: one synthetic dense LP {m}x{nv} (tableau {H}x{W}, {H * W * 8 / 1e6:.1f} MB), grid-wide kernel K4",
             status=int(st.item()), pivots=p, ms=ms, pivots_per_s=p / ms * 1e3, us_per_pivot=ms * 1e3 / max(p, 1),
             dense_bytes_per_pivot=16 * H * W, achieved_gbs=p * 16.0 * H * W / (ms * 1e-3) / 1e9,
             hbm_roofline_frac=(p * 16.0 * H * W / (ms * 1e-3) / 1e9) / HBM_PEAK,
             note="max_pivots capped" if cap < 1e17 else "full solve")
        del d, work

if "netlib" in what:  # every Netlib parity model, one LP at a time through the host API
    rows = []
    for name in NL.names:
        g = NL.get(name)
        opt = E.make_options(check_cycles=g["check_cycles"])
        eng.solve_batch(g["matrix"], g["height"], g["width"], opt)
        t0 = time.perf_counter()
        r = eng.solve_batch(g["matrix"], g["height"], g["width"], opt)
        dt = time.perf_counter() - t0
        ok = int(r["status"][0]) == g["status"] and tuple(int(x) for x in r["pivots"][0]) == g["pivots"]
        rows.append({"name": name, "shape": [g["height"], g["width"]], "pivots": sum(g["pivots"]), "gpu_ms": dt * 1e3,
                     "oracle_cpu_ms": g["oracle_seconds"] * 1e3, "parity": ok})
    tot_g, tot_c = sum(r["gpu_ms"] for r in rows), sum(r["oracle_cpu_ms"] for r in rows)
    emit(workload="netlib: 51 parity models, one LP per call (host API, incl. H2D/D2H)", models=rows,
         total_gpu_ms=tot_g, total_oracle_cpu_ms=tot_c, all_parity=all(r["parity"] for r in rows))

if "config4" in what:  # MILP suite of benchmarks/json/read.ts
    from oracle import model as M
    for name in ("Large Farm MIP", "Monster 2", "Monster Problem", "Vendor Selection", "Knapsack 1"):
        c = next(x for x in load_cases() if x["name"] == name)
        info = {}
        yalps_b200.solve(c["model"], c["options"], engine=eng, info=info)
        t0 = time.perf_counter()
        sol = yalps_b200.solve(c["model"], c["options"], engine=eng, info=info)
        dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        yalps_b200.tableau_model(c["model"])
        build = time.perf_counter() - t0
        t0 = time.perf_counter()
        M.solve(c["model"], {**M.DEFAULT_OPTIONS, **c["options"]})
        cpu = time.perf_counter() - t0
        emit(workload=f"config4: {name} via solve() (tableau build on host + root LP + branch and cut waves)",
             status=sol["status"], result=sol["result"], expected=c["expected"]["result"], gpu_ms=dt * 1e3,
             host_tableau_build_ms=build * 1e3,
             oracle_cpu_ms=cpu * 1e3, nodes=info["nodes"], node_pivots=info["node_pivots"], waves=info["waves"],
             device_nodes=info["device_nodes"], wave_us=info.get("wave_us"), bnb_us=info.get("bnb_us"), root_pivots=list(info["root_pivots"]),
             shape=[info["height"], info["width"]])
eng.close()
