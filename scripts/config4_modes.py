"""Config 4 (MILP suite) through solve(): device-resident search (mode 2/0) against the host wave driver (mode 1).
    python scripts/config4_modes.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yalps_b200
import bench_workloads as BW
eng = yalps_b200.Engine(0)
for name in ("Large Farm MIP", "Knapsack 1", "Fancy Stock Cutting Problem", "Integer Wood Shop Problem", "Monster 2", "Vendor Selection"):
    c = BW.milp_case(name)
    row = {"model": name}
    for mode, label in ((1, "host_waves"), (0, "auto")):
        eng.set_bnb_mode(mode)
        info = {}
        for _ in range(3):
            yalps_b200.solve(c["model"], c["options"], engine=eng, info=info)
        reps = 10
        t0 = time.perf_counter()
        for _ in range(reps):
            sol = yalps_b200.solve(c["model"], c["options"], engine=eng, info=info)
        dt = (time.perf_counter() - t0) / reps
        tm0 = time.perf_counter()
        yalps_b200.tableau_model(c["model"])
        build = time.perf_counter() - tm0
        row[label] = {"ms": round(dt * 1e3, 3), "bnb_ms": info["bnb_us"] / 1e3, "waves": info["waves"], "nodes": info["nodes"],
                      "device_nodes": info["device_nodes"], "result": sol["result"], "host_tableau_build_ms": round(build * 1e3, 3)}
    print(json.dumps(row), flush=True)
eng.close()
