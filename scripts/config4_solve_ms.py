"""solve() wall time of the config-4 models (10 calls after 3 warm-ups): python scripts/config4_solve_ms.py
YALPS_DENSE=1 forces the dense-image entry (yalps_solve) for an A/B against the sparse entry (yalps_solve_sparse)."""
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import yalps_b200, bench_workloads as BW
from yalps_b200 import solver
if os.environ.get("YALPS_DENSE"):
    solver.SPARSE_OVER_BYTES = 1 << 62
eng = yalps_b200.Engine(0)
for name in ("Large Farm MIP", "Monster 2", "Vendor Selection", "Monster Problem"):
    c = BW.milp_case(name)
    info = {}
    for _ in range(3):
        yalps_b200.solve(c["model"], c["options"], engine=eng, info=info)
    t0 = time.perf_counter()
    for _ in range(10):
        sol = yalps_b200.solve(c["model"], c["options"], engine=eng, info=info)
    ms = (time.perf_counter() - t0) / 10 * 1e3
    t0 = time.perf_counter()
    for _ in range(10):
        yalps_b200.tableau_model(c["model"], solver.SPARSE_OVER_BYTES)
    build = (time.perf_counter() - t0) / 10 * 1e3
    print(json.dumps({"model": name, "dense_entry": bool(os.environ.get("YALPS_DENSE")), "solve_ms": round(ms, 3),
                      "host_tableau_build_ms": round(build, 3), "bnb_ms": info["bnb_us"] / 1e3, "result": sol["result"],
                      "nodes": info["nodes"], "root_pivots": info["root_pivots"]}), flush=True)
eng.close()
