"""solve() statistics of the big MILP models (waves, device wave time, root pivots): python scripts/milp_info.py"""
import sys, os, time, json
sys.path.insert(0, os.getcwd())
import yalps_b200, bench_workloads as BW
eng = yalps_b200.Engine(0)
for name in ("Monster 2", "Vendor Selection"):
    c = BW.milp_case(name)
    info = {}
    for _ in range(3): yalps_b200.solve(c["model"], c["options"], engine=eng, info=info)
    t0=time.perf_counter()
    for _ in range(5): yalps_b200.solve(c["model"], c["options"], engine=eng, info=info)
    print(name, round((time.perf_counter()-t0)/5*1e3,3), {k: info[k] for k in info if k not in ("final_pos","final_rhs","final_var")})
eng.close()
