import json, os, sys, time
sys.path.insert(0, os.getcwd())
import yalps_b200
import bench_workloads as BW
eng = yalps_b200.Engine(0)
for name in ("Large Farm MIP", "Knapsack 1", "Fancy Stock Cutting Problem", "Integer Wood Shop Problem"):
    c = BW.milp_case(name)
    eng.set_bnb_mode(0)
    info = {}
    for _ in range(3):
        yalps_b200.solve(c["model"], c["options"], engine=eng, info=info)
    reps = 10
    t0 = time.perf_counter()
    for _ in range(reps):
        sol = yalps_b200.solve(c["model"], c["options"], engine=eng, info=info)
    dt = (time.perf_counter() - t0) / reps
    print(os.environ.get("YALPS_BNB_SPEC"), name, round(dt*1e3,3), "bnb_ms", info["bnb_us"]/1e3, "waves", info["waves"], "nodes", info["nodes"], "device_nodes", info["device_nodes"], sol["result"], flush=True)
eng.close()
