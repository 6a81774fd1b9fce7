import os, sys, time, json
ROOT="/root/repo"; sys.path.insert(0, ROOT); sys.path.insert(0, ROOT+"/tests")
import yalps_b200
from conftest import load_cases
eng = yalps_b200.Engine(0)
for name in ("Large Farm MIP", "Knapsack 1", "Fancy Stock Cutting Problem", "Monster 2"):
    c = next(x for x in load_cases() if x["name"] == name)
    for threads in (0, 32, 64, 128, 256, 512):
        for wave in (16, 64, 256):
            eng.set_tuning(0, threads); eng.set_wave(wave)
            info = {}
            yalps_b200.solve(c["model"], c["options"], engine=eng, info=info)
            t0 = time.perf_counter(); yalps_b200.solve(c["model"], c["options"], engine=eng, info=info); dt = time.perf_counter() - t0
            print(name[:12], "threads", threads, "wave", wave, "ms", round(dt*1e3, 2), "waves", info["waves"], "dev_nodes", info["device_nodes"], "wave_us", info["wave_us"], "us/wave", round(info["wave_us"]/max(info["waves"],1),1))
