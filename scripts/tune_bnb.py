"""Branch-and-cut wave size / CTA shape sweep (config 4 models): python scripts/tune_bnb.py"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import yalps_b200
from conftest import load_cases
eng = yalps_b200.Engine(0)
for name in ("Large Farm MIP", "Knapsack 1", "Fancy Stock Cutting Problem", "Monster 2"):
    c = next(x for x in load_cases() if x["name"] == name)
    for threads, rows in ((0, 0), (128, 4), (256, 4), (256, 8), (512, 8)):
        for wave in (16, 32, 64, 96, 128, 148, 256):
            eng.set_tuning(0, threads, rows); eng.set_wave(wave)
            info = {}
            yalps_b200.solve(c["model"], c["options"], engine=eng, info=info)
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter(); yalps_b200.solve(c["model"], c["options"], engine=eng, info=info); best = min(best, time.perf_counter() - t0)
            print(json.dumps({"model": name, "threads": threads, "rows": rows, "wave": wave, "ms": round(best * 1e3, 2), "waves": info["waves"],
                              "device_nodes": info["device_nodes"], "wave_us": info["wave_us"], "bnb_us": info["bnb_us"],
                              "us_per_wave": round(info["wave_us"] / max(info["waves"], 1), 1)}), flush=True)
