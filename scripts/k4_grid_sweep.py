"""K4 per-pivot time against the number of CTAs (YALPS_GRID_CTAS): fewer CTAs = cheaper grid barriers, less update bandwidth.
    python scripts/k4_grid_sweep.py"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cases = [("0", "0", "1e9", "25FV47"), ("1024", "2048", "400"), ("300", "900", "400"), ("2048", "4096", "100")]
for ctas in (148, 111, 74, 48, 32, 16):
    env = dict(os.environ, YALPS_GRID_CTAS=str(ctas))
    for c in cases:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "k4_case.py"), *c], env=env, capture_output=True, text=True).stdout.strip().splitlines()
        print(ctas, out[-1] if out else "failed", flush=True)
