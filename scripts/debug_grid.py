import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import load_netlib
from oracle import lib as O
import yalps_b200
from yalps_b200 import engine as E
name = sys.argv[1] if len(sys.argv) > 1 else "SCAGR25"
g = load_netlib().get(name)
H, W = g["height"], g["width"]
eng = yalps_b200.Engine(0)
def run(path, mp):
    eng.set_tuning(path, 0)
    return eng.solve_batch(g["matrix"], H, W, E.make_options(max_pivots=mp), want_matrices=True)
def ora(mp):
    m = g["matrix"].copy().reshape(1, -1)
    r = O.simplex_batch(m, W, H, max_pivots=mp); r["matrices"] = m; return r
for rep in range(3):
    a = run(E.PATH_GRID, 8192)
    print("grid full run:", a["status"][0], a["pivots"][0], a["value"][0])
lo, hi = 0, 800
e = ora(hi); a = run(E.PATH_GRID, hi)
print("at", hi, np.array_equal(a["matrices"].view(np.uint64), e["matrices"].view(np.uint64)))
while hi - lo > 1:
    mid = (lo + hi) // 2
    e = ora(mid); a = run(E.PATH_GRID, mid)
    same = np.array_equal(a["matrices"].view(np.uint64), e["matrices"].view(np.uint64)) and np.array_equal(a["pivots"], e["pivots"])
    if same: lo = mid
    else: hi = mid
print("first differing max_pivots:", hi)
e = ora(hi); a = run(E.PATH_GRID, hi); k2 = run(E.PATH_GMEM, hi)
print("k2 same as oracle:", np.array_equal(k2["matrices"].view(np.uint64), e["matrices"].view(np.uint64)))
d = np.flatnonzero(a["matrices"].view(np.uint64).reshape(-1) != e["matrices"].view(np.uint64).reshape(-1))
print("ndiff", d.size, "rows", np.unique(d // W)[:20], "cols", np.unique(d % W)[:20])
print("pivots", a["pivots"], e["pivots"])
for i in d[:10]: print(i // W, i % W, a["matrices"].reshape(-1)[i], e["matrices"].reshape(-1)[i])
