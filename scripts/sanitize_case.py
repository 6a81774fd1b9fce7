"""Small mixed workload for compute-sanitizer (one tool per gpurun call): every kernel path once."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import yalps_b200
from yalps_b200 import engine as E
from oracle import lib as O
from conftest import load_cases
eng = yalps_b200.Engine(0)
for (m, nv, neg, n) in [(32, 64, 4, 24), (7, 40, 3, 8), (40, 7, 10, 8)]:
    mats = O.generate_synthetic(0, n, m, nv, neg)
    exp = O.simplex_batch(mats.copy(), nv + 1, m + 1)
    for path, thr in ((E.PATH_SMEM, 32), (E.PATH_SMEM, 128), (E.PATH_GMEM, 64), (E.PATH_GRID, 0)):
        eng.set_tuning(path, thr)
        got = eng.solve_batch(mats[: (2 if path == E.PATH_GRID else n)], m + 1, nv + 1, want_matrices=True)
        k = got["status"].shape[0]
        assert np.array_equal(got["pivots"], exp["pivots"][:k]), (m, nv, path)
eng.set_tuning(0, 0)
t = np.array([[0, 10, -57, -9, -24], [0, 0.5, -5.5, -2.5, 9], [0, 0.5, -1.5, -0.5, 1], [1, 1, 0, 0, 0]], float)
assert eng.solve_batch(t.reshape(1, -1), 4, 5, E.make_options(check_cycles=True))["status"][0] == 4
c = next(x for x in load_cases() if x["name"] == "Knapsack 1")
assert yalps_b200.solve(c["model"], engine=eng)["result"] == 7534.0
print("sanitize case ok")
eng.close()
