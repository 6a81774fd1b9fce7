"""Where the time of one solve() of AFIRO goes (config 1): python scripts/afiro_latency.py"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import yalps_b200
from yalps_b200 import solver, engine as E
from yalps_b200.mps import netlib_model
model = netlib_model(open("tests/golden/afiro.mps").read())
eng = yalps_b200.Engine(0)
opt = {**solver._DEFAULTS}
copt = solver._c_options(opt)
def t(fn, reps=200):
    for _ in range(20): fn()
    t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    return (time.perf_counter() - t0) / reps * 1e6, r
us_total, _ = t(lambda: yalps_b200.solve(model, engine=eng))
us_tab, tm = t(lambda: yalps_b200.tableau_model(model, solver.SPARSE_OVER_BYTES))
tb = tm.tableau
us_eng, r = t(lambda: eng.solve_tableau(tb.matrix, tb.height, tb.width, tm.integers, tm.sign, copt))
us_sol, _ = t(lambda: solver._solution(tm, "optimal", r["result"], r["rhs"], r["pos"], r["var"], opt))
us_opt, _ = t(lambda: solver._c_options(opt))
mats = np.ascontiguousarray(tb.matrix).reshape(1, -1)
us_batch, _ = t(lambda: eng.solve_batch(mats, tb.height, tb.width, copt))
print(f"solve {us_total:.1f} us = tableau_model {us_tab:.1f} + c_options {us_opt:.1f} + Engine.solve_tableau {us_eng:.1f} + solution {us_sol:.1f}; Engine.solve_batch(n=1) {us_batch:.1f}")
eng.close()
