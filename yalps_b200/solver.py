"""solve(model, options) / solve_many(models, options): the public API of the reference (src/index.ts:1-3,
src/YALPS.ts:73-92) in front of the B200 engine.

Model, Options and Solution keep the reference's shapes and spellings (src/types.ts):
  model   = {"direction", "objective", "constraints", "variables", "integers", "binaries"}
  options = {"precision", "checkCycles", "maxPivots", "tolerance", "timeout", "maxIterations",
             "includeZeroVariables"}
  solution = {"status": "optimal"|"infeasible"|"unbounded"|"timedout"|"cycled", "result", "variables"}
The tableau is built on the host (tableau.py), the simplex and branch-and-cut numerics run on the GPU
through the C ABI; there is no CPU solve path.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np

from .engine import Engine, MultiEngine, STATUS_NAMES, make_options
from .tableau import SPARSE_OVER_BYTES, TableauModel, tableau_model

# src/YALPS.ts:52-60
_DEFAULTS = {
    "precision": 1e-8,
    "checkCycles": False,
    "maxPivots": 8192,
    "tolerance": 0,
    "timeout": math.inf,
    "maxIterations": 32768,
    "includeZeroVariables": False,
}

#: exported mutable copy, like `defaultOptions` (src/YALPS.ts:65); editing it does not change solve()
default_options = dict(_DEFAULTS)
defaultOptions = default_options

_engines: dict = {}


def get_engine(device: int = 0) -> Engine:
    """Process-wide engine per device (created on first use; raises if there is no CUDA device)."""
    eng = _engines.get(device)
    if eng is None:
        eng = _engines[device] = Engine(device)
    return eng


def _js_round(x: float) -> float:
    if math.isnan(x) or math.isinf(x) or abs(x) >= 2.0 ** 52:
        return x
    r = float(math.floor(x))
    if x - r >= 0.5:
        r += 1.0
    return -0.0 if r == 0.0 and math.copysign(1.0, x) < 0 else r


def round_to_precision(num: float, precision: float) -> float:
    """src/util.ts:1-4 (host copy, used only to format reported variable values)."""
    rounding = _js_round(1.0 / precision)
    return _js_round((num + 2.0 ** -52) * rounding) / rounding


def _c_options(opt: dict):
    return make_options(opt["precision"], opt["maxPivots"], opt["checkCycles"], opt["tolerance"], opt["timeout"],
                        opt["maxIterations"])


def _solution(tabmod: TableauModel, status: str, result: float, rhs, pos, var, opt: dict) -> dict:
    """src/YALPS.ts:8-50: needs only column 0 and the basis arrays of the final tableau."""
    width = tabmod.tableau.width
    precision = opt["precision"]
    names = tabmod.variables
    if status == "optimal" or (status == "timedout" and not math.isnan(result)):
        out = []
        for i, (key, _) in enumerate(names):
            row = int(pos[i + 1]) - width
            value = float(rhs[row]) if row >= 0 else 0.0
            if value > precision:
                out.append([key, round_to_precision(value, precision)])
            elif opt["includeZeroVariables"]:
                out.append([key, 0.0])
        return {"status": status, "result": -tabmod.sign * result, "variables": out}
    if status == "unbounded":
        v = int(var[int(result)]) - 1
        return {"status": "unbounded", "result": tabmod.sign * math.inf,
                "variables": [[names[v][0], math.inf]] if 0 <= v < len(names) else []}
    return {"status": status, "result": math.nan, "variables": []}


def solve(model: dict, options: Optional[dict] = None, *, engine: Optional[Engine] = None,
          info: Optional[dict] = None) -> dict:
    """Runs the solver on `model` (src/YALPS.ts:73-92).  `info`, if given, receives engine statistics."""
    opt = {**_DEFAULTS, **(options or {})}
    eng = engine or get_engine()
    # big model tableaus are almost all zeros: they go to the device as (cell, value) pairs (yalps_solve_sparse) and
    # never exist densely on the host; small ones take the library's zero-copy path as a dense image
    sparse_ok = hasattr(eng, "solve_tableau_sparse")
    tabmod = tableau_model(model, SPARSE_OVER_BYTES if sparse_ok else None)
    t = tabmod.tableau
    if t.matrix is None:
        r = eng.solve_tableau_sparse(t.cells, t.values, t.height, t.width, tabmod.integers, tabmod.sign, _c_options(opt))
    else:
        # a MultiEngine shards the branch-and-bound frontier over its GPUs (same search, same result)
        r = eng.solve_tableau(t.matrix, t.height, t.width, tabmod.integers, tabmod.sign, _c_options(opt))
    if info is not None:
        info.update(r["stats"], root_status=STATUS_NAMES[r["root_status"]], root_value=r["root_value"],
                    root_pivots=r["root_pivots"], height=t.height, width=t.width, final_rhs=r["rhs"],
                    final_pos=r["pos"], final_var=r["var"])
    return _solution(tabmod, STATUS_NAMES[r["status"]], r["result"], r["rhs"], r["pos"], r["var"], opt)


def _worker_engines(device: int, count: int) -> list:
    """Extra engines (one ctx each, same GPU) for concurrent branch-and-cut searches; a ctx is single-threaded."""
    out = []
    for k in range(count):
        key = (device, "milp", k)
        eng = _engines.get(key)
        if eng is None:
            eng = _engines[key] = Engine(device)
        out.append(eng)
    return out


def solve_many(models: Sequence[dict], options: Optional[dict] = None, *, engine: Optional[Engine] = None,
               milp_workers: int = 4) -> list:
    """solveMany(models, options): all root LPs go to the device as ONE ragged batch (LPs of different sizes are
    grouped per kernel configuration by the library).  Models with integer variables whose root is optimal then
    run branch and cut; their searches are latency-bound node waves, so up to `milp_workers` of them run
    concurrently, each on its own context / streams of the same GPU (the C ABI releases the GIL for a whole search).
    Results are identical to calling solve() per model."""
    opt = {**_DEFAULTS, **(options or {})}
    copt = _c_options(opt)
    eng = engine or get_engine()
    tabmods = [tableau_model(m) for m in models]
    if not tabmods:
        return []
    if isinstance(eng, MultiEngine):  # one C-ABI call: roots sharded over the GPUs, searches dealt to worker contexts
        res = eng.solve_many_tableaus(tabmods, copt, searches_per_device=milp_workers)
        return [_solution(tm, STATUS_NAMES[r["status"]], r["result"], r["rhs"], r["pos"], r["var"], opt)
                for tm, r in zip(tabmods, res)]
    roots = eng.solve_ragged([tm.tableau.matrix for tm in tabmods],
                             [(tm.tableau.height, tm.tableau.width) for tm in tabmods], copt,
                             want_matrices=any(tm.integers for tm in tabmods))
    out: list = [None] * len(tabmods)
    milp = []
    for i, (tm, r) in enumerate(zip(tabmods, roots)):
        status = STATUS_NAMES[r["status"]]
        if not tm.integers or status != "optimal":
            out[i] = _solution(tm, status, r["value"], r["rhs"], r["pos"], r["var"], opt)
        else:
            milp.append(i)

    def search(e: Engine, i: int):
        tm, r = tabmods[i], roots[i]
        t = tm.tableau
        e.bnb_set_root(r["matrix"], t.height, t.width, r["pos"], r["var"], 2 * len(tm.integers))
        b = e.branch_and_cut(tm.integers, tm.sign, r["value"], copt)
        out[i] = _solution(tm, STATUS_NAMES[b["status"]], b["result"], b["rhs"], b["pos"], b["var"], opt)

    workers = max(1, min(int(milp_workers), len(milp)))
    if workers <= 1:
        for i in milp:
            search(eng, i)
    else:
        import queue
        import threading
        todo: "queue.SimpleQueue[int]" = queue.SimpleQueue()
        for i in milp:
            todo.put(i)
        errors: list = []

        def run(e: Engine):
            while True:
                try:
                    i = todo.get_nowait()
                except queue.Empty:
                    return
                try:
                    search(e, i)
                except Exception as exc:  # surfaced after the join
                    errors.append(exc)
                    return

        threads = [threading.Thread(target=run, args=(e,)) for e in _worker_engines(eng.device, workers)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        if errors:
            raise errors[0]
    return out


solveMany = solve_many
