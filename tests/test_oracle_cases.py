"""Pins the oracle (oracle/yalps_oracle.c + oracle/model.py) on the reference's own fixtures:
every tests/cases/*.json expectation (status exact, objective within the reference's validator tolerance,
tests/helpers/validate.ts) and the committed trajectory vectors.  CPU only."""
import math

import numpy as np
import pytest

from conftest import case_expected_result, load_cases, same_bits, same_value
from oracle import model as M

CASES = load_cases()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_case_matches_reference_expectation(case):
    options = {**M.DEFAULT_OPTIONS, **case["options"]}
    info = {}
    sol = M.solve(case["model"], options, info)
    model = dict(case["model"])
    model["integers"] = set(model.get("integers") or [])
    model["binaries"] = set(model.get("binaries") or [])
    expected = {"status": case["expected"]["status"], "result": case_expected_result(case)}
    assert M.valid_solution_and_status(sol, expected, model, options)  # tests/solver.ts:23-25
    o = case["oracle"]
    assert sol["status"] == o["status"] and same_value(sol["result"], o["result"])
    assert list(info["root_pivots"]) == o["root_pivots"]
    assert info["nodes"] == o["nodes"] and info["node_pivots"] == o["node_pivots"]
    assert np.array_equal(info["final_pos"], o["final_pos"])
    assert same_bits(info["final_rhs"], o["final_rhs"])


def test_status_census():
    """38 optimal, 5 infeasible, 1 unbounded, 2 cycled."""
    from collections import Counter
    c = Counter(x["expected"]["status"] for x in CASES)
    assert c == {"optimal": 38, "infeasible": 5, "unbounded": 1, "cycled": 2}


def test_readme_example():
    """README.md:63-79 of the reference."""
    model = {
        "direction": "maximize", "objective": "profit",
        "constraints": {"wood": {"max": 300}, "labor": {"max": 110}, "storage": {"max": 400}},
        "variables": {"table": {"wood": 30, "labor": 5, "profit": 1200, "storage": 30},
                      "dresser": {"wood": 20, "labor": 10, "profit": 1600, "storage": 50}},
        "integers": ["table", "dresser"],
    }
    sol = M.solve(model)
    assert sol == {"status": "optimal", "result": 14400.0, "variables": [("table", 8.0), ("dresser", 3.0)]}


def test_round_to_precision_matches_c_and_python():
    from oracle import lib
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.normal(size=200) * 10.0 ** rng.integers(-9, 9, 200), [0.5e-8, -0.5e-8, 1.5e-8, -0.0, 0.0,
                         2.5, -2.5, 1e300, math.inf, math.nan]])
    for x in xs:
        for p in (1e-8, 1e-5, 0.5, 3e-9):
            assert same_value(lib.round_to_precision(float(x), p), M.round_to_precision(float(x), p))


def test_js_round_half_up():
    assert M.js_round(2.5) == 3.0 and M.js_round(-2.5) == -2.0 and M.js_round(-0.5) == 0.0
    assert math.copysign(1.0, M.js_round(-0.2)) == -1.0
    assert M.js_round(0.49999999999999994) == 0.0
