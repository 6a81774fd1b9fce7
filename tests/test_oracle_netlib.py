"""Oracle vs benchmarks/netlib/index.json (the reference's expected objectives, 1e-5 relative as in
tests/additional/netlib.ts) and vs the committed trajectory vectors.  The long-running models are
checked on the GPU side only (their golden vectors were produced by this same oracle)."""
import math

import numpy as np
import pytest

from conftest import load_netlib, same_bits, same_value
from oracle import lib, model as M

NL = load_netlib()
FAST = [n for n in NL.names if float(NL.z[f"{n}/oracle_seconds"][0]) < 1.0]


@pytest.mark.parametrize("name", FAST)
def test_netlib_model(name):
    g = NL.get(name)
    m = g["matrix"].copy()
    nv = g["height"] + g["width"]
    pos, var = np.arange(nv, dtype=np.int32), np.arange(nv, dtype=np.int32)
    st, value, piv = lib.simplex(m, g["width"], g["height"], pos, var, 1e-8, 8192, g["check_cycles"])
    assert st == g["status"] and same_value(value, g["value"]) and piv == g["pivots"]
    assert np.array_equal(pos, g["final_pos"])
    assert same_bits(m.reshape(g["height"], g["width"])[:, 0], g["final_rhs"])
    if g["list"] in (0, 1):  # models the reference handles: compare with index.json like validSolution does
        options = {**M.DEFAULT_OPTIONS, "checkCycles": g["check_cycles"]}
        result = -(-1.0) * value if st == 0 else math.nan  # direction "minimize": result = -sign * x
        assert M.result_is_optimal(result, g["index_value"], options)


def test_netlib_set_is_complete():
    assert len(NL.names) == 51
    assert sum(1 for n in NL.names if NL.get(n)["list"] == 1) == 25
