import gzip
import json
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _unjson(x):
    if x == "NaN":
        return math.nan
    if x == "Infinity":
        return math.inf
    if x == "-Infinity":
        return -math.inf
    return x


_cases_cache = None


def load_cases():
    """tests/golden/cases.json.gz -> list of dicts (model in order-preserving pair lists)."""
    global _cases_cache
    if _cases_cache is None:
        with gzip.open(os.path.join(GOLDEN, "cases.json.gz"), "rt", encoding="utf-8") as f:
            cases = json.load(f)
        for c in cases:
            o = c["oracle"]
            o["result"] = _unjson(o["result"])
            o["root_result"] = _unjson(o["root_result"])
            o["variables"] = [[k, _unjson(v)] for k, v in o["variables"]]
            o["final_rhs"] = np.asarray([float.fromhex(h) for h in o.pop("final_rhs_hex")], np.float64)
            o["final_pos"] = np.asarray(o["final_pos"], np.int32)
            m = c["model"]
            m["constraints"] = [(k, v) for k, v in m["constraints"]]
            m["variables"] = [(k, [(ck, cv) for ck, cv in v]) for k, v in m["variables"]]
            if m.get("integers") is None:
                m.pop("integers", None)
            if m.get("binaries") is None:
                m.pop("binaries", None)
            if m.get("direction") is None:
                m.pop("direction", None)
            if m.get("objective") is None:
                m.pop("objective", None)
        _cases_cache = cases
    return _cases_cache


def case_expected_result(c):
    """tests/helpers/read.ts:54-58"""
    e = c["expected"]
    if e["status"] == "optimal":
        return float(e["result"])
    if e["status"] == "unbounded":
        return math.inf * (-1.0 if c["model"].get("direction") == "minimize" else 1.0)
    return math.nan


_netlib_cache = None


class NetlibGolden:
    def __init__(self):
        self.z = np.load(os.path.join(GOLDEN, "netlib.npz"))
        self.names = [str(n) for n in self.z["names"]]

    def get(self, name):
        z = self.z
        h, w = (int(x) for x in z[f"{name}/shape"])
        m = np.zeros(h * w, np.float64)
        m[z[f"{name}/nz_idx"]] = z[f"{name}/nz_val"]
        m[z[f"{name}/neg_zero_idx"]] = -0.0
        return {
            "name": name, "height": h, "width": w, "matrix": m,
            "check_cycles": bool(z[f"{name}/check_cycles"][0]), "index_value": float(z[f"{name}/index_value"][0]),
            "status": int(z[f"{name}/status"][0]), "value": float(z[f"{name}/value"][0]),
            "pivots": tuple(int(x) for x in z[f"{name}/pivots"]), "final_pos": z[f"{name}/final_pos"],
            "final_rhs": z[f"{name}/final_rhs"], "list": int(z[f"{name}/list"][0]),
            "oracle_seconds": float(z[f"{name}/oracle_seconds"][0]),
            "row_groups": z[f"{name}/row_groups"],
        }


def load_netlib():
    global _netlib_cache
    if _netlib_cache is None:
        _netlib_cache = NetlibGolden()
    return _netlib_cache


def bits(a):
    """float64 array -> uint64 view for bit-exact comparison (NaNs and signed zeros included)."""
    return np.ascontiguousarray(a, np.float64).view(np.uint64)


def same_bits(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return a.shape == b.shape and np.array_equal(bits(a), bits(b))


def same_value(a, b):
    """bit-equal, except that every NaN equals every NaN"""
    a, b = float(a), float(b)
    if math.isnan(a) or math.isnan(b):
        return math.isnan(a) and math.isnan(b)
    return a.hex() == b.hex()


@pytest.fixture(scope="session")
def engine():
    import yalps_b200
    eng = yalps_b200.Engine(0)
    yield eng
    eng.close()
