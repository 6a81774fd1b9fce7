"""Fixture for the MPS BOUNDS support: Netlib models WITH a BOUNDS section, which the reference's own
harness filters out (benchmarks/netlib/read.ts:50).  Stored: the MPS text (Netlib test data), the published optimum of
benchmarks/netlib/index.json and the oracle's outcome on the transformed model (apply_bounds -> solve).

Run in the build container (needs /root/reference):  python tests/golden/make_bounded.py
Output: netlib_bounded.json.gz
"""
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import model as M  # noqa: E402
from yalps_b200 import mps as P  # noqa: E402  (only the host-side reader / transformation, no GPU)

REF = "/root/reference/benchmarks/netlib"
NAMES = ["KB2", "RECIPE", "VTP.BASE", "BORE3D", "CAPRI"]
index = {e["name"]: e for e in json.load(open(os.path.join(REF, "index.json")))}
out = []
for name in NAMES:
    text = open(os.path.join(REF, "cases", name.lower() + ".mps")).read()
    model, recover = P.apply_bounds(P.netlib_model(text))
    info = {}
    sol = recover(M.solve(model, {"maxPivots": 8192}, info))
    assert sol["status"] == "optimal" and abs(sol["result"] - index[name]["value"]) <= 1e-5 * abs(index[name]["value"]), name
    out.append({"name": name, "mps": text, "published": index[name]["value"], "oracle_status": sol["status"],
                "oracle_result": sol["result"], "bounded_columns": len(P.netlib_model(text)["bounds"])})
    print(name, sol["status"], sol["result"], index[name]["value"])
with gzip.open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "netlib_bounded.json.gz"), "wt") as f:
    json.dump(out, f)
