"""Generates the committed fixtures under tests/golden/ from the reference's own test data.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py

Inputs : /root/reference/tests/cases/*.json, /root/reference/benchmarks/netlib/{index.json,cases/*.mps}
Outputs: cases.json.gz   -- the 46 test models (order-preserving pair lists), options, the reference's
                            expected {status, result}, and the oracle's trajectory data (pivot counts,
                            node counts, final basis) used as bit-level golden vectors
         netlib.npz      -- 51 Netlib models as sparse initial tableaus + index.json value + oracle outputs
The reference ships no golden data below its 1e-5 objective tolerance; the pivot counts / bases here
are the oracle's (oracle/yalps_oracle.c) after it passed the reference's own expectations.
"""
import glob
import gzip
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import model as M  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

OK = ["AGG2", "AGG3", "BEACONFD", "ISRAEL", "LOTFI", "SC105", "SC205", "SCAGR25", "SCAGR7", "SCFXM1", "SCORPION",
      "SCRS8", "SCSD6", "SCTAP1", "SCTAP2", "SCTAP3", "SHARE1B", "SHIP04L", "SHIP04S", "SHIP08L", "SHIP08S",
      "SHIP12S", "SHIP12L", "STOCFOR1", "KLEIN2"]  # benchmarks/netlib/read.ts:61-65
TIMEOUT = ["25FV47", "AGG", "BANDM", "BNL1", "BRANDY", "DEGEN2", "DEGEN3", "E226", "FFFFF800", "SCFXM2", "SCFXM3",
           "SCSD1", "SCSD8", "STOCFOR2", "WOOD1P", "KLEIN3"]  # benchmarks/netlib/read.ts:55-58
SMALL = ["AFIRO", "ADLITTLE", "SC50A", "SC50B", "BLEND", "SHARE2B", "BGPRTR", "ITEST2", "ITEST6", "KLEIN1"]


def jnum(x):
    if isinstance(x, float) and (math.isnan(x) or math.isinf(x)):
        return {"nan": "NaN", "inf": "Infinity", "-inf": "-Infinity"}[str(x)]
    return x


def cases():
    out = []
    for path in sorted(glob.glob(f"{REF}/tests/cases/*.json")):
        c = M.read_case(path)
        raw = json.load(open(path))
        info = {}
        sol = M.solve(c["model"], c["options"], info)
        assert M.valid_solution_and_status(sol, c["expected"], c["model"], c["options"]), c["name"]
        m = c["model"]
        out.append({
            "name": c["name"],
            "model": {
                "direction": m.get("direction"), "objective": m.get("objective"),
                "constraints": [[k, v] for k, v in m["constraints"]],
                "variables": [[k, [[ck, cv] for ck, cv in v]] for k, v in m["variables"]],
                "integers": raw["model"].get("integers"), "binaries": raw["model"].get("binaries"),
            },
            "options": raw.get("options") or {},
            "expected": {"status": raw["expected"]["status"], "result": raw["expected"].get("result")},
            "oracle": {
                "status": sol["status"], "result": jnum(sol["result"]),
                "variables": [[k, jnum(v)] for k, v in sol["variables"]],
                "height": info["height"], "width": info["width"],
                "root_status": info["root_status"], "root_result": jnum(info["root_result"]),
                "root_pivots": list(info["root_pivots"]), "nodes": info["nodes"], "node_pivots": info["node_pivots"],
                "final_pos": [int(x) for x in info["final_pos"]],
                "final_rhs_hex": [float(x).hex() for x in info["final_rhs"]],
            },
        })
        print(f"case {c['name']}: {sol['status']} {sol['result']}")
    with gzip.open(os.path.join(OUT, "cases.json.gz"), "wt", encoding="utf-8", compresslevel=9) as f:
        json.dump(out, f, separators=(",", ":"))


def netlib():
    idx = {e["name"]: e for e in json.load(open(f"{REF}/benchmarks/netlib/index.json"))}
    data = {}
    names = []
    for name in SMALL + OK + TIMEOUT:
        e = idx[name]
        model = M.netlib_model(open(f"{REF}/benchmarks/netlib/cases/{name.lower()}.mps").read())
        assert len(model["bounds"]) == 0, name
        opt = {**M.DEFAULT_OPTIONS, **(e.get("options") or {})}
        tm = M.tableau_model(model)
        t = tm.tableau
        init = t.matrix.copy()
        t0 = time.time()
        st, value, piv = M._lib.simplex(t.matrix, t.width, t.height, t.pos, t.var, opt["precision"], opt["maxPivots"],
                                        opt["checkCycles"])
        dt = time.time() - t0
        nzi = np.flatnonzero(init)
        names.append(name)
        data[f"{name}/shape"] = np.asarray([t.height, t.width], np.int32)
        data[f"{name}/nz_idx"] = nzi.astype(np.int64)
        data[f"{name}/nz_val"] = init[nzi]
        data[f"{name}/neg_zero_idx"] = np.flatnonzero((init == 0) & np.signbit(init)).astype(np.int64)
        data[f"{name}/row_groups"] = tm.row_groups
        data[f"{name}/check_cycles"] = np.asarray([1 if opt["checkCycles"] else 0], np.int32)
        data[f"{name}/index_value"] = np.asarray([math.nan if e["value"] is None else e["value"]], np.float64)
        data[f"{name}/status"] = np.asarray([st], np.int32)
        data[f"{name}/value"] = np.asarray([value], np.float64)
        data[f"{name}/pivots"] = np.asarray(piv, np.int64)
        data[f"{name}/final_pos"] = t.pos.copy()
        data[f"{name}/final_rhs"] = t.matrix.reshape(t.height, t.width)[:, 0].copy()
        data[f"{name}/list"] = np.asarray([0 if name in SMALL else 1 if name in OK else 2], np.int32)
        data[f"{name}/oracle_seconds"] = np.asarray([dt], np.float64)
        print(f"netlib {name}: {t.height}x{t.width} {M.STATUS_NAMES[st]} {value} {piv} {dt:.2f}s")
    data["names"] = np.asarray(names)
    np.savez_compressed(os.path.join(OUT, "netlib.npz"), **data)


if __name__ == "__main__":
    what = sys.argv[1:] or ["cases", "netlib"]
    if "cases" in what:
        cases()
    if "netlib" in what:
        netlib()
