"""oracle/model.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of everything around the simplex loop that the parity harness
needs because no JS engine exists in this image: the model -> tableau builder,
the JSON case loader, the fixed-column MPS reader, the Netlib constraint
conversion, the test PRNG, `solve()` / `solution()` and the validators.  Only
tests/, __graft_entry__.smoke() and bench.py's CPU legs import this package.

Citations are into /root/reference.  Numbers are kept as Python floats / numpy
float64 (== JS Number); -0.0 is preserved where the reference produces it.
"""
from __future__ import annotations

import json
import math
import os
import re
from typing import Any, Iterable

import numpy as np

from . import lib as _lib

INF = math.inf
STATUS_NAMES = ("optimal", "infeasible", "unbounded", "timedout", "cycled")

# src/YALPS.ts:52-60
DEFAULT_OPTIONS = {
    "precision": 1e-8,
    "checkCycles": False,
    "maxPivots": 8192,
    "tolerance": 0,
    "timeout": INF,
    "maxIterations": 32768,
    "includeZeroVariables": False,
}


# --------------------------------------------------------------------------- helpers

def js_object_entries(obj: dict) -> list:
    """Object.entries order of a JSON.parse'd object: canonical array-index keys
    ascending first, then the rest in insertion order (ECMA-262 OrdinaryOwnPropertyKeys)."""
    idx, rest = [], []
    for k, v in obj.items():
        if isinstance(k, str) and re.fullmatch(r"0|[1-9][0-9]*", k) and int(k) < 2**32 - 1:
            idx.append((int(k), k, v))
        else:
            rest.append((k, v))
    idx.sort(key=lambda t: t[0])
    return [(k, v) for _, k, v in idx] + rest


def _to_iterable(seq) -> Iterable:
    """convertToIterable (src/tableau.ts:33-38): iterables of pairs pass through, plain objects -> entries."""
    if isinstance(seq, dict):
        return js_object_entries(seq)
    return seq


def _to_set(s):
    """convertToSet (src/tableau.ts:41-45)."""
    if s is True:
        return True
    if s is False or s is None:
        return set()
    return s if isinstance(s, (set, frozenset)) else set(s)


def js_round(x: float) -> float:
    """Math.round: halves toward +inf, keeps -0."""
    if math.isnan(x) or math.isinf(x) or abs(x) >= 2.0**52:
        return x
    r = math.floor(x)
    if x - r >= 0.5:
        r += 1
    r = float(r)
    if r == 0.0 and math.copysign(1.0, x) < 0:
        r = -0.0
    return r


def round_to_precision(num: float, precision: float) -> float:
    """src/util.ts:1-4"""
    rounding = js_round(1.0 / precision)
    return js_round((num + 2.0**-52) * rounding) / rounding


# tests/helpers/util.ts:20-41
def _imul(a: int, b: int) -> int:
    return (a * b) & 0xFFFFFFFF


def prospector_hash(n: int) -> int:
    x = n & 0xFFFFFFFF
    x ^= x >> 16
    x = _imul(x, 0x21F0AAAD)
    x ^= x >> 15
    x = _imul(x, 0xD35A2D97)
    x ^= x >> 15
    return x


def hash_string(s: str) -> int:
    x = 42
    for ch in s:
        x = prospector_hash(x ^ ord(ch))
    return x


def new_rand(seed: int):
    state = [seed & 0xFFFFFFFF]

    def rand() -> float:
        state[0] = (state[0] + 0x9E3779B9) & 0xFFFFFFFF
        return prospector_hash(state[0]) / 4294967296.0

    return rand


# --------------------------------------------------------------------------- tableauModel

class Tableau:
    """src/tableau.ts:9-15"""

    __slots__ = ("matrix", "width", "height", "pos", "var")

    def __init__(self, matrix, width, height, pos, var):
        self.matrix, self.width, self.height, self.pos, self.var = matrix, width, height, pos, var


class TableauModel:
    __slots__ = ("tableau", "sign", "variables", "integers", "row_groups")

    def __init__(self, tableau, sign, variables, integers, row_groups=None):
        self.tableau, self.sign, self.variables, self.integers = tableau, sign, variables, integers
        self.row_groups = row_groups  # row -> index of the constraint key it came from (-1: objective / binary rows)


def _get(con, key):
    v = con.get(key) if isinstance(con, dict) else getattr(con, key, None)
    return v


def tableau_model(model: dict) -> TableauModel:
    """src/tableau.ts:47-137, statement for statement."""
    direction = model.get("direction")
    objective = model.get("objective")
    integers = model.get("integers")
    binaries = model.get("binaries")
    sign = -1.0 if direction == "minimize" else 1.0

    constraints_iter = _to_iterable(model["constraints"])
    variables = list(_to_iterable(model["variables"]))

    binary_col: list[int] = []
    ints: list[int] = []
    if integers is not None or binaries is not None:
        binary_vars = _to_set(binaries)
        integer_vars = True if binary_vars is True else _to_set(integers)
        for i in range(1, len(variables) + 1):
            key = variables[i - 1][0]
            if binary_vars is True or key in binary_vars:
                binary_col.append(i)
                ints.append(i)
            elif integer_vars is True or key in integer_vars:
                ints.append(i)

    constraints: dict[Any, list] = {}  # key -> [row, lower, upper]; dict keeps first-seen order like Map
    for key, con in constraints_iter:
        b = constraints.get(key)
        if b is None:
            b = [math.nan, -INF, INF]
        eq, mn, mx = _get(con, "equal"), _get(con, "min"), _get(con, "max")
        lo = eq if eq is not None else (mn if mn is not None else -INF)
        hi = eq if eq is not None else (mx if mx is not None else INF)
        b[1] = max(b[1], float(lo))
        b[2] = min(b[2], float(hi))
        if key not in constraints:
            constraints[key] = b

    num_constraints = 1
    for b in constraints.values():
        b[0] = num_constraints
        num_constraints += (1 if math.isfinite(b[1]) else 0) + (1 if math.isfinite(b[2]) else 0)

    width = len(variables) + 1
    height = num_constraints + len(binary_col)
    num_vars = width + height
    matrix = np.zeros(width * height, dtype=np.float64)
    pos = np.arange(num_vars, dtype=np.int32)
    var = np.arange(num_vars, dtype=np.int32)

    for c in range(1, width):
        for con_key, coef in _to_iterable(variables[c - 1][1]):
            coef = float(coef)
            if objective is not None and con_key == objective:
                matrix[c] = sign * coef
            b = constraints.get(con_key)
            if b is not None:
                if math.isfinite(b[2]):
                    matrix[b[0] * width + c] = coef
                    if math.isfinite(b[1]):
                        matrix[(b[0] + 1) * width + c] = -coef
                elif math.isfinite(b[1]):
                    matrix[b[0] * width + c] = -coef

    for b in constraints.values():
        if math.isfinite(b[2]):
            matrix[b[0] * width] = b[2]
            if math.isfinite(b[1]):
                matrix[(b[0] + 1) * width] = -b[1]
        elif math.isfinite(b[1]):
            matrix[b[0] * width] = -b[1]

    for k, col in enumerate(binary_col):
        row = num_constraints + k
        matrix[row * width] = 1.0
        matrix[row * width + col] = 1.0

    groups = np.full(height, -1, dtype=np.int32)
    for g, b in enumerate(constraints.values()):
        nrows = (1 if math.isfinite(b[1]) else 0) + (1 if math.isfinite(b[2]) else 0)
        groups[b[0]:b[0] + nrows] = g
    return TableauModel(Tableau(matrix, width, height, pos, var), sign, variables, ints, groups)


# --------------------------------------------------------------------------- solve (src/YALPS.ts)

def solution(tabmod: TableauModel, rhs: np.ndarray, pos: np.ndarray, var: np.ndarray, status: str, result: float,
             options: dict) -> dict:
    """src/YALPS.ts:8-50.  rhs = column 0 of the final tableau."""
    precision = options["precision"]
    width = tabmod.tableau.width
    vars_ = tabmod.variables
    if status == "optimal" or (status == "timedout" and not math.isnan(result)):
        out = []
        for i in range(len(vars_)):
            row = int(pos[i + 1]) - width
            value = float(rhs[row]) if row >= 0 else 0.0
            if value > precision:
                out.append((vars_[i][0], round_to_precision(value, precision)))
            elif options["includeZeroVariables"]:
                out.append((vars_[i][0], 0.0))
        return {"status": status, "result": -tabmod.sign * result, "variables": out}
    if status == "unbounded":
        variable = int(var[int(result)]) - 1
        return {
            "status": "unbounded",
            "result": tabmod.sign * INF,
            "variables": [(vars_[variable][0], INF)] if 0 <= variable < len(vars_) else [],
        }
    return {"status": status, "result": math.nan, "variables": []}


def solve(model: dict, options: dict | None = None, info: dict | None = None) -> dict:
    """src/YALPS.ts:73-92 on top of liboracle.so.  `info` (optional dict) receives the
    trajectory data the GPU parity tests compare against."""
    tabmod = tableau_model(model)
    opt = {**DEFAULT_OPTIONS, **(options or {})}
    t = tabmod.tableau
    st, result, pivots = _lib.simplex(t.matrix, t.width, t.height, t.pos, t.var, opt["precision"], opt["maxPivots"],
                                      opt["checkCycles"])
    status = STATUS_NAMES[st]
    if info is not None:
        info.update(root_status=status, root_result=result, root_pivots=pivots, width=t.width, height=t.height,
                    nodes=0, node_pivots=0)
    if len(tabmod.integers) == 0 or status != "optimal":
        rhs = t.matrix.reshape(t.height, t.width)[:, 0].copy() if t.width * t.height else np.zeros(0)
        if info is not None:
            info.update(final_rhs=rhs, final_pos=t.pos.copy(), final_var=t.var.copy())
        return solution(tabmod, rhs, t.pos, t.var, status, result, opt)
    bst, bres, rhs, pos, var, stats, node_log = _lib.branch_and_cut(
        t.matrix, t.width, t.height, t.pos, t.var, np.asarray(tabmod.integers, dtype=np.int32), tabmod.sign, result,
        opt["precision"], opt["maxPivots"], opt["checkCycles"], opt["tolerance"], opt["timeout"],
        opt["maxIterations"])
    if info is not None:
        info.update(nodes=int(stats[0]), node_pivots=int(stats[1]), max_cuts=int(stats[2]), max_heap=int(stats[3]),
                    node_log=node_log, final_rhs=rhs, final_pos=pos, final_var=var)
    return solution(tabmod, rhs, pos, var, STATUS_NAMES[bst], bres, opt)


# --------------------------------------------------------------------------- loaders

def read_case(path: str) -> dict:
    """tests/helpers/read.ts:41-62 for one file."""
    with open(path, "r", encoding="utf-8") as f:
        data = json.load(f)
    name = os.path.splitext(os.path.basename(path))[0]
    m = data["model"]
    constraints = js_object_entries(m["constraints"])
    variables = [(k, js_object_entries(v)) for k, v in js_object_entries(m["variables"])]
    model = dict(m)
    model.update(hash=hash_string(name), constraints=constraints, variables=variables,
                 integers=set(m.get("integers") or []), binaries=set(m.get("binaries") or []))
    options = {**DEFAULT_OPTIONS, **(data.get("options") or {})}
    exp = data["expected"]
    if exp["status"] == "optimal":
        result = float(exp["result"])
    elif exp["status"] == "unbounded":
        result = INF * (-1.0 if m.get("direction") == "minimize" else 1.0)
    else:
        result = math.nan
    return {"name": name, "model": model, "options": options, "expected": {**exp, "result": result}}


_NUM = re.compile(r"\s*([+-]?(?:Infinity|(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?))")


def js_parse_float(s: str) -> float:
    """parseFloat: longest numeric prefix, NaN otherwise."""
    m = _NUM.match(s)
    if not m:
        return math.nan
    tok = m.group(1)
    return float(tok.replace("Infinity", "inf"))


def model_from_mps(text: str, direction: str | None = None) -> dict:
    """benchmarks/mps.ts:304-325 (fixed-column reader; NAME/ROWS/COLUMNS/RHS/RANGES/BOUNDS)."""
    lines = re.split(r"\r?\n", text)
    f1 = lambda l: l[1:3].strip()
    f2 = lambda l: l[4:12].strip()
    f3 = lambda l: l[14:22].strip()
    f4 = lambda l: l[24:36].strip()
    f5 = lambda l: l[39:47].strip()
    f6 = lambda l: l[49:61].strip()

    model = {"name": "", "direction": direction, "objective": None, "constraints": {}, "variables": {},
             "integers": set(), "binaries": set(), "bounds": {}}
    ctypes_: dict[str, str] = {}
    state = {"i": 0}

    def fail(msg):
        raise ValueError(f"Line {state['i'] + 1}: {msg}")

    def next_line():
        for i in range(state["i"] + 1, len(lines)):
            if not lines[i].startswith("*"):
                state["i"] = i
                return lines[i]
        return None

    def not_end(line):
        return line is not None and line.startswith(" ")

    def section():
        return lines[state["i"]].rstrip() if state["i"] < len(lines) else None

    def got(sec):  # sectionErr, benchmarks/mps.ts:62-63
        return "end of file" if sec is None else f"'{sec}'"

    def parse_num(value, what):
        if value == "":
            fail(f"Missing {what} value")
        v = js_parse_float(value)
        if math.isnan(v):
            fail(f"Failed to parse number '{value}'")
        return v

    # NAME (:40-46)
    idx = next((i for i, l in enumerate(lines) if l.startswith("NAME")), -1)
    if idx < 0:
        fail("No NAME section was found")
    model["name"] = f3(lines[idx])
    state["i"] = idx + 1

    # ROWS (:70-98)
    if section() != "ROWS":
        fail(f"Expected section ROWS but got {got(section())}")
    line = next_line()
    while not_end(line):
        name = f2(line)
        if name == "":
            fail("Missing row name")
        if name in ctypes_:
            fail(f"The row '{name}' was already defined")
        typ = f1(line)
        if typ == "L":
            model["constraints"][name] = [-INF, 0.0]
        elif typ == "G":
            model["constraints"][name] = [0.0, INF]
        elif typ == "E":
            model["constraints"][name] = [0.0, 0.0]
        elif typ == "N":
            if model["objective"] is None:
                model["objective"] = name
            model["constraints"][name] = [-INF, INF]
        elif typ == "":
            fail("Missing row type")
        else:
            fail(f"Unexpected row type '{typ}'")
        ctypes_[name] = typ
        line = next_line()

    # COLUMNS (:114-162)
    if section() != "COLUMNS":
        fail(f"Expected section COLUMNS but got {got(section())}")

    def add_coef(variable, row, value):
        if row == "":
            fail("Missing row name")
        if value == "":
            fail("Missing coefficient value")
        if row not in ctypes_:
            fail(f"The row '{row}' was not defined in the ROWS section")
        if row in variable:
            fail(f"The coefficient for row '{row}' was previously set for this column")
        variable[row] = parse_num(value, "coefficient")

    integer_marked = False
    line = next_line()
    while not_end(line):
        if f3(line) == "'MARKER'":
            marker = f4(line)
            if marker == "'INTORG'":
                integer_marked = True
            elif marker == "'INTEND'":
                integer_marked = False
            else:
                fail(f"Unexpected MARKER '{marker}'")
            line = next_line()
            continue
        name = f2(line)
        if name == "":
            fail("Missing column name")
        if name in model["variables"]:
            fail(f"Values for the column '{name}' were previously provided")
        variable: dict[str, float] = {}
        while True:
            add_coef(variable, f3(line), f4(line))
            n2, v2 = f5(line), f6(line)
            if n2 != "" or v2 != "":
                add_coef(variable, n2, v2)
            line = next_line()
            if not (not_end(line) and f2(line) == name):
                break
        model["variables"][name] = variable
        if integer_marked:
            model["integers"].add(name)

    # RHS (:164-209)
    if section() != "RHS":
        fail(f"Expected section RHS but got {got(section())}")

    def add_constraint(row, value):
        if row == "":
            fail("Missing row name")
        if value == "":
            fail("Missing rhs value")
        typ = ctypes_.get(row)
        if typ is None:
            fail(f"The row '{row}' was not defined in the ROWS section")
        val = parse_num(value, "rhs")
        con = model["constraints"][row]
        if typ in ("L", "E"):
            con[1] = val
        if typ in ("G", "E"):
            con[0] = val

    line = next_line()
    while not_end(line):
        add_constraint(f3(line), f4(line))
        n2, v2 = f5(line), f6(line)
        if n2 != "" or v2 != "":
            add_constraint(n2, v2)
        line = next_line()

    def add_range(row, value):
        if row == "":
            fail("Missing row name")
        if value == "":
            fail("Missing range value")
        typ = ctypes_.get(row)
        if typ is None:
            fail(f"The row '{row}' was not defined in the ROWS section")
        val = parse_num(value, "range")
        b = model["constraints"][row]
        if typ == "L" or (typ == "E" and val < 0.0):
            b[0] = b[1] - abs(val)
        if typ == "G" or (typ == "E" and val > 0.0):
            b[1] = b[0] + abs(val)

    sec = section()
    if sec == "RANGES":
        line = next_line()
        while not_end(line):
            add_range(f3(line), f4(line))
            n2, v2 = f5(line), f6(line)
            if n2 != "" or v2 != "":
                add_range(n2, v2)
            line = next_line()
        sec = section()
        if sec not in ("BOUNDS", "ENDATA"):
            fail(f"Expected section BOUNDS or ENDATA but got {got(sec)}")
    elif sec not in ("BOUNDS", "ENDATA"):
        fail(f"Expected section RANGES, BOUNDS, or ENDATA but got {got(sec)}")

    if sec == "BOUNDS":  # :254-302
        def set_bounds(name, lower, upper):
            b = model["bounds"].setdefault(name, [0.0, INF])
            if not math.isnan(lower):
                b[0] = lower
            if not math.isnan(upper):
                b[1] = upper

        line = next_line()
        while not_end(line):
            typ = f1(line)
            col = f3(line)
            if col == "":
                fail("Missing column name")
            if col not in model["variables"]:
                fail(f"The column '{col}' was not defined in the COLUMNS section")
            val = math.nan
            if typ in ("LO", "UP", "FX", "LI", "UI"):
                val = parse_num(f4(line), "bound")
            if typ == "LO":
                set_bounds(col, val, INF)
            elif typ == "UP":
                set_bounds(col, 0.0, val)
            elif typ == "FX":
                set_bounds(col, val, val)
            elif typ == "FR":
                set_bounds(col, -INF, INF)
            elif typ == "MI":
                set_bounds(col, -INF, 0.0)
            elif typ == "PL":
                set_bounds(col, 0.0, INF)
            elif typ == "BV":
                model["binaries"].add(col)
            elif typ == "LI":
                model["integers"].add(col)
                set_bounds(col, val, INF)
            elif typ == "UI":
                model["integers"].add(col)
                set_bounds(col, 0.0, val)
            elif typ == "SC":
                fail("SC bound type is unsupported")
            elif typ == "":
                fail("Missing bound type")
            else:
                fail(f"Unexpected bound type '{typ}'")
            line = next_line()
        if section() != "ENDATA":
            fail(f"Expected section ENDATA but got {got(section())}")
    return model


def netlib_model(text: str) -> dict:
    """benchmarks/netlib/read.ts:16-28,38-41: modelFromMps(.., "minimize") + convertConstraints."""
    mps = model_from_mps(text, "minimize")
    cons = {}
    for key, (mn, mx) in mps["constraints"].items():
        if math.isfinite(mn) and math.isfinite(mx):
            cons[key] = {"equal": mn} if mn == mx else {"min": mn, "max": mx}
        elif math.isfinite(mn):
            cons[key] = {"min": mn}
        elif math.isfinite(mx):
            cons[key] = {"max": mx}
    out = dict(mps)
    # Maps in the reference: keep as pair lists so integer-like names are NOT reordered
    out["constraints"] = list(cons.items())
    out["variables"] = [(k, list(v.items())) for k, v in mps["variables"].items()]
    return out


# --------------------------------------------------------------------------- validators (tests/helpers/validate.ts)

MAX_DIFF = 1e-5


def _rel_from(delta, expected, precision):
    return (delta - precision) / max(abs(expected), 1.0)


def result_is_optimal(result, expected, options) -> bool:
    if math.isnan(expected):
        return math.isnan(result)
    if not math.isfinite(expected):
        return expected == result
    return math.isfinite(result) and _rel_from(abs(result - expected), expected, options["precision"]) <= max(
        options["tolerance"], MAX_DIFF)


def constraints_are_satisfied(sol, model, precision) -> bool:
    variables = dict(model["variables"])
    sums: dict = {}
    for key, num in sol["variables"]:
        for con, coef in variables[key]:
            sums[con] = num * coef + sums.get(con, 0.0)
    for key, con in model["constraints"]:
        s = sums.get(key, 0.0)
        eq, mn, mx = con.get("equal"), con.get("min"), con.get("max")
        if eq is not None:
            if _rel_from(abs(s - eq), eq, precision) > MAX_DIFF:
                return False
        else:
            if mn is not None and _rel_from(mn - s, mn, precision) > MAX_DIFF:
                return False
            if mx is not None and _rel_from(s - mx, mx, precision) > MAX_DIFF:
                return False
    return True


def variables_have_valid_values(sol, model, precision) -> bool:
    ints, bins = model["integers"], model["binaries"]
    for v, n in sol["variables"]:
        if not n >= -precision:
            return False
        if (v in ints or v in bins) and not abs(n - js_round(n)) <= precision:
            return False
        if v in bins and not n <= 1 + precision:
            return False
    return True


def valid_solution(sol, expected, model, options) -> bool:
    return (result_is_optimal(sol["result"], expected, options)
            and variables_have_valid_values(sol, model, options["precision"])
            and (not math.isfinite(expected) or constraints_are_satisfied(sol, model, options["precision"])))


def valid_solution_and_status(sol, expected, model, options) -> bool:
    if sol["status"] != expected["status"]:
        return False
    if sol["status"] == "timedout" and math.isnan(sol["result"]):
        return True
    return valid_solution(sol, expected["result"], model, options)
