"""Fixed-column MPS reader (host side; SURVEY 8f-1: the data format in front of the hot path).

Same accepted dialect and error behaviour as the reference's benchmark reader `modelFromMps`
(benchmarks/mps.ts:304-325): sections NAME, ROWS, COLUMNS (with MARKER INTORG/INTEND), RHS, optional RANGES and
BOUNDS, ENDATA; fields at the fixed columns of benchmarks/mps.ts:31-36; `*` comment lines; the first N row is the
objective and RHS entries on N rows are ignored; numbers are parsed like JS parseFloat (longest numeric prefix).
`netlib_model` applies the Netlib conversion of benchmarks/netlib/read.ts:16-28 (min == max -> equal, free rows
dropped, direction "minimize") and returns a model dict for `yalps_b200.solve`.

Implementation note: a table-driven single pass over the lines (section -> handler), not the reference's chain of
reader functions.
"""
from __future__ import annotations

import math
import re
from typing import Optional

INF = math.inf
_FIELDS = ((1, 3), (4, 12), (14, 22), (24, 36), (39, 47), (49, 61))  # benchmarks/mps.ts:31-36
_FLOAT = re.compile(r"\s*([+-]?(?:Infinity|(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?))")


class MpsError(ValueError):
    pass


def _parse_float(text: str) -> float:
    m = _FLOAT.match(text)
    return float(m.group(1).replace("Infinity", "inf")) if m else math.nan


def _fields(line: str):
    return tuple(line[a:b].strip() for a, b in _FIELDS)


def model_from_mps(text: str, direction: Optional[str] = None) -> dict:
    lines = re.split(r"\r?\n", text)
    start = next((i for i, l in enumerate(lines) if l.startswith("NAME")), None)
    if start is None:
        raise MpsError("Line 1: No NAME section was found")
    model = {"name": _fields(lines[start])[2], "direction": direction, "objective": None, "constraints": {},
             "variables": {}, "bounds": {}, "column_bounds": {}, "integers": set(), "binaries": set()}
    row_type: dict = {}
    state = {"section": None, "int": False, "col": None, "lineno": start + 1, "bound_lower_seen": set()}

    def fail(msg):
        raise MpsError(f"Line {state['lineno'] + 1}: {msg}")

    def number(value, what):
        if value == "":
            fail(f"Missing {what} value")
        v = _parse_float(value)
        if math.isnan(v):
            fail(f"Failed to parse number '{value}'")
        return v

    def known_row(row):
        if row == "":
            fail("Missing row name")
        if row not in row_type:
            fail(f"The row '{row}' was not defined in the ROWS section")
        return row_type[row]

    def on_rows(f):
        typ, name = f[0], f[1]
        if name == "":
            fail("Missing row name")
        if name in row_type:
            fail(f"The row '{name}' was already defined")
        if typ == "":
            fail("Missing row type")
        if typ not in ("L", "G", "E", "N"):
            fail(f"Unexpected row type '{typ}'")
        model["constraints"][name] = {"L": [-INF, 0.0], "G": [0.0, INF], "E": [0.0, 0.0], "N": [-INF, INF]}[typ]
        if typ == "N" and model["objective"] is None:
            model["objective"] = name
        row_type[name] = typ

    def on_columns(f):
        if f[2] == "'MARKER'":
            if f[3] not in ("'INTORG'", "'INTEND'"):
                fail(f"Unexpected MARKER '{f[3]}'")
            state["int"] = f[3] == "'INTORG'"
            state["col"] = None
            return
        name = f[1]
        if name == "":
            fail("Missing column name")
        if name != state["col"]:
            if name in model["variables"]:
                fail(f"Values for the column '{name}' were previously provided -- all values for a column must come "
                     "consecutively")
            model["variables"][name] = {}
            if state["int"]:
                model["integers"].add(name)
            state["col"] = name
        coefs = model["variables"][name]
        for row, value in ((f[2], f[3]), (f[4], f[5])):
            if (row, value) == ("", "") and row is f[4]:
                continue
            if row == "":
                fail("Missing row name")
            if value == "":
                fail("Missing coefficient value")
            known_row(row)
            if row in coefs:
                fail(f"The coefficient for row '{row}' was previously set for this column")
            coefs[row] = number(value, "coefficient")

    def on_rhs(f):
        for row, value in ((f[2], f[3]), (f[4], f[5])):
            if (row, value) == ("", "") and row is f[4]:
                continue
            typ = known_row(row)
            v = number(value, "rhs")
            con = model["constraints"][row]
            if typ in ("L", "E"):
                con[1] = v
            if typ in ("G", "E"):
                con[0] = v

    def on_ranges(f):
        for row, value in ((f[2], f[3]), (f[4], f[5])):
            if (row, value) == ("", "") and row is f[4]:
                continue
            typ = known_row(row)
            v = number(value, "range")
            b = model["constraints"][row]
            if typ == "L" or (typ == "E" and v < 0.0):
                b[0] = b[1] - abs(v)
            if typ == "G" or (typ == "E" and v > 0.0):
                b[1] = b[0] + abs(v)

    def on_bounds(f):
        typ, col = f[0], f[2]
        if col == "":
            fail("Missing column name")
        if col not in model["variables"]:
            fail(f"The column '{col}' was not defined in the COLUMNS section")
        v = number(f[3], "bound") if typ in ("LO", "UP", "FX", "LI", "UI") else math.nan
        lo_hi = {"LO": (v, INF), "UP": (0.0, v), "FX": (v, v), "FR": (-INF, INF), "MI": (-INF, 0.0), "PL": (0.0, INF),
                 "LI": (v, INF), "UI": (0.0, v)}
        if typ == "BV":
            model["binaries"].add(col)
            return
        if typ == "SC":
            fail("SC bound type is unsupported")
        if typ == "":
            fail("Missing bound type")
        if typ not in lo_hi:
            fail(f"Unexpected bound type '{typ}'")
        if typ in ("LI", "UI"):
            model["integers"].add(col)
        b = model["bounds"].setdefault(col, [0.0, INF])
        lo, hi = lo_hi[typ]
        if not math.isnan(lo):
            b[0] = lo
        if not math.isnan(hi):
            b[1] = hi
        # `bounds` above is the reference reader's table, quirk included: every BOUNDS line rewrites BOTH sides
        # (benchmarks/mps.ts:254-258,287-288), so `LO x 1` + `UP x 4` ends as [0, 4] or [1, inf) depending on the line
        # order.  The reference never reads it (benchmarks/netlib/read.ts:50 drops such models); apply_bounds() does,
        # so the usual MPS meaning -- one side per line -- is kept separately in `column_bounds`.
        cb = model["column_bounds"].setdefault(col, [0.0, INF])
        seen = state["bound_lower_seen"]
        if typ in ("LO", "LI"):
            cb[0] = v
            seen.add(col)
        elif typ in ("UP", "UI"):
            cb[1] = v
            if v < 0.0 and col not in seen:  # the common convention: a negative upper bound without a lower one
                cb[0] = -INF
        elif typ == "FX":
            cb[0] = cb[1] = v
            seen.add(col)
        elif typ == "FR":
            cb[0], cb[1] = -INF, INF
            seen.add(col)
        elif typ == "MI":
            cb[0] = -INF
            seen.add(col)
        elif typ == "PL":
            cb[1] = INF

    handlers = {"ROWS": on_rows, "COLUMNS": on_columns, "RHS": on_rhs, "RANGES": on_ranges, "BOUNDS": on_bounds}
    order = {"ROWS": ("COLUMNS",), "COLUMNS": ("RHS",), "RHS": ("RANGES", "BOUNDS", "ENDATA"),
             "RANGES": ("BOUNDS", "ENDATA"), "BOUNDS": ("ENDATA",)}
    wording = {("ROWS",): "ROWS", ("COLUMNS",): "COLUMNS", ("RHS",): "RHS",
               ("RANGES", "BOUNDS", "ENDATA"): "RANGES, BOUNDS, or ENDATA", ("BOUNDS", "ENDATA"): "BOUNDS or ENDATA",
               ("ENDATA",): "ENDATA"}
    expected = ("ROWS",)
    last = None  # last non-comment line seen: what the reference reports as the "section" at end of file
    for i in range(start + 1, len(lines)):
        line = lines[i]
        if line.startswith("*"):
            continue
        state["lineno"] = last = i
        if line.startswith(" "):
            if state["section"] is None:
                fail(f"Expected section {wording[expected]} but got '{line.rstrip()}'")
            handlers[state["section"]](_fields(line))
            continue
        name = line.rstrip()
        if name not in expected:
            fail(f"Expected section {wording[expected]} but got '{name}'")
        if name == "ENDATA":
            return model
        state["section"] = name
        expected = order[name]
    if last is None:
        state["lineno"] = start + 1
        fail(f"Expected section {wording[expected]} but got end of file")
    state["lineno"] = last
    fail(f"Expected section {wording[expected]} but got '{lines[last].rstrip()}'")


def netlib_model(text: str) -> dict:
    """benchmarks/netlib/read.ts:16-28,38-41: the model dict solve() takes (pair lists keep MPS order)."""
    mps = model_from_mps(text, "minimize")
    constraints = []
    for key, (lo, hi) in mps["constraints"].items():
        if math.isfinite(lo) and math.isfinite(hi):
            constraints.append((key, {"equal": lo} if lo == hi else {"min": lo, "max": hi}))
        elif math.isfinite(lo):
            constraints.append((key, {"min": lo}))
        elif math.isfinite(hi):
            constraints.append((key, {"max": hi}))
    return {"name": mps["name"], "direction": "minimize", "objective": mps["objective"], "constraints": constraints,
            "variables": [(k, list(v.items())) for k, v in mps["variables"].items()], "integers": mps["integers"],
            "binaries": mps["binaries"], "bounds": mps["bounds"], "column_bounds": mps["column_bounds"]}


def apply_bounds(model: dict) -> tuple:
    """Column bounds (the MPS BOUNDS section) expressed in a model solve() can take.

    The reference's variables are implicitly >= 0 and its Netlib harness drops every model with a BOUNDS section
    (benchmarks/netlib/read.ts:50); this is the host-side transformation SURVEY 8(f)-4 asks for.  For a column x
    with bounds [l, u] (default [0, inf), benchmarks/mps.ts:254-258):

      l finite   x = l + x'            x' >= 0; every constraint bound and the objective shift by coef * l;
                                       u finite adds the row  x' <= u - l  (a fixed column, l == u, becomes a constant)
      l = -inf   x = u - x' (u finite) x' >= 0, coefficients negated, shifts by coef * u
                 x = x+ - x-  (free)   two non-negative columns

    Returns (model without "bounds", recover) where recover(solution) maps a Solution of the transformed model back
    to the original variables and objective value.  `model` is a netlib_model()-style dict (pair lists).
    """
    # `column_bounds` (one side per BOUNDS line) when the model came from model_from_mps(); a hand-made model may give
    # the same table under "bounds"
    bounds = model.get("column_bounds") if model.get("column_bounds") is not None else (model.get("bounds") or {})
    integers = set(model.get("integers") or ())
    binaries = set(model.get("binaries") or ())
    new_integers = set(integers)
    shift = {}       # constraint key -> sum of coef * offset (moved to the right-hand side)
    obj_const = 0.0
    variables, extra_rows, back = [], [], {}
    for name, coefs in model["variables"]:
        lo, hi = bounds.get(name, (0.0, math.inf))
        coefs = list(coefs)
        if name in integers and name not in binaries:
            # an integer column only takes integer values: tightening its bounds to integers changes nothing and
            # keeps the shifted column x' = x - l integral
            lo = float(math.ceil(lo)) if math.isfinite(lo) else lo
            hi = float(math.floor(hi)) if math.isfinite(hi) else hi
        elif name in binaries and (lo, hi) != (0.0, math.inf):
            raise ValueError(f"column {name}: a binary column cannot carry other bounds")
        if lo > hi:
            raise ValueError(f"column {name}: lower bound {lo} above upper bound {hi}")
        if math.isfinite(lo):
            offset, sign = lo, 1.0
        elif math.isfinite(hi):
            offset, sign = hi, -1.0
        else:  # free column: x = x+ - x-
            variables.append((name, coefs))
            variables.append((name + "__neg", [(k, -c) for k, c in coefs]))
            back[name] = ("free", name, name + "__neg")
            if name in integers:  # x integer <=> both parts can be taken integer
                new_integers.add(name + "__neg")
            continue
        if offset != 0.0:
            for k, c in coefs:
                if k == model["objective"]:
                    obj_const += c * offset
                shift[k] = shift.get(k, 0.0) + c * offset
        if sign < 0:
            coefs = [(k, -c) for k, c in coefs]
        width = hi - lo if sign > 0 else math.inf  # range left for x'
        if width == 0.0:
            back[name] = ("const", offset)
            continue  # fixed column: only its constant contribution remains
        if math.isfinite(width):
            row = name + "__ub"
            coefs.append((row, 1.0))
            extra_rows.append((row, {"max": width}))
        variables.append((name, coefs))
        back[name] = ("affine", offset, sign)
    constraints = []
    for key, con in model["constraints"]:
        s = shift.get(key, 0.0)
        constraints.append((key, {f: v - s for f, v in con.items()} if s != 0.0 else dict(con)))
    out = {k: v for k, v in model.items() if k not in ("bounds", "column_bounds")}
    out["constraints"] = constraints + extra_rows
    out["variables"] = variables
    out["integers"] = new_integers - {n for n, kind in back.items() if kind[0] == "const"}

    def recover(solution: dict) -> dict:
        sol = dict(solution)
        if solution["status"] not in ("optimal", "timedout") or solution["result"] != solution["result"]:
            return sol
        vals = dict(solution["variables"])
        orig = []
        for name, _ in model["variables"]:
            kind = back[name]
            if kind[0] == "const":
                x = kind[1]
            elif kind[0] == "free":
                x = vals.get(kind[1], 0.0) - vals.get(kind[2], 0.0)
            else:
                x = kind[1] + kind[2] * vals.get(name, 0.0)
            if x != 0.0:
                orig.append([name, x])
        sol["variables"] = orig
        sol["result"] = solution["result"] + obj_const
        return sol

    return out, recover
