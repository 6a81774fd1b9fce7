"""Host-side model -> tableau builder (yalps_b200/tableau.py) against the oracle's literal restatement of
src/tableau.ts and against the layout properties tests/tableau.ts pins (bit-exact, signed zeros included)."""
import numpy as np
import pytest

from conftest import load_cases, same_bits
from oracle import model as M
from yalps_b200.tableau import tableau_model

CASES = [c for c in load_cases() if c["name"] not in ("Monster 2", "Monster Problem", "Vendor Selection")]
ALL = load_cases()


def build(model):
    tm = tableau_model(model)
    t = tm.tableau
    return tm, t


def assert_same(a, b):
    ta, tb = a.tableau, b.tableau
    assert (ta.width, ta.height) == (tb.width, tb.height)
    mb = tb.matrix
    assert same_bits(ta.matrix, mb)
    pa = ta.position_of_variable
    pb = tb.position_of_variable if hasattr(tb, "position_of_variable") else tb.pos
    va = ta.variable_at_position
    vb = tb.variable_at_position if hasattr(tb, "variable_at_position") else tb.var
    assert np.array_equal(pa, pb) and np.array_equal(va, vb)
    assert a.sign == b.sign and list(a.integers) == list(b.integers)
    assert [k for k, _ in a.variables] == [k for k, _ in b.variables]


@pytest.mark.parametrize("case", ALL, ids=[c["name"] for c in ALL])
def test_matches_oracle_builder(case):
    assert_same(tableau_model(case["model"]), M.tableau_model(case["model"]))


def test_empty_model():
    """tests/tableau.ts:12-27"""
    tm = tableau_model({"variables": {}, "constraints": {}})
    t = tm.tableau
    assert t.width == 1 and t.height == 1 and same_bits(t.matrix, np.zeros(1))
    assert t.position_of_variable.tolist() == [0, 1] and t.variable_at_position.tolist() == [0, 1]
    assert tm.sign == 1.0 and tm.variables == [] and tm.integers == []


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_layout_properties(case):
    model = case["model"]
    base = tableau_model(model)
    W = base.tableau.width
    rand = M.new_rand(M.hash_string(case["name"]))

    # tests/tableau.ts:49-54: no objective -> objective row is zero
    no_obj = tableau_model({k: v for k, v in model.items() if k != "objective"})
    exp = base.tableau.matrix.copy()
    exp[:W] = 0.0
    assert same_bits(no_obj.tableau.matrix, exp)

    # :56-67: opposite direction negates the objective row (and sign)
    other = "maximize" if model.get("direction") == "minimize" else "minimize"
    flipped = tableau_model({**model, "direction": other})
    exp = base.tableau.matrix.copy()
    exp[:W] = np.where(exp[:W] == 0.0, 0.0, -exp[:W])
    assert same_bits(flipped.tableau.matrix, exp) and flipped.sign == -base.sign

    # :104-133: object / array / Map forms are equivalent
    as_dict = tableau_model({**model, "constraints": dict(model["constraints"]),
                             "variables": {k: dict(v) for k, v in model["variables"]}})
    if not any(str(k).isdigit() for k, _ in model["constraints"]) and not any(
            str(k).isdigit() for k, _ in model["variables"]):
        assert_same(as_dict, base)

    # :135-191: integer / binary forms
    keys = [k for k, _ in model["variables"]]
    assert_same(tableau_model({**model, "integers": False}), tableau_model({**model, "integers": []}))
    assert_same(tableau_model({**model, "integers": True}), tableau_model({**model, "integers": set(keys)}))
    assert_same(tableau_model({**model, "binaries": True}), tableau_model({**model, "binaries": list(keys)}))
    key = keys[int(rand() * len(keys))]
    assert_same(tableau_model({**model, "integers": [key], "binaries": [key]}),
                tableau_model({**model, "integers": [], "binaries": [key]}))  # :184-191 binary wins

    # :223-242: equal has precedence over min/max
    cons = list(model["constraints"])
    eq = [i for i, (_, c) in enumerate(cons) if c.get("equal") is not None]
    if eq:
        i = eq[int(rand() * len(eq))]
        k, c = cons[i]
        mod = list(cons)
        mod[i] = (k, {"equal": c["equal"], "min": c["equal"] + 1.0, "max": c["equal"] - 1.0})
        assert_same(tableau_model({**model, "constraints": mod}), base)

    # :244-265: constraints with the same key merge
    i = int(rand() * len(cons))
    k, c = cons[i]
    other_c = {"max": rand() * 100.0 + (c.get("max") or 0.0), "min": rand() * 100.0 + (c.get("min") or 0.0)}
    hi = c.get("equal") if c.get("equal") is not None else (c.get("max") if c.get("max") is not None else float("inf"))
    lo = c.get("equal") if c.get("equal") is not None else (c.get("min") if c.get("min") is not None else -float("inf"))
    merged = list(cons)
    merged[i] = (k, {"max": min(hi, other_c["max"]), "min": max(lo, other_c["min"])})
    assert_same(tableau_model({**model, "constraints": cons + [(k, other_c)]}),
                tableau_model({**model, "constraints": merged}))

    # :282-300: the last coefficient with the same key wins
    variables = list(model["variables"])
    vi = int(rand() * len(variables))
    vk, coefs = variables[vi]
    coefs = list(coefs)
    if coefs:
        ci = int(rand() * len(coefs))
        ck, cv = coefs[ci]
        dup = list(coefs)
        dup[ci] = (ck, cv + rand() * 100.0)
        dup.append((ck, cv))
        nv = list(variables)
        nv[vi] = (vk, dup)
        assert same_bits(tableau_model({**model, "variables": nv}).tableau.matrix, base.tableau.matrix)

    # :308-333: removing a constraint removes its rows
    i = int(rand() * len(cons))
    removed = tableau_model({**model, "constraints": cons[:i] + cons[i + 1:]})
    if sum(1 for k2, _ in cons if k2 == cons[i][0]) == 1:
        nrows = lambda c: 2 if c.get("equal") is not None else (c.get("max") is not None) + (c.get("min") is not None)
        row = 1 + sum(nrows(c2) for _, c2 in cons[:i])
        k_rows = nrows(cons[i][1])
        exp = np.delete(base.tableau.matrix.reshape(base.tableau.height, W), range(row, row + k_rows), axis=0)
        assert same_bits(removed.tableau.matrix, exp.reshape(-1))
        assert removed.tableau.height == base.tableau.height - k_rows


def test_integer_like_keys_follow_js_property_order():
    """Object.entries puts canonical array-index keys first in ascending order (restatement rule 7)."""
    model = {"objective": "o", "constraints": {"c": {"max": 10}},
             "variables": {"b": {"c": 1, "o": 1}, "10": {"c": 2, "o": 2}, "2": {"c": 3, "o": 3}}}
    tm = tableau_model(model)
    assert [k for k, _ in tm.variables] == ["2", "10", "b"]
    assert_same(tm, M.tableau_model(model))


@pytest.mark.parametrize("case", ALL, ids=[c["name"] for c in ALL])
def test_sparse_form_is_the_dense_image(case):
    """The (cells, values) form handed to yalps_solve_sparse is the ordered list of the stores src/tableau.ts:100-134
    makes into its zero-filled matrix: replaying it gives the dense image bit for bit (signed zeros included)."""
    dense = tableau_model(case["model"])
    sparse = tableau_model(case["model"], 0)
    t = sparse.tableau
    assert t.matrix is None and t.cells.dtype == np.int32 and t.values.dtype == np.float64
    assert t.cells.size == t.values.size and (t.cells.size == 0 or (0 <= t.cells.min() and t.cells.max() < t.width * t.height))
    replay = np.zeros(t.width * t.height)
    for c, v in zip(t.cells.tolist(), t.values.tolist()):  # in order, like update() calls
        replay[c] = v
    assert same_bits(replay, dense.tableau.matrix)
    assert same_bits(t.dense(), dense.tableau.matrix) and t.matrix is not None
    assert_same(sparse, M.tableau_model(case["model"]))


def test_sparse_form_keeps_the_order_of_duplicate_coefficients():
    """src/tableau.ts:101: a constraint key listed twice by one variable stores twice, the last store wins."""
    model = {"direction": "minimize", "objective": "cost",
             "constraints": [("a", {"max": 4}), ("b", {"min": 1, "max": 3})],
             "variables": [("x", [("a", 1), ("cost", 2), ("a", 5), ("b", -0.0)]), ("y", [("b", 2), ("b", 7), ("cost", 0)])]}
    sparse = tableau_model(model, 0)
    t = sparse.tableau
    assert len(set(t.cells.tolist())) < t.cells.size  # duplicates are kept, not resolved, in the sparse form
    assert same_bits(tableau_model(model).tableau.matrix, t.dense())
    assert_same(sparse, M.tableau_model(model))
    m = t.dense().reshape(t.height, t.width)
    assert m[1, 1] == 5.0 and m[2, 2] == 7.0 and m[3, 2] == -7.0 and np.signbit(m[0, 2]) and np.signbit(m[2, 1])


def _both_builders(model):
    """tableau_model through the native core (csrc/tabfast.c) and through the Python loops it restates."""
    from yalps_b200 import tableau as T
    assert T._NATIVE is not None, "yalps_b200/_tabfast*.so is missing: run `make -C yalps_b200/csrc`"
    native = T.tableau_model(model, 0)
    saved, T._NATIVE = T._NATIVE, None
    try:
        python = T.tableau_model(model, 0)
    finally:
        T._NATIVE = saved
    a, b = native.tableau, python.tableau
    assert (a.width, a.height) == (b.width, b.height) and native.integers == python.integers
    assert np.array_equal(a.cells, b.cells) and same_bits(a.values, b.values)
    return native


@pytest.mark.parametrize("case", ALL, ids=[c["name"] for c in ALL])
def test_native_builder_makes_the_same_stores_as_the_python_loops(case):
    _both_builders(case["model"])


def test_native_builder_on_the_corners_of_the_model_format():
    class Con:  # any object with min / max / equal attributes is a constraint (src/types.ts:7-31)
        def __init__(self, **kw):
            self.__dict__.update(kw)

    models = [
        {"variables": {}, "constraints": {}},
        # integer-like keys are iterated first, ascending (JS property order), in variables, constraints and coefficients
        {"objective": "10", "direction": "minimize",
         "constraints": {"b": {"max": 5}, "10": {"min": 1}, "2": {"equal": 3}, "02": {"max": 9}},
         "variables": {"y": {"b": 1, "10": 2, "2": 3, "02": 4}, "7": {"2": "1.5", "b": -0.0}, "x": [("b", 2), ["2", 1]]},
         "integers": ["7"], "binaries": {"x"}},
        # objects, tuples, generators, keys without a finite bound, NaN bounds, merged duplicates, objective == constraint
        {"objective": "c", "constraints": [("c", Con(max=10)), ("free", Con()), ("c", {"min": -2, "max": 12}),
                                           ("nan", {"max": float("nan")}), ("e", Con(equal=4, min=0)), (3, {"min": 1}),
                                           (3.0, {"max": 8})],
         "variables": [("u", (p for p in [("c", 1), ("free", "7"), ("nan", 5), (3, 2), ("e", True)])),
                       ("v", {"c": 2, "e": -1, 3: 0, "unknown": 1}), ("w", ())],
         "binaries": True},
        # coefficients of keys that have no row are never looked at (src/tableau.ts:104: `bounds != null` comes first)
        {"constraints": {"a": {"max": 1}, "free": {}}, "variables": [("x", {"a": 1, "free": "junk", "other": object()})]},
    ]
    for model in models:
        if isinstance(model["variables"], list):  # generators are consumed once: rebuild them per builder
            def fresh():
                return {**model, "variables": [(k, list(c) if not isinstance(c, (dict, tuple)) else c)
                                               for k, c in model["variables"]]}
            model = fresh()
        tm = _both_builders(model)
        if "free" not in model["constraints"] or isinstance(model["constraints"], list):
            assert same_bits(tm.tableau.dense(), M.tableau_model(model).tableau.matrix)
    for bad in ({"variables": [("x", [("a", "abc")])], "constraints": {"a": {"max": 1}}},
                {"variables": [("x", [("a", 1, 2)])], "constraints": {"a": {"max": 1}}},
                {"variables": [("x", [([], 1)])], "constraints": {"a": {"max": 1}}}):
        with pytest.raises((ValueError, TypeError)):
            tableau_model(bad)
