"""yalps_b200: B200-native batched simplex engine behind the YALPS solve() API.

Public surface mirrors src/index.ts:1-3 of the reference: `solve`, `default_options` and the constraint
helpers, plus the new `solve_many`.  Everything numeric runs in libyalps_b200.so (hand-written sm_100a
kernels, C ABI in include/yalps_b200.h); importing this package without the built library, or calling
it without a CUDA device, raises.
"""
from .constraint import equal_to, equalTo, greater_eq, greaterEq, in_range, inRange, less_eq, lessEq
from .engine import Engine, MultiEngine, STATUS_NAMES, make_options
from .solver import default_options, defaultOptions, get_engine, solve, solve_many, solveMany
from .mps import apply_bounds, model_from_mps, netlib_model
from .tableau import Tableau, TableauModel, tableau_model

__all__ = [
    "solve", "solve_many", "solveMany", "default_options", "defaultOptions", "less_eq", "greater_eq", "equal_to",
    "in_range", "lessEq", "greaterEq", "equalTo", "inRange", "Engine", "MultiEngine", "make_options", "STATUS_NAMES",
    "tableau_model", "Tableau", "TableauModel", "get_engine", "model_from_mps", "netlib_model", "apply_bounds",
]
