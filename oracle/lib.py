"""oracle/lib.py -- TEST INFRASTRUCTURE.  ctypes face of liboracle.so (oracle/yalps_oracle.c)."""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "yalps_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.oracle_round_to_precision.restype = C.c_double
        _lib.oracle_round_to_precision.argtypes = [C.c_double, C.c_double]
        _lib.oracle_simplex.restype = C.c_int
        _lib.oracle_simplex.argtypes = [_dp, C.c_int32, C.c_int32, _ip, _ip, C.c_double, C.c_double, C.c_int32, _dp, _lp]
        _lib.oracle_simplex_batch.restype = C.c_int
        _lib.oracle_simplex_batch.argtypes = [C.c_int64, _dp, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32,
                                              _ip, _dp, _lp, _dp, _ip, _ip, C.c_int32]
        _lib.oracle_branch_and_cut.restype = C.c_int
        _lib.oracle_branch_and_cut.argtypes = [_dp, C.c_int32, C.c_int32, _ip, _ip, _ip, C.c_int32, C.c_double,
                                               C.c_double, C.c_double, C.c_double, C.c_int32, C.c_double, C.c_double,
                                               C.c_double, _dp, _ip, _dp, _ip, _ip, _lp, _dp, C.c_int64]
        _lib.oracle_apply_cuts.restype = None
        _lib.oracle_apply_cuts.argtypes = [_dp, C.c_int32, C.c_int32, _ip, _ip, _dp, _ip, _dp, C.c_int32, _dp, _ip, _ip]
        _lib.oracle_generate_synthetic.restype = None
        _lib.oracle_generate_synthetic.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, _dp]
        _lib.oracle_prospector_hash.restype = C.c_uint32
        _lib.oracle_prospector_hash.argtypes = [C.c_uint32]
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _l(a):
    return a.ctypes.data_as(_lp)


def round_to_precision(x: float, precision: float) -> float:
    return load().oracle_round_to_precision(x, precision)


def simplex(matrix, width, height, pos, var, precision=1e-8, max_pivots=8192, check_cycles=False):
    """In place on matrix/pos/var (numpy float64 / int32).  Returns (status:int, value:float, (p1, p2))."""
    assert matrix.dtype == np.float64 and matrix.flags.c_contiguous and matrix.size == width * height
    assert pos.dtype == np.int32 and var.dtype == np.int32 and pos.size == width + height
    res = C.c_double()
    piv = np.zeros(2, dtype=np.int64)
    st = load().oracle_simplex(_d(matrix), width, height, _i(pos), _i(var), precision, float(max_pivots),
                               int(bool(check_cycles)), C.byref(res), _l(piv))
    return st, res.value, (int(piv[0]), int(piv[1]))


def simplex_batch(matrices, width, height, precision=1e-8, max_pivots=8192, check_cycles=False, nthreads=1,
                  want_pos=True):
    """matrices: (n, height*width) float64, solved in place.  Returns dict of arrays."""
    n = matrices.shape[0]
    assert matrices.dtype == np.float64 and matrices.flags.c_contiguous
    status = np.zeros(n, dtype=np.int32)
    value = np.zeros(n, dtype=np.float64)
    pivots = np.zeros((n, 2), dtype=np.int64)
    rhs = np.zeros((n, height), dtype=np.float64)
    pos = np.zeros((n, width + height), dtype=np.int32) if want_pos else None
    var = np.zeros((n, width + height), dtype=np.int32) if want_pos else None
    load().oracle_simplex_batch(n, _d(matrices), width, height, precision, float(max_pivots), int(bool(check_cycles)),
                                _i(status), _d(value), _l(pivots), _d(rhs), _i(pos) if want_pos else None,
                                _i(var) if want_pos else None, int(nthreads))
    return {"status": status, "value": value, "pivots": pivots, "rhs": rhs, "pos": pos, "var": var}


def branch_and_cut(root_m, width, height, root_pos, root_var, ints, sign, init_result, precision=1e-8,
                   max_pivots=8192, check_cycles=False, tolerance=0.0, timeout=math.inf, max_iterations=32768,
                   node_log_cap=1 << 16):
    nints = int(ints.size)
    cap_rows = height + 2 * nints
    res = C.c_double()
    out_h = C.c_int32()
    rhs = np.zeros(cap_rows, dtype=np.float64)
    pos = np.zeros(width + cap_rows, dtype=np.int32)
    var = np.zeros(width + cap_rows, dtype=np.int32)
    stats = np.zeros(4, dtype=np.int64)
    log = np.zeros((node_log_cap, 4), dtype=np.float64)
    st = load().oracle_branch_and_cut(_d(root_m), width, height, _i(root_pos), _i(root_var), _i(ints), nints,
                                      float(sign), float(init_result), precision, float(max_pivots),
                                      int(bool(check_cycles)), float(tolerance), float(timeout),
                                      float(max_iterations), C.byref(res), C.byref(out_h), _d(rhs), _i(pos), _i(var),
                                      _l(stats), _d(log), node_log_cap)
    h = out_h.value
    return st, res.value, rhs[:h].copy(), pos[:width + h].copy(), var[:width + h].copy(), stats, log[:min(int(stats[0]), node_log_cap)].copy()


def apply_cuts(root_m, width, height, root_pos, root_var, cut_sign, cut_var, cut_value):
    k = int(len(cut_var))
    out_m = np.zeros((height + k) * width, dtype=np.float64)
    out_pos = np.zeros(width + height + k, dtype=np.int32)
    out_var = np.zeros(width + height + k, dtype=np.int32)
    cs = np.ascontiguousarray(cut_sign, dtype=np.float64)
    cv = np.ascontiguousarray(cut_var, dtype=np.int32)
    cx = np.ascontiguousarray(cut_value, dtype=np.float64)
    load().oracle_apply_cuts(_d(root_m), width, height, _i(root_pos), _i(root_var), _d(cs), _i(cv), _d(cx), k,
                             _d(out_m), _i(out_pos), _i(out_var))
    return out_m, out_pos, out_var


def generate_synthetic(first: int, n: int, m: int, nvars: int, neg_rows: int = 0, salt: int = 0x5BD1E995):
    """(n, (m+1)*(nvars+1)) tableaus of SURVEY 8(d) config 2 / 5."""
    out = np.zeros((n, (m + 1) * (nvars + 1)), dtype=np.float64)
    load().oracle_generate_synthetic(first, n, m, nvars, neg_rows, salt, _d(out))
    return out
