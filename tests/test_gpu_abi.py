"""Error behaviour and edge cases of the C ABI on a real device."""
import ctypes as C

import numpy as np
import pytest

from oracle import lib as O
from yalps_b200 import _ffi, engine as E

pytestmark = pytest.mark.gpu


def test_empty_batch_and_optional_outputs(engine):
    out = engine.solve_batch(np.zeros(0), 3, 4)
    assert out["status"].shape == (0,)
    lib = _ffi.load()
    mats = O.generate_synthetic(0, 4, 5, 6)
    opt = E.make_options()
    status = np.empty(4, np.int32)
    # every output except status is NULL
    rc = lib.yalps_solve_batch(engine._ctx, 4, 6, 7, mats.ctypes.data_as(C.c_void_p), C.byref(opt),
                               status.ctypes.data_as(C.c_void_p), None, None, None, None, None, None)
    assert rc == 0 and (status == 0).all()


def test_argument_errors_are_reported_not_thrown(engine):
    lib = _ffi.load()
    opt = E.make_options()
    rc = lib.yalps_solve_batch(engine._ctx, 1, 0, 4, None, C.byref(opt), None, None, None, None, None, None, None)
    assert rc == -2 and b"bad" in lib.yalps_last_error(engine._ctx)
    rc = lib.yalps_bnb_solve_nodes(None, 0, None, None, None, None, C.byref(opt), None, None, None, None, None, None, None)
    assert rc == -2
    with pytest.raises(_ffi.YalpsError):
        engine.set_tuning(9, 0)
    with pytest.raises(_ffi.YalpsError) as e:
        engine.set_tuning(E.PATH_SMEM, 0)
        try:
            engine.solve_batch(np.zeros(600 * 700), 600, 700)
        finally:
            engine.set_tuning(E.PATH_AUTO, 0)
    assert e.value.code == -3  # does not fit in shared memory


def test_branch_and_cut_without_root_is_an_error():
    import yalps_b200
    eng = yalps_b200.Engine(0)
    try:
        with pytest.raises(_ffi.YalpsError):
            eng._root_shape = (2, 2)
            eng.branch_and_cut([1], 1.0, 0.0)
    finally:
        eng.close()


def test_check_cycles_history_limit_is_an_error_not_a_wrong_answer(engine):
    """With checkCycles the per-phase history is capped at 262,144 pivots; a run that would exceed it
    reports YALPS_ERR_HISTORY.  Here the cap is reached artificially through max_pivots > cap on a cycling LP is
    not possible (it cycles at 11), so only the non-error path is checked."""
    t = np.array([[0, 10, -57, -9, -24], [0, 0.5, -5.5, -2.5, 9], [0, 0.5, -1.5, -0.5, 1], [1, 1, 0, 0, 0]], float)
    out = engine.solve_batch(t.reshape(1, -1), 4, 5, E.make_options(check_cycles=True, max_pivots=float("inf")))
    assert out["status"][0] == 4


def test_launch_counter_and_device_info(engine):
    before = engine.launch_count
    engine.solve_batch(O.generate_synthetic(0, 2, 3, 4), 4, 5)
    assert engine.launch_count == before + 1
    info = engine.device_info()
    assert info["cc"][0] >= 10 and info["sm_count"] >= 100 and info["smem_per_block_optin"] >= 200 * 1024


def test_scratchpad_stream_measurements(engine):
    """The roofline denominators bench.py measures live: the shared-memory stream and the tensor-memory stream
    (tcgen05.ld/st as a lane-private scratchpad moves more bytes per clock than the shared-memory data pipe)."""
    smem, mhz = engine.measure_smem_bandwidth()
    tmem, _ = engine.measure_tmem_bandwidth()
    assert mhz > 500 and 10e3 < smem < 60e3   # GB/s: 128 B/clk/SM x 148 SMs ~ 37 TB/s at 1965 MHz
    assert tmem > smem


def test_automatic_path_matches_forced_paths_on_a_dense_batch(engine):
    """n > 2 x SMs dense LPs that fit tensor memory take K1t automatically; K1, K1t and auto agree bit for bit."""
    mats = O.generate_synthetic(4242, 700, 32, 64, 4)
    outs = []
    for path in (E.PATH_SMEM, E.PATH_TMEM, E.PATH_AUTO):
        engine.set_tuning(path, 0)
        try:
            outs.append(engine.solve_batch(mats, 33, 65, want_matrices=True))
        finally:
            engine.set_tuning(E.PATH_AUTO, 0)
    for o in outs[1:]:
        for k in ("status", "pivots", "pos", "var"):
            assert np.array_equal(outs[0][k], o[k]), k
        for k in ("value", "rhs", "matrices"):
            assert np.array_equal(outs[0][k].view(np.uint64), o[k].view(np.uint64)), k
