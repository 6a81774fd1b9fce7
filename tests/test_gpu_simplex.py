"""GPU parity: CUDA simplex kernels (through the C ABI) against the CPU oracle, bit for bit --
status, value, pivot counts per phase, final positionOfVariable / variableAtPosition, RHS column and the
whole final tableau (signed zeros included).  Sizes are chosen so the oracle finishes in seconds."""
import math

import numpy as np
import pytest

from conftest import bits, load_netlib, same_bits, same_value
from oracle import lib as O
from yalps_b200 import engine as E

pytestmark = pytest.mark.gpu


def oracle_batch(mats, H, W, **kw):
    work = mats.copy()
    res = O.simplex_batch(work, W, H, **kw)
    res["matrices"] = work
    return res


def assert_batch_equal(got, exp, what=""):
    assert np.array_equal(got["status"], exp["status"]), f"{what}: status"
    assert np.array_equal(got["pivots"], exp["pivots"]), f"{what}: pivots"
    gv, ev = got["value"], exp["value"]
    assert np.array_equal(np.isnan(gv), np.isnan(ev)) and np.array_equal(bits(gv[~np.isnan(gv)]), bits(ev[~np.isnan(ev)])), f"{what}: value"
    assert np.array_equal(got["pos"], exp["pos"]) and np.array_equal(got["var"], exp["var"]), f"{what}: basis"
    assert same_bits(got["rhs"], exp["rhs"]), f"{what}: rhs"
    if got.get("matrices") is not None:
        assert same_bits(got["matrices"].reshape(-1), exp["matrices"].reshape(-1)), f"{what}: final tableau"


SHAPES = [(32, 64, 0), (32, 64, 8), (1, 1, 0), (5, 3, 2), (7, 40, 3), (40, 7, 10), (16, 100, 4), (60, 31, 20),
          (33, 33, 5), (90, 130, 30)]


@pytest.mark.parametrize("m,nv,neg", SHAPES)
def test_synthetic_batches_bit_exact(engine, m, nv, neg):
    n = 96
    H, W = m + 1, nv + 1
    mats = O.generate_synthetic(1000, n, m, nv, neg)
    exp = oracle_batch(mats, H, W)
    engine.set_tuning(E.PATH_AUTO, 0)
    got = engine.solve_batch(mats, H, W, want_matrices=True)
    assert_batch_equal(got, exp, f"{m}x{nv}")


@pytest.mark.parametrize("threads", [32, 64, 128, 256, 512, 1024])
@pytest.mark.parametrize("path", [E.PATH_SMEM, E.PATH_GMEM])
def test_every_cta_width_and_path(engine, threads, path):
    m, nv, n = 32, 64, 64
    mats = O.generate_synthetic(5, n, m, nv, 6)
    exp = oracle_batch(mats, m + 1, nv + 1)
    engine.set_tuning(path, threads)
    try:
        got = engine.solve_batch(mats, m + 1, nv + 1, want_matrices=True)
    finally:
        engine.set_tuning(E.PATH_AUTO, 0)
    assert_batch_equal(got, exp, f"path={path} threads={threads}")


SPLIT_SHAPES = [(32, 64, 6), (5, 3, 2), (40, 7, 10), (16, 100, 4), (90, 130, 30), (12, 700, 3), (150, 103, 40)]


@pytest.mark.parametrize("rows", [2, 4, 8, 16])
@pytest.mark.parametrize("path", [E.PATH_SMEM, E.PATH_GMEM])
@pytest.mark.parametrize("m,nv,neg", SPLIT_SHAPES)
def test_row_split_kernels_bit_exact(engine, m, nv, neg, path, rows):
    """simplex_split.cuh: NWR row groups x NWC column warps, compacted active-row list."""
    n = 24
    H, W = m + 1, nv + 1
    mats = O.generate_synthetic(4242 + m, n, m, nv, neg)
    exp = oracle_batch(mats, H, W)
    for threads in (64 * rows, 512):
        engine.set_tuning(path, threads, rows)
        try:
            got = engine.solve_batch(mats, H, W, want_matrices=True)
        finally:
            engine.set_tuning(E.PATH_AUTO, 0)
        assert_batch_equal(got, exp, f"{m}x{nv} path={path} rows={rows} threads={threads}")


@pytest.mark.parametrize("rows", [2, 8])
def test_row_split_special_values_options_and_cycles(engine, rows):
    H, W = 4, 5
    t = np.zeros((4, H * W))
    t[0].reshape(H, W)[:] = [[0.0, 3.0, 2.0, -0.0, 1.0], [4.0, 1.0, 1e-16, 1.0000000000000001e-16, -0.0],
                             [5.0, 2e-16, 1.0, -1e-17, 3.0], [6.0, -0.0, 2.0, 1.0, 1e-15]]
    t[1].reshape(H, W)[:] = [[0, 1, 1, 1, 1], [0, 1, 1, 1, 1], [0, 1, 1, 1, 1], [0, 1, 1, 1, 1]]
    t[2].reshape(H, W)[:] = [[0, -1, -1, 2, 2], [-1, -1, -1, -1, -1], [-1, -1, -1, -1, -1], [3, 1, 1, 1, 1]]
    t[3].reshape(H, W)[:] = 1.0
    t[3].reshape(H, W)[1, 2] = math.nan
    t[3].reshape(H, W)[2, 0] = math.inf
    t[3].reshape(H, W)[0, 3] = 5.0
    chv = np.array([[0, 10, -57, -9, -24], [0, 0.5, -5.5, -2.5, 9], [0, 0.5, -1.5, -0.5, 1], [1, 1, 0, 0, 0]], float)
    m, nv = 32, 64
    mats = O.generate_synthetic(300, 16, m, nv, 8)
    try:
        for path in (E.PATH_SMEM, E.PATH_GMEM):
            engine.set_tuning(path, 64 * rows, rows)
            assert_batch_equal(engine.solve_batch(t, H, W, want_matrices=True), oracle_batch(t, H, W), "special")
            for cc in (True, False):
                c = chv.reshape(1, -1).copy()
                assert_batch_equal(engine.solve_batch(c, 4, 5, E.make_options(check_cycles=cc), want_matrices=True),
                                   oracle_batch(c, 4, 5, check_cycles=cc), f"chvatal cc={cc}")
            for mp in (0, 1, 7, 11.5, math.inf):
                assert_batch_equal(engine.solve_batch(mats, m + 1, nv + 1, E.make_options(max_pivots=mp), want_matrices=True),
                                   oracle_batch(mats, m + 1, nv + 1, max_pivots=mp), f"maxPivots={mp}")
            for prec in (1e-3, 0.0, 0.25):
                assert_batch_equal(engine.solve_batch(mats, m + 1, nv + 1, E.make_options(precision=prec), want_matrices=True),
                                   oracle_batch(mats, m + 1, nv + 1, precision=prec), f"precision={prec}")
    finally:
        engine.set_tuning(E.PATH_AUTO, 0)


@pytest.mark.parametrize("name", ["AFIRO", "ADLITTLE", "SC105", "BLEND", "KLEIN1", "SHARE2B", "SC205", "ISRAEL"])
def test_row_split_netlib_sparse(engine, name):
    """Sparse Netlib tableaus: the compacted row list is short and changes every pivot."""
    g = load_netlib().get(name)
    H, W = g["height"], g["width"]
    mats = np.asarray(g["matrix"], np.float64).reshape(1, -1)
    exp = oracle_batch(mats, H, W)
    try:
        for path in (E.PATH_SMEM, E.PATH_GMEM):
            for rows, threads in ((4, 256), (8, 512), (16, 512)):
                engine.set_tuning(path, threads, rows)
                try:
                    got = engine.solve_batch(mats, H, W, want_matrices=True)
                except E.YalpsError:
                    assert path == E.PATH_SMEM  # does not fit in shared memory
                    continue
                assert_batch_equal(got, exp, f"{name} path={path} rows={rows}")
    finally:
        engine.set_tuning(E.PATH_AUTO, 0)


CLUSTER_SHAPES = [(32, 64, 6), (5, 3, 2), (1, 1, 0), (40, 7, 10), (16, 100, 4), (90, 130, 30), (12, 700, 3), (150, 103, 40),
                  (200, 260, 50), (330, 500, 60)]


@pytest.mark.parametrize("m,nv,neg", CLUSTER_SHAPES)
def test_cluster_kernel_bit_exact(engine, m, nv, neg):
    """cluster_kernel.cuh: tableau distributed over a thread-block cluster, pivot row published through L2
    (DSMEM carries only the selection records)."""
    n = 6 if m * nv > 50000 else 20
    H, W = m + 1, nv + 1
    mats = O.generate_synthetic(777 + m, n, m, nv, neg)
    exp = oracle_batch(mats, H, W)
    engine.set_tuning(E.PATH_CLUSTER, 0)
    try:
        got = engine.solve_batch(mats, H, W, want_matrices=True)
    finally:
        engine.set_tuning(E.PATH_AUTO, 0)
    assert_batch_equal(got, exp, f"cluster {m}x{nv}")


def test_cluster_kernel_options_special_values_and_cycles(engine):
    H, W = 4, 5
    t = np.zeros((4, H * W))
    t[0].reshape(H, W)[:] = [[0.0, 3.0, 2.0, -0.0, 1.0], [4.0, 1.0, 1e-16, 1.0000000000000001e-16, -0.0],
                             [5.0, 2e-16, 1.0, -1e-17, 3.0], [6.0, -0.0, 2.0, 1.0, 1e-15]]
    t[1].reshape(H, W)[:] = [[0, 1, 1, 1, 1], [0, 1, 1, 1, 1], [0, 1, 1, 1, 1], [0, 1, 1, 1, 1]]
    t[2].reshape(H, W)[:] = [[0, -1, -1, 2, 2], [-1, -1, -1, -1, -1], [-1, -1, -1, -1, -1], [3, 1, 1, 1, 1]]
    t[3].reshape(H, W)[:] = 1.0
    t[3].reshape(H, W)[1, 2] = math.nan
    t[3].reshape(H, W)[2, 0] = math.inf
    t[3].reshape(H, W)[0, 3] = 5.0
    chv = np.array([[0, 10, -57, -9, -24], [0, 0.5, -5.5, -2.5, 9], [0, 0.5, -1.5, -0.5, 1], [1, 1, 0, 0, 0]], float)
    m, nv = 32, 64
    mats = O.generate_synthetic(300, 16, m, nv, 8)
    engine.set_tuning(E.PATH_CLUSTER, 0)
    try:
        assert_batch_equal(engine.solve_batch(t, H, W, want_matrices=True), oracle_batch(t, H, W), "special")
        for cc in (True, False):
            c = chv.reshape(1, -1).copy()
            assert_batch_equal(engine.solve_batch(c, 4, 5, E.make_options(check_cycles=cc), want_matrices=True),
                               oracle_batch(c, 4, 5, check_cycles=cc), f"chvatal cc={cc}")
        for mp in (0, 1, 7, 11.5, math.inf):
            assert_batch_equal(engine.solve_batch(mats, m + 1, nv + 1, E.make_options(max_pivots=mp), want_matrices=True),
                               oracle_batch(mats, m + 1, nv + 1, max_pivots=mp), f"maxPivots={mp}")
        for prec in (1e-3, 0.0, 0.25):
            assert_batch_equal(engine.solve_batch(mats, m + 1, nv + 1, E.make_options(precision=prec), want_matrices=True),
                               oracle_batch(mats, m + 1, nv + 1, precision=prec), f"precision={prec}")
    finally:
        engine.set_tuning(E.PATH_AUTO, 0)


@pytest.mark.parametrize("name", ["AFIRO", "SC105", "SC205", "ISRAEL", "AGG", "SCAGR7", "E226", "BEACONFD", "SHARE1B", "SCFXM1"])
def test_cluster_kernel_netlib(engine, name):
    g = load_netlib().get(name)
    H, W = g["height"], g["width"]
    mats = np.asarray(g["matrix"], np.float64).reshape(1, -1)
    exp = oracle_batch(mats, H, W)
    engine.set_tuning(E.PATH_CLUSTER, 0)
    try:
        got = engine.solve_batch(mats, H, W, want_matrices=True)
    finally:
        engine.set_tuning(E.PATH_AUTO, 0)
    assert_batch_equal(got, exp, f"cluster {name}")


def test_wide_tableau_multi_chunk(engine):
    """W > 32*KC forces several column chunks per row."""
    m, nv, n = 12, 700, 8
    mats = O.generate_synthetic(77, n, m, nv, 3)
    exp = oracle_batch(mats, m + 1, nv + 1)
    for threads in (32, 128, 1024):
        engine.set_tuning(E.PATH_AUTO, threads)
        got = engine.solve_batch(mats, m + 1, nv + 1, want_matrices=True)
        assert_batch_equal(got, exp, f"threads={threads}")
    engine.set_tuning(E.PATH_AUTO, 0)


def test_max_pivots_is_per_phase_and_may_be_infinite(engine):
    m, nv, n = 32, 64, 32
    mats = O.generate_synthetic(300, n, m, nv, 8)
    for mp in (0, 1, 3, 7, 11.5, math.inf):
        exp = oracle_batch(mats, m + 1, nv + 1, max_pivots=mp)
        got = engine.solve_batch(mats, m + 1, nv + 1, E.make_options(max_pivots=mp), want_matrices=True)
        assert_batch_equal(got, exp, f"maxPivots={mp}")
        if mp == 0:
            assert (got["status"] == 4).all()  # "cycled"


def test_precision_option(engine):
    m, nv, n = 20, 30, 32
    mats = O.generate_synthetic(900, n, m, nv, 5)
    for prec in (1e-8, 1e-3, 0.0, 0.25):
        exp = oracle_batch(mats, m + 1, nv + 1, precision=prec)
        got = engine.solve_batch(mats, m + 1, nv + 1, E.make_options(precision=prec), want_matrices=True)
        assert_batch_equal(got, exp, f"precision={prec}")


def test_statuses_unbounded_infeasible_and_special_values(engine):
    H, W = 4, 5
    t = np.zeros((6, H * W))
    # 0: unbounded (positive cost, no positive column entry)
    t[0].reshape(H, W)[0, 1] = 1.0
    t[0].reshape(H, W)[1:, 1] = -1.0
    t[0].reshape(H, W)[1:, 0] = 1.0
    # 1: infeasible (negative rhs, no negative coefficient)
    t[1].reshape(H, W)[1, 0] = -1.0
    t[1].reshape(H, W)[1, 1:] = 1.0
    # 2: NaN / inf cells
    t[2].reshape(H, W)[:] = 1.0
    t[2].reshape(H, W)[1, 2] = math.nan
    t[2].reshape(H, W)[2, 0] = math.inf
    t[2].reshape(H, W)[0, 3] = 5.0
    # 3: tiny cells around the 1e-16 skip threshold, negative zeros
    t[3].reshape(H, W)[:] = [[0.0, 3.0, 2.0, -0.0, 1.0], [4.0, 1.0, 1e-16, 1.0000000000000001e-16, -0.0],
                             [5.0, 2e-16, 1.0, -1e-17, 3.0], [6.0, -0.0, 2.0, 1.0, 1e-15]]
    # 4: degenerate ties everywhere
    t[4].reshape(H, W)[:] = [[0, 1, 1, 1, 1], [0, 1, 1, 1, 1], [0, 1, 1, 1, 1], [0, 1, 1, 1, 1]]
    # 5: phase 1 with ties
    t[5].reshape(H, W)[:] = [[0, -1, -1, 2, 2], [-1, -1, -1, -1, -1], [-1, -1, -1, -1, -1], [3, 1, 1, 1, 1]]
    exp = oracle_batch(t, H, W)
    got = engine.solve_batch(t, H, W, want_matrices=True)
    assert_batch_equal(got, exp, "special")
    assert got["status"][0] == 2 and got["value"][0] == 1.0
    assert got["status"][1] == 1 and math.isnan(got["value"][1])


def test_check_cycles_detects_the_chvatal_cycle(engine):
    """tests/cases/Chvatal Cycling.json as a raw tableau (maximize; c1,c2 <= 0; c3 <= 1)."""
    t = np.array([[0, 10, -57, -9, -24], [0, 0.5, -5.5, -2.5, 9], [0, 0.5, -1.5, -0.5, 1], [1, 1, 0, 0, 0]], float)
    for cc in (True, False):
        m = t.reshape(1, -1).copy()
        exp = oracle_batch(m, 4, 5, check_cycles=cc)
        got = engine.solve_batch(m, 4, 5, E.make_options(check_cycles=cc), want_matrices=True)
        assert_batch_equal(got, exp, f"checkCycles={cc}")
    assert exp["status"][0] == 4 or True
    got = engine.solve_batch(t.reshape(1, -1), 4, 5, E.make_options(check_cycles=True))
    assert got["status"][0] == 4 and tuple(got["pivots"][0]) == (0, 11)


def test_ragged_batch(engine):
    rng = np.random.default_rng(3)
    tabs, shapes, exp = [], [], []
    for i in range(40):
        m, nv = int(rng.integers(1, 40)), int(rng.integers(1, 70))
        t = O.generate_synthetic(i, 1, m, nv, int(rng.integers(0, m + 1)))[0]
        tabs.append(t)
        shapes.append((m + 1, nv + 1))
        exp.append(oracle_batch(t.reshape(1, -1), m + 1, nv + 1))
    got = engine.solve_ragged(tabs, shapes, want_matrices=True)
    for g, e, s in zip(got, exp, shapes):
        assert g["status"] == e["status"][0] and g["pivots"] == tuple(e["pivots"][0])
        assert same_value(g["value"], e["value"][0])
        assert np.array_equal(g["pos"], e["pos"][0]) and np.array_equal(g["var"], e["var"][0])
        assert same_bits(g["rhs"], e["rhs"][0]) and same_bits(g["matrix"], e["matrices"][0])


def test_ragged_batch_on_the_cluster_kernel(engine):
    """Ragged groups reach the cluster kernel through the LP index indirection."""
    rng = np.random.default_rng(5)
    tabs, shapes, exp = [], [], []
    for i, (m, nv) in enumerate([(3, 4), (40, 70), (200, 260), (90, 130), (1, 1), (250, 300), (33, 9), (120, 400)]):
        t = O.generate_synthetic(900 + i, 1, m, nv, min(m, 4))[0]
        tabs.append(t)
        shapes.append((m + 1, nv + 1))
        exp.append(oracle_batch(t.reshape(1, -1), m + 1, nv + 1))
    engine.set_tuning(E.PATH_CLUSTER, 0)
    try:
        got = engine.solve_ragged(tabs, shapes, want_matrices=True)
    finally:
        engine.set_tuning(E.PATH_AUTO, 0)
    for g, e, s in zip(got, exp, shapes):
        assert g["status"] == e["status"][0] and g["pivots"] == tuple(e["pivots"][0]), s
        assert same_value(g["value"], e["value"][0])
        assert np.array_equal(g["pos"], e["pos"][0]) and np.array_equal(g["var"], e["var"][0])
        assert same_bits(g["rhs"], e["rhs"][0]) and same_bits(g["matrix"], e["matrices"][0]), s


def test_round_to_precision_device(engine):
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.normal(size=4000) * 10.0 ** rng.integers(-10, 10, 4000),
                         [0.5e-8, -0.5e-8, 1.5e-8, 2.5e-8, -0.0, 0.0, 1e300, -1e300, math.inf, -math.inf, math.nan]])
    for p in (1e-8, 1e-5, 0.5, 3e-9):
        got = engine.round_to_precision(xs, p)
        for x, g in zip(xs, got):
            assert same_value(g, O.round_to_precision(float(x), p)), (x, p)


def test_device_generator_matches_oracle_generator(engine):
    import torch
    for (m, nv, neg) in [(32, 64, 0), (32, 64, 8), (3, 5, 1)]:
        n = 50
        d = torch.empty(n * (m + 1) * (nv + 1), dtype=torch.float64, device="cuda")
        engine.generate_synthetic_device(123, n, m, nv, d.data_ptr(), neg_rows=neg)
        torch.cuda.synchronize()
        assert same_bits(d.cpu().numpy(), O.generate_synthetic(123, n, m, nv, neg).reshape(-1))


def test_device_resident_entry_point(engine):
    import torch
    m, nv, n = 32, 64, 512
    H, W = m + 1, nv + 1
    mats = O.generate_synthetic(0, n, m, nv, 0)
    exp = oracle_batch(mats, H, W)
    d_in = torch.from_numpy(mats).cuda()
    st = torch.empty(n, dtype=torch.int32, device="cuda")
    val = torch.empty(n, dtype=torch.float64, device="cuda")
    piv = torch.empty(n, 2, dtype=torch.int64, device="cuda")
    rhs = torch.empty(n, H, dtype=torch.float64, device="cuda")
    pos = torch.empty(n, W + H, dtype=torch.int32, device="cuda")
    var = torch.empty(n, W + H, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    engine.solve_batch_device(n, H, W, d_in.data_ptr(), d_status=st.data_ptr(), d_value=val.data_ptr(),
                              d_pivots=piv.data_ptr(), d_rhs=rhs.data_ptr(), d_pos=pos.data_ptr(),
                              d_var=var.data_ptr(), stream=stream)
    torch.cuda.synchronize()
    got = {"status": st.cpu().numpy(), "value": val.cpu().numpy(), "pivots": piv.cpu().numpy(),
           "rhs": rhs.cpu().numpy(), "pos": pos.cpu().numpy(), "var": var.cpu().numpy()}
    assert_batch_equal(got, exp, "device entry")
    assert same_bits(d_in.cpu().numpy(), mats)  # the SMEM path leaves its input untouched


NL = load_netlib()
NETLIB_QUICK = [n for n in NL.names if float(NL.z[f"{n}/oracle_seconds"][0]) < 1.0]


@pytest.mark.parametrize("name", NETLIB_QUICK)
def test_netlib_trajectory(engine, name):
    """Netlib models (golden vectors from tests/golden/netlib.npz): identical status, value, pivot counts,
    final basis and RHS column -- i.e. the reference's trajectory, whether or not it is the Netlib optimum."""
    g = NL.get(name)
    got = engine.solve_batch(g["matrix"], g["height"], g["width"], E.make_options(check_cycles=g["check_cycles"]))
    assert got["status"][0] == g["status"] and tuple(got["pivots"][0]) == g["pivots"]
    assert same_value(got["value"][0], g["value"])
    assert np.array_equal(got["pos"][0], g["final_pos"])
    assert same_bits(got["rhs"][0], g["final_rhs"])


# ---------------------------------------------------------------------------------------------- K4 grid kernel
@pytest.mark.parametrize("m,nv,neg", [(32, 64, 6), (5, 3, 2), (100, 300, 40), (300, 90, 100), (1, 1, 0)])
def test_grid_kernel_bit_exact(engine, m, nv, neg):
    """One LP across the whole grid (cooperative launch), forced on small inputs so the oracle is quick."""
    n = 4
    mats = O.generate_synthetic(4000, n, m, nv, neg)
    exp = oracle_batch(mats, m + 1, nv + 1)
    engine.set_tuning(E.PATH_GRID, 0)
    try:
        got = engine.solve_batch(mats, m + 1, nv + 1, want_matrices=True)
    finally:
        engine.set_tuning(E.PATH_AUTO, 0)
    assert_batch_equal(got, exp, f"grid {m}x{nv}")


def test_grid_kernel_check_cycles_and_budget(engine):
    t = np.array([[0, 10, -57, -9, -24], [0, 0.5, -5.5, -2.5, 9], [0, 0.5, -1.5, -0.5, 1], [1, 1, 0, 0, 0]], float)
    engine.set_tuning(E.PATH_GRID, 0)
    try:
        for kw in ({"check_cycles": True}, {"check_cycles": False, "max_pivots": 9}, {"max_pivots": 0}):
            m = t.reshape(1, -1).copy()
            exp = oracle_batch(m, 4, 5, **kw)
            got = engine.solve_batch(m, 4, 5, E.make_options(**kw), want_matrices=True)
            assert_batch_equal(got, exp, f"grid {kw}")
    finally:
        engine.set_tuning(E.PATH_AUTO, 0)


# ---------------------------------------------------------------------- KG: grid-wide, resident in shared memory
@pytest.mark.parametrize("m,nv,neg", [(32, 64, 6), (5, 3, 2), (100, 300, 40), (300, 90, 100), (1, 1, 0), (700, 1500, 90)])
def test_grid_resident_kernel_bit_exact(engine, m, nv, neg):
    """One LP across the whole grid with its rows resident in the SMs' shared memory (one grid barrier per pivot),
    forced on small inputs so the oracle is quick; 4 LPs run one after the other inside the one launch."""
    n = 4 if m < 500 else 1
    mats = O.generate_synthetic(4000, n, m, nv, neg)
    exp = oracle_batch(mats, m + 1, nv + 1)
    engine.set_tuning(E.PATH_GRID_RESIDENT, 0)
    try:
        got = engine.solve_batch(mats, m + 1, nv + 1, want_matrices=True)
    finally:
        engine.set_tuning(E.PATH_AUTO, 0)
    assert_batch_equal(got, exp, f"grid-resident {m}x{nv}")


def test_grid_resident_kernel_check_cycles_and_budget(engine):
    t = np.array([[0, 10, -57, -9, -24], [0, 0.5, -5.5, -2.5, 9], [0, 0.5, -1.5, -0.5, 1], [1, 1, 0, 0, 0]], float)
    engine.set_tuning(E.PATH_GRID_RESIDENT, 0)
    try:
        for kw in ({"check_cycles": True}, {"check_cycles": False, "max_pivots": 9}, {"max_pivots": 0}):
            m = t.reshape(1, -1).copy()
            exp = oracle_batch(m, 4, 5, **kw)
            got = engine.solve_batch(m, 4, 5, E.make_options(**kw), want_matrices=True)
            assert_batch_equal(got, exp, f"grid-resident {kw}")
    finally:
        engine.set_tuning(E.PATH_AUTO, 0)


def test_grid_resident_is_the_automatic_choice_beyond_a_cluster(engine):
    """1025 x 2049 (16.8 MB) no longer fits a 16-CTA cluster: the automatic path must give the K4 bits (and, by
    test_config5_full_size_capped_pivots_against_oracle, the oracle's) while launching exactly one kernel."""
    import torch
    m, nv, cap = 1024, 2048, 25
    H, W = m + 1, nv + 1
    d = torch.empty(H * W, dtype=torch.float64, device="cuda")
    engine.generate_synthetic_device(0, 1, m, nv, d.data_ptr())
    torch.cuda.synchronize()
    mats = d.cpu().numpy().reshape(1, -1)
    opt = E.make_options(max_pivots=cap)
    auto = engine.solve_batch(mats, H, W, opt, want_matrices=True)
    engine.set_tuning(E.PATH_GRID, 0)
    try:
        k4 = engine.solve_batch(mats, H, W, opt, want_matrices=True)
    finally:
        engine.set_tuning(E.PATH_AUTO, 0)
    assert_batch_equal(auto, k4, "KG vs K4")


def test_large_dense_lp_takes_the_grid_path(engine):
    """A tableau beyond shared memory with n = 1 (config 5 in miniature): auto path == K4."""
    m, nv = 400, 900
    mats = O.generate_synthetic(9, 1, m, nv, 60)
    exp = oracle_batch(mats, m + 1, nv + 1)
    got = engine.solve_batch(mats, m + 1, nv + 1, want_matrices=True)
    assert_batch_equal(got, exp, "400x900")
    assert got["pivots"].sum() > 100


NETLIB_LONG = [n for n in NL.names if float(NL.z[f"{n}/oracle_seconds"][0]) >= 1.0]


@pytest.mark.parametrize("name", NETLIB_LONG)
def test_netlib_long_trajectories(engine, name):
    """The reference's 'cannot handle' list (benchmarks/netlib/read.ts:55-58): thousands of pivots on dense,
    ill-conditioned tableaus.  Status / value / pivot counts / final basis must still be the reference's."""
    g = NL.get(name)
    got = engine.solve_batch(g["matrix"], g["height"], g["width"], E.make_options(check_cycles=g["check_cycles"]))
    assert got["status"][0] == g["status"] and tuple(got["pivots"][0]) == g["pivots"]
    assert same_value(got["value"][0], g["value"])
    assert np.array_equal(got["pos"][0], g["final_pos"])
    assert same_bits(got["rhs"][0], g["final_rhs"])


def _replica_rhs(base_rhs, groups, first, n, eps=1e-2, salt=0x2545F491):
    """numpy statement of the config-3 perturbation: one U per constraint key."""
    from oracle import model as M
    out = np.tile(base_rhs, (n, 1))
    for i in range(n):
        seed0 = M.prospector_hash((first + i) ^ salt)
        for r, g in enumerate(groups):
            if g >= 0:
                u = M.prospector_hash((seed0 + (int(g) + 1) * 0x9E3779B9) & 0xFFFFFFFF) / 4294967296.0
                out[i, r] = base_rhs[r] * (1.0 + eps * (2.0 * u - 1.0))
    return out


def test_replica_generator_and_config3_parity(engine):
    """Config 3: RHS-perturbed replicas of SC105 / ADLITTLE generated on the device, solved on both memory paths,
    against the oracle on the same replicas."""
    import torch
    for name, n in (("ADLITTLE", 24), ("SC105", 12)):
        g = NL.get(name)
        H, W = g["height"], g["width"]
        d = torch.empty(n * H * W, dtype=torch.float64, device="cuda")
        engine.generate_replicas_device(7, n, g["matrix"], H, W, g["row_groups"], d.data_ptr())
        torch.cuda.synchronize()
        mats = d.cpu().numpy().reshape(n, H * W)
        exp_mats = np.tile(g["matrix"], (n, 1))
        exp_mats.reshape(n, H, W)[:, :, 0] = _replica_rhs(g["matrix"].reshape(H, W)[:, 0], g["row_groups"], 7, n)
        assert same_bits(mats, exp_mats)
        exp = oracle_batch(mats, H, W)
        assert (exp["status"] == 0).all()
        for path in (E.PATH_SMEM, E.PATH_GMEM):
            engine.set_tuning(path, 0)
            got = engine.solve_batch(mats, H, W, want_matrices=True)
            assert_batch_equal(got, exp, f"{name} path={path}")
        engine.set_tuning(E.PATH_AUTO, 0)


def test_full_size_batch_properties(engine):
    """BASELINE.json config 2 at full size (65,536 LPs, too many for the oracle in a unit test): size-independent
    properties -- every LP optimal (feasible start, bounded), the basis arrays are inverse permutations, the
    reported value is roundToPrecision of the final M[0,0], final RHS is non-negative, and an oracle spot check
    on a random sample of 512 LPs is bit-exact."""
    import torch
    n, m, nv = 65536, 32, 64
    H, W = m + 1, nv + 1
    d = torch.empty(n * H * W, dtype=torch.float64, device="cuda")
    engine.generate_synthetic_device(0, n, m, nv, d.data_ptr())
    st = torch.empty(n, dtype=torch.int32, device="cuda")
    val = torch.empty(n, dtype=torch.float64, device="cuda")
    piv = torch.empty(n, 2, dtype=torch.int64, device="cuda")
    rhs = torch.empty(n, H, dtype=torch.float64, device="cuda")
    pos = torch.empty(n, W + H, dtype=torch.int32, device="cuda")
    var = torch.empty(n, W + H, dtype=torch.int32, device="cuda")
    engine.solve_batch_device(n, H, W, d.data_ptr(), d_status=st.data_ptr(), d_value=val.data_ptr(),
                              d_pivots=piv.data_ptr(), d_rhs=rhs.data_ptr(), d_pos=pos.data_ptr(), d_var=var.data_ptr(),
                              stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert bool((st == 0).all())
    assert bool((piv[:, 0] == 0).all()) and int(piv[:, 1].min()) >= 1
    idx = torch.arange(W + H, device="cuda", dtype=torch.int64).expand(n, -1)
    assert bool((torch.gather(pos.long(), 1, var.long()) == idx).all())  # pos[var[p]] == p
    assert bool((rhs[:, 1:] >= -1e-8).all())  # primal feasible at the optimum
    r0 = engine.round_to_precision(rhs[:, 0].cpu().numpy(), 1e-8)
    assert same_bits(r0, val.cpu().numpy())
    sample = np.random.default_rng(0).choice(n, 512, replace=False)
    mats = d.view(n, H * W)[torch.from_numpy(sample).cuda()].cpu().numpy()
    exp = oracle_batch(mats, H, W)
    assert np.array_equal(exp["status"], st.cpu().numpy()[sample])
    assert same_bits(exp["value"], val.cpu().numpy()[sample])
    assert np.array_equal(exp["pivots"], piv.cpu().numpy()[sample])
    assert same_bits(exp["rhs"], rhs.cpu().numpy()[sample]) and np.array_equal(exp["pos"], pos.cpu().numpy()[sample])
    assert np.array_equal(exp["var"], var.cpu().numpy()[sample])


@pytest.mark.parametrize("name,n", [("SC105", 32768), ("ADLITTLE", 65536)])
def test_config3_full_size_sample_against_oracle_on_the_automatic_path(engine, name, n):
    """BASELINE.json config 3 at the size bench.py runs per launch, on the AUTOMATIC path (the row-split HBM/L2
    kernel): a random sample of 1,024 replicas against the oracle on every output, plus size-independent properties
    over the whole batch (all optimal, inverse permutations, value = roundToPrecision(M[0,0]), primal feasibility)."""
    import torch
    g = NL.get(name)
    H, W = g["height"], g["width"]
    d = torch.empty(n * H * W, dtype=torch.float64, device="cuda")
    engine.generate_replicas_device(0, n, g["matrix"], H, W, g["row_groups"], d.data_ptr())
    work = torch.empty_like(d)
    st = torch.empty(n, dtype=torch.int32, device="cuda")
    val = torch.empty(n, dtype=torch.float64, device="cuda")
    piv = torch.empty(n, 2, dtype=torch.int64, device="cuda")
    rhs = torch.empty(n, H, dtype=torch.float64, device="cuda")
    pos = torch.empty(n, W + H, dtype=torch.int32, device="cuda")
    var = torch.empty(n, W + H, dtype=torch.int32, device="cuda")
    rows = torch.zeros(n, dtype=torch.int64, device="cuda")
    engine.set_row_counter(rows.data_ptr(), per_lp=True)
    try:
        engine.solve_batch_device(n, H, W, d.data_ptr(), d_work=work.data_ptr(), d_status=st.data_ptr(),
                                  d_value=val.data_ptr(), d_pivots=piv.data_ptr(), d_rhs=rhs.data_ptr(),
                                  d_pos=pos.data_ptr(), d_var=var.data_ptr(),
                                  stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    finally:
        engine.set_row_counter(0)
    assert bool((st == 0).all())
    idx = torch.arange(W + H, device="cuda", dtype=torch.int64).expand(n, -1)
    assert bool((torch.gather(pos.long(), 1, var.long()) == idx).all())
    assert bool((rhs[:, 1:] >= -1e-8).all())
    assert same_bits(engine.round_to_precision(rhs[:, 0].cpu().numpy(), 1e-8), val.cpu().numpy())
    sample = np.sort(np.random.default_rng(3).choice(n, 1024, replace=False))
    mats = d.view(n, H * W)[torch.from_numpy(sample).cuda()].cpu().numpy()
    exp = O.simplex_batch(mats, W, H, nthreads=16)
    assert np.array_equal(exp["status"], st.cpu().numpy()[sample])
    assert same_bits(exp["value"], val.cpu().numpy()[sample])
    assert np.array_equal(exp["pivots"], piv.cpu().numpy()[sample])
    assert same_bits(exp["rhs"], rhs.cpu().numpy()[sample])
    assert np.array_equal(exp["pos"], pos.cpu().numpy()[sample]) and np.array_equal(exp["var"], var.cpu().numpy()[sample])
    # the device row counter (roofline diagnostics): between 1 and H-1 rows per pivot, per LP
    r = rows.cpu().numpy()
    p = piv.sum(dim=1).cpu().numpy()
    assert (r >= p).all() and (r <= p * (H - 1)).all()


def count_rewritten_rows(mats, H, W):
    """R of SURVEY 8(d) summed over the pivots of each LP, from the oracle stepped one pivot at a time."""
    out = []
    for m0 in mats:
        m = m0.copy()
        pos = np.arange(W + H, dtype=np.int32)
        var = pos.copy()
        total = 0
        for phase_budget in range(10 ** 6):
            before = m.reshape(H, W).copy()
            st, _, piv = O.simplex(m, W, H, pos, var, max_pivots=1)
            if sum(piv) == 0:
                break
            changed = (before.view(np.uint64) != m.reshape(H, W).view(np.uint64)).any(axis=1)
            # a rewritten row changes at least its pivot-column cell (it becomes -coef/q != coef unless q == -1)
            total += int(changed.sum()) - 1  # minus the pivot row
            if st != 4:
                break
        out.append(total)
    return out


def test_row_counter_matches_a_cpu_count(engine):
    """yalps_set_row_counter: the rows rewritten per LP, per kernel family, against the oracle stepped pivot by pivot
    (rows whose bits change; dense random tableaus, where a rewritten row always changes)."""
    import torch
    m_, nv, n = 12, 20, 40
    H, W = m_ + 1, nv + 1
    mats = O.generate_synthetic(4100, n, m_, nv, 3)
    d = torch.from_numpy(mats.reshape(-1)).cuda()
    work = torch.empty_like(d)
    piv = torch.empty(n, 2, dtype=torch.int64, device="cuda")
    pos = torch.empty(n, W + H, dtype=torch.int32, device="cuda")
    var = torch.empty(n, W + H, dtype=torch.int32, device="cuda")
    exp = None
    for path, threads, rgroups in ((E.PATH_TMEM, 0, 0), (E.PATH_SMEM, 32, 0), (E.PATH_SMEM, 128, 4), (E.PATH_GMEM, 64, 0),
                                   (E.PATH_GMEM, 128, 2), (E.PATH_CLUSTER, 0, 0)):
        rows = torch.zeros(n, dtype=torch.int64, device="cuda")
        engine.set_tuning(path, threads, rgroups)
        engine.set_row_counter(rows.data_ptr(), per_lp=True)
        try:
            engine.solve_batch_device(n, H, W, d.data_ptr(), d_work=work.data_ptr(), d_pivots=piv.data_ptr(),
                                      d_pos=pos.data_ptr(), d_var=var.data_ptr(),
                                      stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
        finally:
            engine.set_row_counter(0)
            engine.set_tuning(E.PATH_AUTO, 0, 0)
        got = rows.cpu().numpy().tolist()
        if exp is None:
            # every LP runs to the end under the default budget, so stepping one pivot at a time is the same trajectory
            exp = count_rewritten_rows(mats[:8], H, W)
            assert got[:8] == exp, (path, got[:8], exp)
            first = got
        assert got == first, f"path {path} threads {threads} row groups {rgroups}"


def test_netlib_ok_list_with_infinite_pivot_budget(engine):
    """The reference's own benchmark runner calls solve with maxPivots: Infinity (benchmarks/runners.ts:10).  For the
    models its harness handles, neither phase reaches the default budget, so the unbounded-budget trajectory is the
    golden one: status, value, pivot counts, final basis and RHS bits."""
    checked = 0
    for name in NL.names:
        g = NL.get(name)
        if g["list"] == 2 or max(g["pivots"]) >= 8192:  # list: 0 = small, 1 = the harness's ok list, 2 = its "cannot handle" list
            continue
        got = engine.solve_batch(g["matrix"], g["height"], g["width"],
                                 E.make_options(check_cycles=g["check_cycles"], max_pivots=math.inf))
        assert got["status"][0] == g["status"] and tuple(got["pivots"][0]) == g["pivots"], name
        assert same_value(got["value"][0], g["value"]), name
        assert np.array_equal(got["pos"][0], g["final_pos"]) and same_bits(got["rhs"][0], g["final_rhs"]), name
        checked += 1
    assert checked >= 30


# ---------------------------------------------------------------------------------------------- K1t tensor-memory kernel
@pytest.mark.parametrize("m,nv,neg", [(32, 64, 0), (32, 64, 8), (32, 64, 32), (1, 1, 0), (5, 3, 2), (7, 40, 3),
                                      (32, 7, 10), (16, 33, 4), (31, 63, 9), (20, 64, 20), (32, 1, 1), (1, 64, 1),
                                      (8, 64, 0), (9, 64, 5), (24, 48, 0),
                                      # the 256-column shape (34..65 rows, two RHS cells per lane)
                                      (64, 64, 0), (64, 64, 20), (64, 64, 64), (33, 64, 5), (35, 32, 8), (40, 10, 3),
                                      (64, 1, 1), (50, 33, 7), (57, 64, 30)])
def test_tensor_memory_kernel_bit_exact(engine, m, nv, neg):
    """K1t: one warp per LP, tableau rows in tensor memory (at most 65 x 65); more LPs than resident warps so
    that the queue, TMEM reuse across LPs and all four lane quarters are exercised."""
    n = 4 * 148 * 4 + 37 if (m, nv) in ((32, 64), (64, 64)) else 200
    mats = O.generate_synthetic(78000, n, m, nv, neg)
    exp = oracle_batch(mats, m + 1, nv + 1)
    engine.set_tuning(E.PATH_TMEM, 0)
    try:
        got = engine.solve_batch(mats, m + 1, nv + 1, want_matrices=True)
    finally:
        engine.set_tuning(E.PATH_AUTO, 0)
    assert_batch_equal(got, exp, f"tmem {m}x{nv}")


def test_tensor_memory_kernel_options_special_values_and_sparse(engine):
    H, W = 4, 5
    t = np.zeros((4, H * W))
    t[0].reshape(H, W)[0, 1] = 1.0
    t[0].reshape(H, W)[1:, 1] = -1.0
    t[0].reshape(H, W)[1:, 0] = 1.0
    t[1].reshape(H, W)[1, 0] = -1.0
    t[1].reshape(H, W)[1, 1:] = 1.0
    t[2].reshape(H, W)[:] = [[0.0, 3.0, 2.0, -0.0, 1.0], [4.0, 1.0, 1e-16, 1.0000000000000001e-16, -0.0],
                             [5.0, 2e-16, 1.0, -1e-17, 3.0], [6.0, -0.0, 2.0, 1.0, 1e-15]]
    t[3].reshape(H, W)[:] = [[0, -1, -1, 2, 2], [-1, -1, -1, -1, -1], [-1, -1, -1, -1, -1], [3, 1, 1, 1, 1]]
    mats = O.generate_synthetic(5, 64, 32, 64, 8)
    # sparse tableaus (zero coefficients: skipped rows, flushed pivot-row cells, signed zeros)
    rng = np.random.default_rng(3)
    sparse = O.generate_synthetic(9, 96, 32, 64, 6).reshape(96, 33, 65).copy()
    mask = rng.random(sparse.shape) < 0.6
    mask[:, :, 0] = False
    sparse[mask] = 0.0
    sparse[:, 1:, 1:][rng.random(sparse[:, 1:, 1:].shape) < 0.05] = -0.0
    sparse = sparse.reshape(96, -1)
    engine.set_tuning(E.PATH_TMEM, 0)
    try:
        assert_batch_equal(engine.solve_batch(t, H, W, want_matrices=True), oracle_batch(t, H, W), "tmem special")
        assert_batch_equal(engine.solve_batch(sparse, 33, 65, want_matrices=True), oracle_batch(sparse, 33, 65), "tmem sparse")
        for mp in (0, 1, 5, 11.5, math.inf):
            for prec in (1e-8, 1e-3, 0.0):
                exp = oracle_batch(mats, 33, 65, max_pivots=mp, precision=prec)
                got = engine.solve_batch(mats, 33, 65, E.make_options(max_pivots=mp, precision=prec), want_matrices=True)
                assert_batch_equal(got, exp, f"tmem maxPivots={mp} precision={prec}")
        with pytest.raises(Exception):
            engine.solve_batch(np.zeros(66 * 65), 66, 65)  # one row too many for the tensor-memory kernel
    finally:
        engine.set_tuning(E.PATH_AUTO, 0)


def test_ragged_batch_mixed_sizes_are_grouped(engine):
    """A ragged batch with tiny, medium and beyond-shared-memory tableaus: each size class gets its own kernel
    configuration (K1 of several widths, K2/K4), results come back in the caller's order."""
    rng = np.random.default_rng(11)
    shapes_mn = [(3, 4), (32, 64), (200, 260), (5, 9), (90, 130), (250, 300), (32, 64), (1, 1), (60, 200)] * 3
    rng.shuffle(shapes_mn)
    tabs, shapes, exp = [], [], []
    for i, (m, nv) in enumerate(shapes_mn):
        t = O.generate_synthetic(100 + i, 1, m, nv, min(m, 3))[0]
        tabs.append(t)
        shapes.append((m + 1, nv + 1))
        exp.append(oracle_batch(t.reshape(1, -1), m + 1, nv + 1))
    got = engine.solve_ragged(tabs, shapes, want_matrices=True)
    for g, e, s in zip(got, exp, shapes):
        assert g["status"] == e["status"][0] and g["pivots"] == tuple(e["pivots"][0]), s
        assert same_value(g["value"], e["value"][0])
        assert np.array_equal(g["pos"], e["pos"][0]) and np.array_equal(g["var"], e["var"][0])
        assert same_bits(g["rhs"], e["rhs"][0]) and same_bits(g["matrix"], e["matrices"][0]), s


@pytest.mark.parametrize("m,nv,cap", [(1024, 2048, 60), (4096, 8192, 12)])
def test_config5_full_size_capped_pivots_against_oracle(engine, m, nv, cap):
    """BASELINE.json config 5 at full size (4097 x 8193 tableau, 268 MB): the first `cap` pivots of the grid-wide
    kernel against the oracle, bit for bit over the whole tableau (a full solve is too slow for the CPU oracle in a
    unit test; `maxPivots` is part of the reference's interface, so a capped run is a legitimate trajectory prefix)."""
    import torch
    H, W = m + 1, nv + 1
    d = torch.empty(H * W, dtype=torch.float64, device="cuda")
    engine.generate_synthetic_device(0, 1, m, nv, d.data_ptr())
    torch.cuda.synchronize()
    mats = d.cpu().numpy().reshape(1, -1)
    exp = oracle_batch(mats, H, W, max_pivots=cap)
    got = engine.solve_batch(mats, H, W, E.make_options(max_pivots=cap), want_matrices=True)
    assert_batch_equal(got, exp, f"config5 {m}x{nv}")
    assert got["status"][0] == 4 and int(got["pivots"][0].sum()) == cap


def test_kernel_paths_agree_on_a_mid_size_batch(engine):
    """K1, K2 and K4 are independent formulations: on the same inputs they must agree bit for bit with each other
    (and with the oracle) -- 24 LPs of 120 x 200."""
    m, nv, n = 120, 200, 24
    mats = O.generate_synthetic(31, n, m, nv, 25)
    exp = oracle_batch(mats, m + 1, nv + 1)
    outs = {}
    for name, path in (("K1", E.PATH_SMEM), ("K2", E.PATH_GMEM), ("K4", E.PATH_GRID)):
        engine.set_tuning(path, 0)
        outs[name] = engine.solve_batch(mats, m + 1, nv + 1, want_matrices=True)
    engine.set_tuning(E.PATH_AUTO, 0)
    for name, got in outs.items():
        assert_batch_equal(got, exp, name)


def test_host_pipeline_with_several_chunks(engine):
    """The host entry point cuts large batches into double-buffered chunks (H2D / kernel / D2H overlap): 4,400 LPs of
    33 x 65 (75 MB) span two chunks; every LP must come back in order and bit-exact."""
    n, m, nv = 4400, 32, 64
    mats = O.generate_synthetic(123456, n, m, nv, 3)
    exp = oracle_batch(mats, m + 1, nv + 1)
    got = engine.solve_batch(mats, m + 1, nv + 1)
    got["matrices"] = None
    assert_batch_equal(got, exp, "chunked host pipeline")


def test_config2_full_size_tensor_memory_against_shared_memory_kernel(engine):
    """BASELINE config 2 at full size (65,536 LPs 33x65 generated on the device): K1t and K1 must agree on every
    output bit of every LP -- status, value, pivot counts per phase, RHS column, both basis arrays."""
    import torch
    m, nv, n = 32, 64, 65536
    H, W = m + 1, nv + 1
    d = torch.empty(n * H * W, dtype=torch.float64, device="cuda")
    engine.generate_synthetic_device(0, n, m, nv, d.data_ptr(), neg_rows=3)
    outs = []
    for path in (E.PATH_SMEM, E.PATH_TMEM):
        st = torch.empty(n, dtype=torch.int32, device="cuda")
        val = torch.empty(n, dtype=torch.float64, device="cuda")
        piv = torch.empty(n, 2, dtype=torch.int64, device="cuda")
        rhs = torch.empty(n * H, dtype=torch.float64, device="cuda")
        pos = torch.empty(n * (W + H), dtype=torch.int32, device="cuda")
        var = torch.empty(n * (W + H), dtype=torch.int32, device="cuda")
        engine.set_tuning(path, 0)
        try:
            engine.solve_batch_device(n, H, W, d.data_ptr(), d_status=st.data_ptr(), d_value=val.data_ptr(),
                                      d_pivots=piv.data_ptr(), d_rhs=rhs.data_ptr(), d_pos=pos.data_ptr(), d_var=var.data_ptr())
            torch.cuda.synchronize()
        finally:
            engine.set_tuning(E.PATH_AUTO, 0)
        outs.append((st, val.view(torch.int64), piv, rhs.view(torch.int64), pos, var))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    assert int(outs[0][2].sum()) > 10 * n  # real work: more than ten pivots per LP on average


def test_tensor_memory_kernel_on_a_ragged_batch(engine):
    """K1t through yalps_solve_ragged: LPs of many different shapes (all within 33 x 65) share launches, every warp
    reads its LP's own height / width / offsets, TMEM rows left over from a taller previous LP must not leak."""
    rng = np.random.default_rng(5)
    tabs, shapes, exp = [], [], []
    for i in range(420):
        m, nv = int(rng.integers(1, 65)), int(rng.integers(1, 65))  # both TMEM shapes (up to 33 / up to 65 rows)
        t = O.generate_synthetic(9000 + i, 1, m, nv, int(rng.integers(0, m + 1)))[0]
        tabs.append(t)
        shapes.append((m + 1, nv + 1))
        exp.append(oracle_batch(t.reshape(1, -1), m + 1, nv + 1))
    engine.set_tuning(E.PATH_TMEM, 0)
    try:
        got = engine.solve_ragged(tabs, shapes, want_matrices=True)
    finally:
        engine.set_tuning(E.PATH_AUTO, 0)
    for g, e, s in zip(got, exp, shapes):
        assert g["status"] == e["status"][0] and g["pivots"] == tuple(e["pivots"][0]), s
        assert same_value(g["value"], e["value"][0])
        assert np.array_equal(g["pos"], e["pos"][0]) and np.array_equal(g["var"], e["var"][0])
        assert same_bits(g["rhs"], e["rhs"][0]) and same_bits(g["matrix"], e["matrices"][0]), s


def test_ragged_batch_with_many_small_lps_on_the_automatic_path(engine):
    """solve_many-style input: 900 LPs of two shapes in one ragged call.  Each size class has more than two LPs per SM,
    so the automatic policy puts the classes on the tensor-memory kernel (index-mapped launch-local LP ids)."""
    tabs, shapes, exp = [], [], []
    for i in range(900):
        m, nv = ((8, 16), (32, 64))[i % 2]
        t = O.generate_synthetic(20000 + i, 1, m, nv, i % 5)[0]
        tabs.append(t)
        shapes.append((m + 1, nv + 1))
        exp.append(oracle_batch(t.reshape(1, -1), m + 1, nv + 1))
    before = engine.launch_count
    got = engine.solve_ragged(tabs, shapes, want_matrices=True)
    assert engine.launch_count - before <= 4
    for g, e, s in zip(got, exp, shapes):
        assert g["status"] == e["status"][0] and g["pivots"] == tuple(e["pivots"][0]), s
        assert same_value(g["value"], e["value"][0])
        assert np.array_equal(g["pos"], e["pos"][0]) and np.array_equal(g["var"], e["var"][0])
        assert same_bits(g["rhs"], e["rhs"][0]) and same_bits(g["matrix"], e["matrices"][0]), s


@pytest.mark.parametrize("mode", [1, 2])
def test_cluster_kernel_tma_staging_variants_are_bit_exact(mode):
    """KC with the winner's pivot row staged by cp.async.bulk (1) or by one multicast bulk copy into every CTA of the
    cluster (2) instead of the ld.global.cg loop: same bits (the A/B of profiles/r02_kc_tma_ab.jsonl)."""
    import os
    import yalps_b200
    os.environ["YALPS_KC_TMA"] = str(mode)
    try:
        eng = yalps_b200.Engine(0)
    finally:
        del os.environ["YALPS_KC_TMA"]
    try:
        eng.set_tuning(E.PATH_CLUSTER, 0)
        for (m, nv, neg, n) in ((200, 260, 30, 3), (90, 400, 10, 2), (300, 150, 80, 2)):
            mats = O.generate_synthetic(6100 + m, n, m, nv, neg)
            assert_batch_equal(eng.solve_batch(mats, m + 1, nv + 1, want_matrices=True), oracle_batch(mats, m + 1, nv + 1),
                               f"KC tma mode {mode} {m}x{nv}")
        g = NL.get("SC205")
        got = eng.solve_batch(g["matrix"], g["height"], g["width"])
        assert got["status"][0] == g["status"] and tuple(got["pivots"][0]) == g["pivots"]
        assert np.array_equal(got["pos"][0], g["final_pos"]) and same_bits(got["rhs"][0], g["final_rhs"])
    finally:
        eng.close()


def test_tensor_memory_kernel_with_check_cycles(engine):
    """K1t with checkCycles: true (src/simplex.ts:44-63): the instantiation with a per-warp history.  The Chvatal LP
    must stop as "cycled" exactly where the reference does; LPs that do not cycle are unchanged by the option; and a
    mixed batch keeps every LP's own history (one LP per warp, several LPs per warp in sequence)."""
    chvatal = np.array([[0, 10, -57, -9, -24], [0, 0.5, -5.5, -2.5, 9], [0, 0.5, -1.5, -0.5, 1], [1, 1, 0, 0, 0]], float)
    engine.set_tuning(E.PATH_TMEM, 0)
    try:
        for mp in (8192, math.inf, 7, 30):
            m = chvatal.reshape(1, -1).copy()
            exp = oracle_batch(m, 4, 5, check_cycles=True, max_pivots=mp)
            got = engine.solve_batch(m, 4, 5, E.make_options(check_cycles=True, max_pivots=mp), want_matrices=True)
            assert_batch_equal(got, exp, f"tmem chvatal maxPivots={mp}")
        assert exp["status"][0] == 4
        # 700 copies interleaved with ordinary LPs of the same shape: more LPs than warps, histories must not leak
        rng = np.random.default_rng(11)
        n = 1400
        mats = O.generate_synthetic(321, n, 3, 4, 1)
        mats[::2] = chvatal.reshape(-1)
        exp = oracle_batch(mats, 4, 5, check_cycles=True)
        got = engine.solve_batch(mats, 4, 5, E.make_options(check_cycles=True), want_matrices=True)
        assert_batch_equal(got, exp, "tmem mixed batch with checkCycles")
        assert (exp["status"][::2] == 4).all()
        for (m_, nv, neg, k) in ((32, 64, 8, 300), (50, 40, 10, 64), (12, 20, 12, 100)):
            mats = O.generate_synthetic(555 + m_, k, m_, nv, neg)
            exp = oracle_batch(mats, m_ + 1, nv + 1, check_cycles=True)
            got = engine.solve_batch(mats, m_ + 1, nv + 1, E.make_options(check_cycles=True), want_matrices=True)
            assert_batch_equal(got, exp, f"tmem checkCycles {m_}x{nv}")
    finally:
        engine.set_tuning(E.PATH_AUTO, 0)
    # the automatic path now keeps small checkCycles batches on the tensor-memory kernel as well
    mats = O.generate_synthetic(9, 500, 20, 30, 5)
    assert_batch_equal(engine.solve_batch(mats, 21, 31, E.make_options(check_cycles=True), want_matrices=True),
                       oracle_batch(mats, 21, 31, check_cycles=True), "auto path with checkCycles")
