"""Secondary workloads of bench.py: BASELINE.json configs 3, 4 and 5 (config 2 is the headline, config 1 the parity
anchor).  Every entry of the `secondary` array carries ms, pivots/s, LPs/s and a roofline whose bytes follow
SURVEY 8(d) exactly:

    bytes(pivot) = 16*W*(1+R) + 8*(2(H-1) + 2(W-1)),   R = rows the rank-1 update rewrites (|coef| > 1e-16)

with R COUNTED ON THE DEVICE (yalps_set_row_counter: the kernels add the R of every pivot to a 64-bit counter) in a
measuring pass over the same inputs; the timed passes run without the counter.  `traffic` is the DRAM traffic of the
same kernel from the committed ncu --set full summaries under profiles/ (null where none was captured).

Inputs come from tests/golden/ (the Netlib base tableaus and the reference's MILP test cases, generated from the
reference by tests/golden/make_golden.py); nothing here reads /root/reference.
"""
import gzip
import json
import math
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def netlib_base(name):
    z = np.load(os.path.join(GOLDEN, "netlib.npz"))
    h, w = (int(x) for x in z[f"{name}/shape"])
    m = np.zeros(h * w, np.float64)
    m[z[f"{name}/nz_idx"]] = z[f"{name}/nz_val"]
    m[z[f"{name}/neg_zero_idx"]] = -0.0
    return {"height": h, "width": w, "matrix": m, "row_groups": z[f"{name}/row_groups"],
            "pivots": tuple(int(x) for x in z[f"{name}/pivots"]), "status": int(z[f"{name}/status"][0])}


def milp_case(name):
    with gzip.open(os.path.join(GOLDEN, "cases.json.gz"), "rt", encoding="utf-8") as f:
        cases = json.load(f)
    c = next(x for x in cases if x["name"] == name)
    m = c["model"]
    m["constraints"] = [(k, v) for k, v in m["constraints"]]
    m["variables"] = [(k, [(ck, cv) for ck, cv in v]) for k, v in m["variables"]]
    for key in ("integers", "binaries", "direction", "objective"):
        if m.get(key) is None:
            m.pop(key, None)
    return c


def pivot_bytes(H, W, pivots, rows):
    """SURVEY 8(d) summed over `pivots` pivots that rewrote `rows` rows in total."""
    return 16.0 * W * (pivots + rows) + pivots * 8.0 * (2 * (H - 1) + 2 * (W - 1))


def ncu_traffic(name, lps=None):
    """DRAM bytes per launch from a committed ncu --set full summary; captures taken on a smaller launch of the same
    kernel carry `dram_bytes_per_lp` and are scaled to the `lps` LPs of the benchmarked launch."""
    try:
        j = json.load(open(os.path.join(ROOT, "profiles", name)))
        if lps is not None and "dram_bytes_per_lp" in j:
            return j["dram_bytes_per_lp"] * lps
        return j["dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        return None


def _events(torch, fn, reps, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for e0, e1 in evs:
        e0.record()
        fn()
        e1.record()
    torch.cuda.synchronize()
    return sum(e0.elapsed_time(e1) for e0, e1 in evs) / reps


def _counter(torch, eng, fn):
    """One pass of fn() with the device row counter set: total rows rewritten."""
    c = torch.zeros(1, dtype=torch.int64, device="cuda")
    eng.set_row_counter(c.data_ptr(), per_lp=False)
    try:
        fn()
        torch.cuda.synchronize()
    finally:
        eng.set_row_counter(0)
    return int(c.item())


def config3(torch, eng, name, n, hbm_peak, cpu_cores, l2_peak):
    """n RHS-perturbed replicas of a Netlib model, working copies in HBM (the automatic policy picks the kernel).
    One step = one launch: the kernel reads the pristine replicas from `d`, writes its working copy into `work` and
    solves there, so the tableau crosses HBM once in and once out per LP."""
    from yalps_b200.engine import make_options
    from oracle import lib as O
    g = netlib_base(name)
    H, W = g["height"], g["width"]
    cells = H * W
    opt = make_options()
    stream = torch.cuda.current_stream().cuda_stream
    d = torch.empty(n * cells, dtype=torch.float64, device="cuda")
    eng.generate_replicas_device(0, n, g["matrix"], H, W, g["row_groups"], d.data_ptr(), stream=stream)
    work = torch.empty_like(d)
    st = torch.empty(n, dtype=torch.int32, device="cuda")
    val = torch.empty(n, dtype=torch.float64, device="cuda")
    piv = torch.empty(n, 2, dtype=torch.int64, device="cuda")
    rhs = torch.empty(n, H, dtype=torch.float64, device="cuda")
    pos = torch.empty(n, W + H, dtype=torch.int32, device="cuda")
    var = torch.empty(n, W + H, dtype=torch.int32, device="cuda")

    def step():
        eng.solve_batch_device(n, H, W, d.data_ptr(), opt, d_work=work.data_ptr(), d_status=st.data_ptr(),
                               d_value=val.data_ptr(), d_pivots=piv.data_ptr(), d_rhs=rhs.data_ptr(),
                               d_pos=pos.data_ptr(), d_var=var.data_ptr(), stream=stream)

    launches0 = eng.launch_count
    step()
    torch.cuda.synchronize()
    launches = eng.launch_count - launches0
    rows = _counter(torch, eng, step)
    ms = _events(torch, step, reps=3, warm=1)
    pivots = int(piv.sum().item())
    optimal = int((st == 0).sum().item())
    alg = pivot_bytes(H, W, pivots, rows)
    achieved = alg / (ms * 1e-3) / 1e9

    # end to end through the data-reducing replica entry: base once + n*H right-hand sides from pinned host memory
    h_rhs = eng.pinned_empty((n, H), np.float64)
    h_rhs[:] = d.view(n, H, W)[:, :, 0].cpu().numpy()
    out = eng.batch_outputs(n, H, W, pinned=True)
    eng.solve_replicas(g["matrix"], h_rhs, H, W, opt, out=out)  # warm-up (allocates the device pool)
    assert int(out["pivots"].sum()) == pivots and int((out["status"] == 0).sum()) == optimal
    t0 = time.perf_counter()
    reps = 2
    for _ in range(reps):
        eng.solve_replicas(g["matrix"], h_rhs, H, W, opt, out=out)
    e2e_s = (time.perf_counter() - t0) / reps
    forks = eng.replica_forks
    shared_bits = (out["pivots"].copy(), out["rhs"].copy(), out["pos"].copy())
    # the same call with path sharing off: every replica expanded to its own working copy and solved from scratch
    eng.set_replica_sharing(False)
    try:
        eng.solve_replicas(g["matrix"], h_rhs, H, W, opt, out=out)
        t0 = time.perf_counter()
        eng.solve_replicas(g["matrix"], h_rhs, H, W, opt, out=out)
        e2e_plain_s = time.perf_counter() - t0
    finally:
        eng.set_replica_sharing(True)
    sharing_same = bool(np.array_equal(shared_bits[0], out["pivots"]) and np.array_equal(shared_bits[2], out["pos"]) and
                        np.array_equal(shared_bits[1].view(np.uint64), out["rhs"].view(np.uint64)))

    # CPU port on a bounded sample, all cores
    cn = min(n, 64 * cpu_cores)
    sample = d.view(n, cells)[:cn].cpu().numpy().copy()
    t0 = time.perf_counter()
    ref = O.simplex_batch(sample, W, H, nthreads=cpu_cores, want_pos=False)
    cpu_dt = time.perf_counter() - t0
    same = bool(np.array_equal(ref["pivots"], piv[:cn].cpu().numpy()) and
                np.array_equal(ref["rhs"].view(np.uint64), rhs[:cn].cpu().numpy().view(np.uint64)))
    return {
        "workload": f"config3_{name.lower()}: {n} RHS-perturbed (eps 1e-2) replicas of Netlib {name}, tableau {H}x{W}, "
                    f"working copies in HBM ({n * cells * 8 / 1e9:.1f} GB, larger than L2), BASELINE.json configs[2]",
        "ms": ms, "pivots_per_s": pivots / ms * 1e3, "lps_per_s": n / ms * 1e3, "pivots": pivots, "optimal": optimal,
        "gpu_launches": launches,
        "roofline": {"bound": "l2", "achieved": achieved, "peak": l2_peak, "unit": "GB/s", "frac": achieved / l2_peak,
                     "peak_source": "measured live: ld-mul-sub-st stream over a 48 MB L2-resident buffer "
                                    "(yalps_measure_l2_bandwidth)",
                     "traffic": ncu_traffic(f"r02_k2s_{name.lower()}_ncu_summary.json", n),
                     "traffic_note": "dram__bytes_read+write of the same kernel (ncu --set full on a launch of a quarter of "
                                     "the LPs, scaled per LP; profiles/r02_k2s_*_ncu_summary.json)",
                     "bytes_formula": "16*W*(pivots + rows_rewritten) + pivots*8*(2(H-1)+2(W-1)), SURVEY 8(d)",
                     "rows_rewritten": rows, "mean_rows_per_pivot": rows / max(pivots, 1), "rows_dense": H - 1,
                     "hbm": {"achieved": achieved, "peak": hbm_peak, "frac": achieved / hbm_peak,
                             "note": "the SURVEY 8(d) bytes against the HBM peak can exceed 1: a working copy "
                                     f"({cells * 8 / 1e3:.0f} KB) stays in L2 for the ~{pivots // max(n, 1)} pivots its CTA "
                                     "spends on it, so HBM sees far fewer bytes than the pivots touch (`traffic`: what the live copies that no longer fit L2 cost in DRAM)"},
                     "note": "R counted on the device.  The medium that serves the per-pivot bytes is L2, hence the "
                             "denominator; the kernel itself is bound by the per-pivot dependent chain of one CTA per LP"},
        "e2e": {"api": "yalps_solve_replicas (base tableau once + n*H right-hand sides from pinned host memory; "
                       "status/value/pivots/RHS/basis back)", "ms": e2e_s * 1e3, "lps_per_s": n / e2e_s,
                "pivots_per_s": pivots / e2e_s, "h2d_bytes_per_step": cells * 8 + n * H * 8,
                "d2h_bytes_per_step": n * (4 + 8 + 16 + H * 8 + 2 * (W + H) * 4),
                "path_sharing": {"replicas_that_left_the_shared_path": forks, "of": n,
                                 "ms_without_sharing": e2e_plain_s * 1e3, "speedup": e2e_plain_s / e2e_s,
                                 "bit_identical_to_without": sharing_same,
                                 "note": "the replicas follow the recorded pivot trace of the base tableau with their RHS "
                                         "column only and continue alone from the leader's snapshot where they would choose "
                                         "differently (include/yalps_b200.h: yalps_set_replica_sharing); `ms` above is WITH "
                                         "sharing, the kernel-only line of this workload is the plain HBM-resident batch"}},
        "cpu_baseline": {"value": int(ref["pivots"].sum()) / cpu_dt, "unit": "pivots/s", "lps_per_s": cn / cpu_dt,
                         "cores": cpu_cores, "kind": "port", "sample": f"the first {cn} replicas, {cpu_dt:.2f} s",
                         "bit_identical_to_gpu": same},
    }


def config4(torch, eng, hbm_peak):
    """The MILP suite of benchmarks/json/read.ts through solve(): host tableau build + root LP + branch and cut."""
    import yalps_b200
    from oracle import model as M
    out = []
    for name in ("Large Farm MIP", "Monster 2", "Vendor Selection", "Monster Problem"):
        c = milp_case(name)
        info = {}
        yalps_b200.solve(c["model"], c["options"], engine=eng, info=info)  # warm-up
        rows = _counter(torch, eng, lambda: yalps_b200.solve(c["model"], c["options"], engine=eng))
        launches0 = eng.launch_count
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            sol = yalps_b200.solve(c["model"], c["options"], engine=eng, info=info)
        dt = (time.perf_counter() - t0) / reps
        launches = (eng.launch_count - launches0) // reps
        t0 = time.perf_counter()
        ref = M.solve(c["model"], {**M.DEFAULT_OPTIONS, **c["options"]})
        cpu = time.perf_counter() - t0
        H, W = info["height"], info["width"]
        piv = sum(info["root_pivots"]) + info["node_pivots"]
        alg = pivot_bytes(H, W, piv, rows)  # node tableaus have a few more rows than the root: a lower bound
        same = sol["status"] == ref["status"] and (sol["result"] == ref["result"] or
                                                   (math.isnan(sol["result"]) and math.isnan(ref["result"])))
        out.append({
            "workload": f"config4_{name.lower().replace(' ', '_')}: {name} via solve() (tableau {H}x{W}, "
                        f"{len(c['model'].get('integers') or [])} integer variables), BASELINE.json configs[3]",
            "ms": dt * 1e3, "pivots_per_s": piv / dt, "lps_per_s": (1 + info["device_nodes"]) / dt,
            "nodes": info["nodes"], "node_pivots": info["node_pivots"], "root_pivots": list(info["root_pivots"]),
            "waves": info["waves"], "device_nodes": info["device_nodes"], "nodes_per_s": info["nodes"] / dt,
            "status": sol["status"], "result": sol["result"], "expected": c["expected"]["result"],
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": alg / dt / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": alg / dt / 1e9 / hbm_peak, "traffic": None, "rows_rewritten": rows,
                         "note": "latency-bound by construction: a search is a chain of dependent waves of a few "
                                 "node LPs (SURVEY 8d); wall time includes the host tableau build and the replay"},
            "cpu_baseline": {"ms": cpu * 1e3, "cores": 1, "kind": "port", "sample": "the same solve(), once",
                             "same_status_and_result": bool(same)},
        })
    return out


def config1(torch, eng, hbm_peak):
    """BASELINE.json configs[0], the reference's own CPU-runnable case: Netlib AFIRO from its MPS file through the
    host MPS reader, the benchmarks/netlib/read.ts conversion and solve() -- the single-LP latency anchor."""
    import yalps_b200
    from oracle import model as M
    from yalps_b200.mps import netlib_model
    text = open(os.path.join(GOLDEN, "afiro.mps")).read()
    model = netlib_model(text)
    info = {}
    for _ in range(3):
        yalps_b200.solve(model, engine=eng, info=info)
    rows = _counter(torch, eng, lambda: yalps_b200.solve(model, engine=eng))
    launches0 = eng.launch_count
    reps = 50
    t0 = time.perf_counter()
    for _ in range(reps):
        sol = yalps_b200.solve(model, engine=eng, info=info)
    dt = (time.perf_counter() - t0) / reps
    launches = (eng.launch_count - launches0) // reps
    t0 = time.perf_counter()
    for _ in range(reps):
        ref = M.solve(model, dict(M.DEFAULT_OPTIONS))
    cpu = (time.perf_counter() - t0) / reps
    H, W = info["height"], info["width"]
    piv = sum(info["root_pivots"])
    alg = pivot_bytes(H, W, piv, rows)
    return {
        "workload": f"config1_afiro: Netlib AFIRO (afiro.mps -> model -> tableau {H}x{W}) via solve(), "
                    "BASELINE.json configs[0]",
        "ms": dt * 1e3, "pivots_per_s": piv / dt, "lps_per_s": 1.0 / dt, "pivots": piv,
        "root_pivots": list(info["root_pivots"]), "status": sol["status"], "result": sol["result"],
        "expected": -464.75314286, "matches_index_json": bool(abs(sol["result"] + 464.75314286) <= 1e-5 * 464.75314286),
        "gpu_launches": launches,
        "roofline": {"bound": "smem", "achieved": alg / dt / 1e9, "peak": None, "unit": "GB/s", "frac": None,
                     "traffic": None, "rows_rewritten": rows,
                     "note": "latency-bound by construction: ONE 9.5 KB tableau, 20 dependent pivots on one SM; wall "
                             "time = host tableau build + one zero-copy launch + synchronise (no roofline applies)"},
        "cpu_baseline": {"ms": cpu * 1e3, "cores": 1, "kind": "port", "sample": f"the same solve(), {reps} times",
                         "same_status_and_result": bool(sol["status"] == ref["status"] and sol["result"] == ref["result"])},
    }


def config5(torch, eng, hbm_peak, l2_peak):
    """One large LP on the whole GPU: the 4097x8193 synthetic tableau (268.5 MB > L2: K4, rows in HBM) for a capped
    number of pivots; the 1025x2049 one (16.8 MB) and Netlib 25FV47 (1338x1572, to the end) on KG, the grid-resident
    kernel that keeps the rows in the SMs' shared memory."""
    from yalps_b200.engine import make_options
    from oracle import lib as O
    out = []
    stream = torch.cuda.current_stream().cuda_stream
    cases = [("synthetic 4096x8192", None, 4096, 8192, 48.0), ("synthetic 1024x2048", None, 1024, 2048, 200.0),
             ("Netlib 25FV47", "25FV47", 0, 0, math.inf)]
    smem_peak = eng.measure_smem_bandwidth()[0]
    for label, netlib, m, nv, cap in cases:
        if netlib:
            g = netlib_base(netlib)
            H, W = g["height"], g["width"]
            d = torch.from_numpy(g["matrix"]).cuda()
        else:
            H, W = m + 1, nv + 1
            d = torch.empty(H * W, dtype=torch.float64, device="cuda")
            eng.generate_synthetic_device(0, 1, m, nv, d.data_ptr(), stream=stream)
        work = torch.empty_like(d)
        st = torch.empty(1, dtype=torch.int32, device="cuda")
        piv = torch.empty(1, 2, dtype=torch.int64, device="cuda")
        opt = make_options(max_pivots=cap)
        copy_ms = _events(torch, lambda: work.copy_(d), reps=3)

        def step():
            work.copy_(d)
            eng.solve_batch_device(1, H, W, work.data_ptr(), opt, d_work=work.data_ptr(), d_status=st.data_ptr(),
                                   d_pivots=piv.data_ptr(), stream=stream)

        launches0 = eng.launch_count
        step()
        torch.cuda.synchronize()
        launches = eng.launch_count - launches0
        rows = _counter(torch, eng, step)
        ms = _events(torch, step, reps=2, warm=0) - copy_ms
        p = int(piv.sum().item())
        alg = pivot_bytes(H, W, p, rows)
        resident = H * W * 8 < 24e6 and W <= 2049  # what plan_gridres (csrc/yalps_b200.cu) accepts: KG instead of K4
        entry = {
            "kernel": "k_simplex_cluster<..., kGrid> (KG: rows resident in the shared memory of all SMs, one all-to-all of "
                      "selection records per pivot)" if resident else "k_simplex_grid (K4: rows in HBM/L2, two grid barriers per pivot)",
            "workload": f"config5: one LP on the whole GPU, {label} (tableau {H}x{W}, {H * W * 8 / 1e6:.1f} MB, "
                        f"{'larger than' if H * W * 8 > 126e6 else 'resident in'} L2), BASELINE.json configs[4]"
                        + (f"; first {int(cap)} pivots per phase" if math.isfinite(cap) else "; full solve"),
            "ms": ms, "pivots": p, "pivots_per_s": p / ms * 1e3, "lps_per_s": 1e3 / ms, "us_per_pivot": ms * 1e3 / max(p, 1),
            "status": int(st.item()), "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / hbm_peak,
                         "traffic": ncu_traffic("r02_k4_ncu_summary.json") if not resident else
                         (ncu_traffic("r02_kg_ncu_summary.json") if not netlib else None),
                         "rows_rewritten": rows, "mean_rows_per_pivot": rows / max(p, 1), "rows_dense": H - 1,
                         "note": None if not resident else "rows never leave shared memory: bound by the latency of one all-to-all "
                                                           "of selection records through L2 per pivot, not by bandwidth; "
                                                           "the smem view is the medium's own roofline",
                         "smem": None if not resident else {"achieved": alg / (ms * 1e-3) / 1e9, "peak": smem_peak,
                                                            "frac": alg / (ms * 1e-3) / 1e9 / smem_peak},
                         "l2": None if not resident else {"achieved": alg / (ms * 1e-3) / 1e9, "peak": l2_peak,
                                                          "frac": alg / (ms * 1e-3) / 1e9 / l2_peak}},
        }
        if netlib:  # the reference's own outcome on this model: infeasible after 3110 phase-1 pivots
            entry["matches_oracle_golden"] = bool(int(st.item()) == g["status"] and
                                                  tuple(int(x) for x in piv[0].tolist()) == g["pivots"])
        elif not resident:
            k = 6  # CPU port: the first k pivots of the same tableau, one thread (a pivot is a sequential rank-1 update)
            host = d.cpu().numpy().copy()
            posv = np.arange(W + H, dtype=np.int32)
            varv = posv.copy()
            t0 = time.perf_counter()
            O.simplex(host, W, H, posv, varv, max_pivots=k)
            cpu_dt = time.perf_counter() - t0
            entry["cpu_baseline"] = {"value": k / cpu_dt, "unit": "pivots/s", "cores": 1, "kind": "port",
                                     "sample": f"the first {k} pivots of the same tableau, {cpu_dt:.2f} s"}
        out.append(entry)
        del d, work
    return out


def run_all(torch, eng, hbm_peak, cpu_cores):
    out = []
    l2_peak = eng.measure_l2_bandwidth()
    for name, n in (("SC105", 32768), ("ADLITTLE", 65536)):
        out.append(config3(torch, eng, name, n, hbm_peak, cpu_cores, l2_peak))
        torch.cuda.empty_cache()
    out.extend(config4(torch, eng, hbm_peak))
    out.extend(config5(torch, eng, hbm_peak, l2_peak))
    out.append(config1(torch, eng, hbm_peak))  # (last: the documents cite the entries above by position)
    return out


def config5_multi(torch, eng, devices, m=4096, nv=8192, cap=48.0):
    """the config-5 tableau with its rows dealt over several GPUs (yalps_multi_solve_large, K4m), driven
    by ONE process; device time of the slowest rank's kernel, against K4 on one GPU on the same tableau, bit for bit."""
    import yalps_b200
    from yalps_b200.engine import make_options
    H, W = m + 1, nv + 1
    stream = torch.cuda.current_stream().cuda_stream
    d = torch.empty(H * W, dtype=torch.float64, device="cuda")
    eng.generate_synthetic_device(0, 1, m, nv, d.data_ptr(), stream=stream)
    torch.cuda.synchronize()
    mats = d.cpu().numpy().reshape(1, -1)
    opt = make_options(max_pivots=cap)
    work = torch.empty_like(d)
    piv = torch.empty(1, 2, dtype=torch.int64, device="cuda")

    def one_gpu():
        eng.solve_batch_device(1, H, W, work.data_ptr(), opt, d_work=work.data_ptr(), d_pivots=piv.data_ptr(), stream=stream)

    best1 = None
    for _ in range(3):
        work.copy_(d)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        one_gpu()
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1)
        best1 = t if best1 is None else min(best1, t)
    p = int(piv.sum().item())
    ref_bits = work.cpu().numpy().view(np.uint64)
    del d
    with yalps_b200.MultiEngine(devices) as me:
        ms = None
        launches0 = me.launch_count
        for _ in range(3):
            got = me.solve_large(mats, H, W, opt, want_matrix=True)
            ms = got["kernel_ms"] if ms is None else min(ms, got["kernel_ms"])
        launches = (me.launch_count - launches0) // 3
    same = bool(np.array_equal(got["matrix"].view(np.uint64), ref_bits) and sum(got["pivots"]) == p)
    return {
        "workload": f"config5 over {len(devices)} GPUs (SURVEY 8f-3): one LP, synthetic {m}x{nv} (tableau {H}x{W}, "
                    f"{H * W * 8 / 1e6:.1f} MB), rows dealt round robin over the GPUs, first {int(cap)} pivots; one process, "
                    f"one persistent kernel per GPU, pivot row / column exchanged by peer-memory stores inside the kernels",
        "api": "yalps_multi_solve_large", "n_gpus": len(devices), "pivots": p,
        "ms": ms, "us_per_pivot": ms * 1e3 / max(p, 1), "pivots_per_s": p / ms * 1e3, "gpu_launches": launches,
        "one_gpu": {"kernel": "K4", "ms": best1, "us_per_pivot": best1 * 1e3 / max(p, 1)},
        "speedup_vs_one_gpu": best1 / ms, "bit_identical_to_one_gpu": same, "scaling": "strong",
        "timing": "CUDA events around each rank's kernel, max over ranks, best of 3",
    }
