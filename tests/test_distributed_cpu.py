"""Host-side multi-GPU logic on CPU: world_size 2 over gloo.  The node / LP evaluator is the oracle here (the GPU
engine is injected on the GPU box); what is under test is the sharding, the gathers, the incumbent min-allreduce
and that the sharded branch-and-cut replay reproduces the reference search exactly on every rank."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_cases, same_bits, same_value
from oracle import lib as O, model as M
from yalps_b200 import distributed as D


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 65536, 1000003):
        for world in (1, 2, 3, 8):
            ranges = [D.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1


def oracle_eval_nodes(t, opt):
    def eval_nodes(cut_lists):
        out = []
        for cuts in cut_lists:
            m, p, v = O.apply_cuts(t.matrix, t.width, t.height, t.pos, t.var, [c[0] for c in cuts],
                                   [c[1] for c in cuts], [c[2] for c in cuts])
            h = t.height + len(cuts)
            st, val, piv = O.simplex(m, t.width, h, p, v, opt["precision"], opt["maxPivots"], opt["checkCycles"])
            out.append({"status": st, "value": val, "pivots": sum(piv), "rhs": m.reshape(h, t.width)[:, 0].copy(),
                        "pos": p, "var": v})
        return out
    return eval_nodes


def _worker(rank, world, port, case_name, queue):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case = next(c for c in load_cases() if c["name"] == case_name)
        opt = {**M.DEFAULT_OPTIONS, **case["options"]}
        tm = M.tableau_model(case["model"])
        t = tm.tableau
        st, value, _ = O.simplex(t.matrix, t.width, t.height, t.pos, t.var, opt["precision"], opt["maxPivots"], False)
        rhs = t.matrix.reshape(t.height, t.width)[:, 0].copy()
        res = D.branch_and_cut_sharded(oracle_eval_nodes(t, opt), rhs, t.pos, t.var, t.width, t.height, tm.integers,
                                       tm.sign, value, opt, wave=16, allreduce_every=2)

        # LP sharding: 101 synthetic LPs, each rank solves its range, everybody ends with all rows
        def solve_local(lo, hi):
            mats = O.generate_synthetic(lo, hi - lo, 8, 12, 2)
            r = O.simplex_batch(mats, 13, 9)
            return {"status": r["status"], "value": r["value"], "pivots": r["pivots"], "rhs": r["rhs"]}

        lp = D.solve_batch_sharded(solve_local, 101)
        inc = D.allreduce_min(5.0 + rank)
        queue.put((rank, res["status"], res["result"], res["rhs"].tolist(), res["pos"].tolist(), res["stats"],
                   lp["status"].tolist(), lp["value"].tolist(), lp["pivots"].tolist(), inc))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case_name", ["Knapsack 1", "Fancy Stock Cutting Problem", "Integer Wood Shop Problem"])
def test_sharded_branch_and_cut_world2(case_name):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case_name, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    case = next(c for c in load_cases() if c["name"] == case_name)
    o = case["oracle"]
    sign = M.tableau_model(case["model"]).sign
    ref_lp = O.simplex_batch(O.generate_synthetic(0, 101, 8, 12, 2), 13, 9)
    for rank, status, result, rhs, pos, stats, lp_status, lp_value, lp_piv, inc in outs:
        assert status == "optimal"
        assert same_value(-sign * result, o["result"])
        assert stats["nodes"] == o["nodes"] and stats["node_pivots"] == o["node_pivots"]
        assert pos == o["final_pos"].tolist() and same_bits(np.asarray(rhs), o["final_rhs"])
        assert stats["allreduces"] >= 1
        assert lp_status == ref_lp["status"].tolist() and lp_piv == ref_lp["pivots"].tolist()
        assert same_bits(np.asarray(lp_value), ref_lp["value"])
        assert inc == 5.0
    assert outs[0][1:] == outs[1][1:]  # both ranks hold identical results


def _timeout_worker(rank, world, port, queue):
    """A finite timeout with ranks whose node evaluation takes different wall time: without a collective stop decision
    the fast rank would leave the loop while the slow one enters the next wave's all_gather."""
    import time
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case = next(c for c in load_cases() if c["name"] == "Fancy Stock Cutting Problem")
        opt = {**M.DEFAULT_OPTIONS, **case["options"], "timeout": 60.0}
        tm = M.tableau_model(case["model"])
        t = tm.tableau
        st, value, _ = O.simplex(t.matrix, t.width, t.height, t.pos, t.var, opt["precision"], opt["maxPivots"], False)
        rhs = t.matrix.reshape(t.height, t.width)[:, 0].copy()
        inner = oracle_eval_nodes(t, opt)

        def slow_eval(cut_lists):
            time.sleep(0.015 * (1 + rank))  # rank 1 is twice as slow: the ranks' clocks cross the deadline in different iterations
            return inner(cut_lists)

        res = D.branch_and_cut_sharded(slow_eval, rhs, t.pos, t.var, t.width, t.height, tm.integers, tm.sign, value,
                                       opt, wave=4, allreduce_every=2)
        queue.put((rank, res["status"], repr(res["result"]), res["stats"]["nodes"], res["stats"]["waves"]))
    finally:
        dist.destroy_process_group()


def test_sharded_branch_and_cut_finite_timeout_is_collective():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_timeout_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert outs[0][1:] == outs[1][1:], outs     # same status, result, node and wave counts on both ranks
    assert outs[0][1] == "timedout", outs        # 129 nodes at >= 15 ms per wave of 4 cannot finish in 60 ms
