"""Multi-GPU host layer: one process per GPU (torch.distributed), LPs and branch-and-bound frontier nodes
sharded across ranks.

* Independent LPs shard with NO data-path collective: LP i belongs to rank floor(i*G/N); every rank solves its
  contiguous range on its own GPU and the host gathers the (small) results.
* Branch and cut shards per wave: the root tableau is replicated on every GPU, the nodes of a wave are dealt
  round-robin, results are all-gathered, and every rank replays the reference loop (src/branchAndCut.ts:122-164)
  on the same data, so all ranks hold the same heap and incumbent.  Every `allreduce_every` waves the incumbent
  objective goes through a min-allreduce (NCCL over NVLink on GPUs, gloo in the CPU tests) as north_star
  asks; with the replicated replay it doubles as a consistency check.

The node evaluator is injected (`eval_nodes`), so the host logic is testable on CPU with world_size 2.
"""
from __future__ import annotations

import heapq
import math
import time
from typing import Callable, Optional, Sequence

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple:
    """Contiguous index range of `rank`: item i -> rank floor(i*world/n)."""
    lo = -(-rank * n // world)        # ceil(rank*n/world)
    hi = -(-(rank + 1) * n // world)
    return lo, hi


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def allreduce_min(value: float, device=None) -> float:
    """Min-allreduce of one fp64 (the incumbent objective; lower is better internally, src/branchAndCut.ts:124)."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return value
    import torch
    t = torch.tensor([value if not math.isnan(value) else math.inf], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return float(t.item())


def allreduce_flag(flag: bool, device=None) -> bool:
    """Collective OR of a stop flag: every rank gets True as soon as one rank raises it."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return bool(flag)
    import torch
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return bool(int(t.item()))


def gather_arrays(local: np.ndarray, counts: Sequence[int], device=None) -> np.ndarray:
    """all_gather of per-rank row blocks with known row counts -> concatenated array on every rank."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return local
    import torch
    width = int(np.prod(local.shape[1:])) if local.ndim > 1 else 1
    cap = max(counts)
    buf = torch.zeros(cap * width, dtype=torch.from_numpy(np.zeros(1, local.dtype)).dtype, device=device or "cpu")
    if local.size:
        buf[:local.size] = torch.from_numpy(np.ascontiguousarray(local).reshape(-1)).to(buf.device)
    parts = [torch.empty_like(buf) for _ in counts]
    dist.all_gather(parts, buf)
    out = [p[:c * width].cpu().numpy().reshape((c,) + local.shape[1:]) for p, c in zip(parts, counts)]
    return np.concatenate(out, axis=0)


def solve_batch_sharded(solve_local: Callable[[int, int], dict], n: int, device=None) -> dict:
    """Shards n independent LPs over the ranks.  solve_local(lo, hi) solves LPs [lo, hi) on this rank's GPU
    and returns {"status", "value", "pivots", ...} arrays with hi-lo rows; the result holds all n rows on
    every rank.  No collective touches the tableaus."""
    dist = _dist()
    world = dist.get_world_size() if dist else 1
    rank = dist.get_rank() if dist else 0
    lo, hi = shard_range(n, rank, world)
    local = solve_local(lo, hi)
    counts = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
    return {k: gather_arrays(np.asarray(v), counts, device) for k, v in local.items() if v is not None}


# ------------------------------------------------------------------------------------------ branch and cut

class _Branch:
    """Heap entry ordered by `eval` only, like the reference comparator x[0]-y[0] (src/branchAndCut.ts:100);
    heapq is the algorithm npm heap@0.2.7 ports."""

    __slots__ = ("eval", "cuts", "id")

    def __init__(self, ev, cuts, bid):
        self.eval, self.cuts, self.id = ev, cuts, bid

    def __lt__(self, other):
        return self.eval - other.eval < 0


def _js_round(x: float) -> float:
    if math.isnan(x) or math.isinf(x) or abs(x) >= 2.0 ** 52:
        return x
    r = float(math.floor(x))
    if x - r >= 0.5:
        r += 1.0
    return r


def _most_fractional(rhs, pos, width, integers):
    """src/branchAndCut.ts:64-85"""
    best, var, val = 0.0, 0, 0.0
    for iv in integers:
        row = int(pos[iv]) - width
        if row < 0:
            continue
        v = float(rhs[row])
        frac = abs(v - _js_round(v))
        if frac > best:
            best, var, val = frac, iv, v
    return var, val, best


def branch_and_cut_sharded(eval_nodes: Callable[[list], list], root_rhs, root_pos, root_var, width: int, height: int,
                           integers: Sequence[int], sign: float, init_result: float, options: dict, wave: int = 64,
                           allreduce_every: int = 4, device=None) -> dict:
    """branchAndCut (src/branchAndCut.ts:89-176) with the node LPs of each wave sharded over the ranks.

    eval_nodes(list_of_cut_lists) -> list of {"status", "value", "pivots", "rhs", "pos", "var"} evaluated on
    THIS rank's GPU (e.g. Engine.bnb_solve_nodes against the replicated root).  Every rank calls this function
    with the same arguments and returns the same result.

    `timeout` (src/branchAndCut.ts:115-116,162) is a wall-clock test, and the ranks' clocks differ: a rank that
    decided on its own could leave the loop while the others enter the next wave's collectives.  With more than one
    rank the decision is therefore collective: the clock is read where all ranks arrive in lockstep -- once before
    the loop and at every wave boundary -- and OR-reduced; between waves the (sticky) agreed flag is used.  With one
    rank, or with the default `timeout = inf` (no clock test needed at all), the loop is the reference's.
    """
    dist = _dist()
    world = dist.get_world_size() if dist else 1
    rank = dist.get_rank() if dist else 0
    precision = options["precision"]
    stats = {"nodes": 0, "node_pivots": 0, "waves": 0, "device_nodes": 0, "allreduces": 0}

    var0, val0, frac0 = _most_fractional(root_rhs, root_pos, width, integers)
    if frac0 <= precision:
        return {"status": "optimal", "result": init_result, "height": height, "rhs": np.asarray(root_rhs),
                "pos": np.asarray(root_pos), "var": np.asarray(root_var), "stats": stats}

    next_id = 0
    heap: list = []
    for cut in ((-1.0, var0, float(math.ceil(val0))), (1.0, var0, float(math.floor(val0)))):
        heapq.heappush(heap, _Branch(init_result, [cut], next_id))
        next_id += 1

    cache: dict = {}
    threshold = init_result * (1.0 - sign * options["tolerance"])
    stop_time = options["timeout"] + math.floor(time.time() * 1000.0)
    collective_clock = world > 1 and math.isfinite(options["timeout"])

    def clock_says_stop() -> bool:
        return math.floor(time.time() * 1000.0) >= stop_time

    timedout = allreduce_flag(clock_says_stop(), device) if collective_clock else clock_says_stop()
    found, best_eval, best = False, math.inf, None
    it = 0

    def run_wave(needed):
        peek = list(heap)  # a heap; popping from the copy gives the reference's upcoming pop order
        batch = [needed]
        while len(batch) < wave and peek:
            b = heapq.heappop(peek)
            if b.eval > best_eval:
                break
            if b.id not in cache:
                batch.append(b)
        mine = [i for i in range(len(batch)) if i % world == rank]
        local = eval_nodes([batch[i].cuts for i in mine]) if mine else []
        stats["waves"] += 1
        stats["device_nodes"] += len(batch)
        if world == 1:
            for i, r in zip(mine, local):
                cache[batch[i].id] = r
            return
        # all-gather the wave: fixed-size records (status, value, pivots, height, rhs, pos, var)
        hcap = height + max(len(b.cuts) for b in batch)
        rec = 4 + hcap + 2 * (width + hcap)
        counts = [len(range(r, len(batch), world)) for r in range(world)]
        block = np.zeros((len(mine), rec), np.float64)
        for k, r in enumerate(local):
            h = len(r["rhs"])
            block[k, 0:4] = (r["status"], r["value"], r["pivots"], h)
            block[k, 4:4 + h] = r["rhs"]
            block[k, 4 + hcap:4 + hcap + width + h] = r["pos"]
            block[k, 4 + hcap + width + hcap:4 + hcap + width + hcap + width + h] = r["var"]
        allrec = gather_arrays(block, counts, device)
        offs = np.concatenate([[0], np.cumsum(counts)])
        for r_ in range(world):
            for k, i in enumerate(range(r_, len(batch), world)):
                row = allrec[offs[r_] + k]
                h = int(row[3])
                cache[batch[i].id] = {
                    "status": int(row[0]), "value": float(row[1]), "pivots": int(row[2]),
                    "rhs": row[4:4 + h].copy(),
                    "pos": row[4 + hcap:4 + hcap + width + h].astype(np.int32),
                    "var": row[4 + hcap + width + hcap:4 + hcap + width + hcap + width + h].astype(np.int32),
                }

    while it < options["maxIterations"] and heap and best_eval >= threshold and not timedout:
        br = heapq.heappop(heap)
        if br.eval > best_eval:
            break
        if br.id not in cache:
            if collective_clock and allreduce_flag(clock_says_stop(), device):
                heapq.heappush(heap, br)  # same heap contents on every rank: the status rule below sees it non-empty
                timedout = True
                break
            run_wave(br)
            if allreduce_every and stats["waves"] % allreduce_every == 0:
                agreed = allreduce_min(best_eval, device)
                stats["allreduces"] += 1
                if agreed != best_eval and not (math.isinf(agreed) and math.isinf(best_eval)):
                    raise RuntimeError(f"rank {rank}: incumbent {best_eval} disagrees with the allreduced {agreed}")
        node = cache.pop(br.id)
        stats["nodes"] += 1
        stats["node_pivots"] += node["pivots"]
        if node["status"] == 0 and node["value"] < best_eval:
            var, val, frac = _most_fractional(node["rhs"], node["pos"], width, integers)
            if frac <= precision:
                found, best_eval, best = True, node["value"], node
            else:
                upper, lower = [], []
                for cut in br.cuts:
                    if cut[1] == var:
                        (lower if cut[0] < 0 else upper).append(cut)
                    else:
                        upper.append(cut)
                        lower.append(cut)
                lower.append((1.0, var, float(math.floor(val))))
                upper.append((-1.0, var, float(math.ceil(val))))
                heapq.heappush(heap, _Branch(node["value"], upper, next_id))
                heapq.heappush(heap, _Branch(node["value"], lower, next_id + 1))
                next_id += 2
        if not collective_clock:
            timedout = clock_says_stop()
        it += 1

    unfinished = (timedout or it >= options["maxIterations"]) and bool(heap) and best_eval >= threshold
    status = "timedout" if unfinished else ("optimal" if found else "infeasible")
    if found:
        return {"status": status, "result": best_eval, "height": len(best["rhs"]), "rhs": best["rhs"],
                "pos": best["pos"], "var": best["var"], "stats": stats}
    return {"status": status, "result": math.nan, "height": height, "rhs": np.asarray(root_rhs),
            "pos": np.asarray(root_pos), "var": np.asarray(root_var), "stats": stats}


def engine_node_evaluator(engine, options_struct) -> Callable[[list], list]:
    """eval_nodes backed by Engine.bnb_solve_nodes (root already uploaded with Engine.bnb_set_root)."""

    def eval_nodes(cut_lists):
        out = engine.bnb_solve_nodes(cut_lists, options_struct)
        H, W = engine._root_shape
        res = []
        for j, cuts in enumerate(cut_lists):
            h = H + len(cuts)
            res.append({"status": int(out["status"][j]), "value": float(out["value"][j]),
                        "pivots": int(out["pivots"][j].sum()), "rhs": out["rhs"][j][:h].copy(),
                        "pos": out["pos"][j][:W + h].copy(), "var": out["var"][j][:W + h].copy()})
        return res

    return eval_nodes
