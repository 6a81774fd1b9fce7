#!/usr/bin/env python
"""BASELINE.json config 3 at full size: 1 M RHS-perturbed (eps = 1e-2) replicas of Netlib SC105 / ADLITTLE, tableaus
generated and solved in HBM chunk by chunk, replicas sharded over the GPUs of one box with no data-path collective
(replica i -> rank floor(i*world/N)).

    python scripts/config3_sharded.py [SC105|ADLITTLE] [total_replicas]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/config3_sharded.py SC105

Prints one JSON line: LPs/s and pivots/s (max time over ranks, CUDA events around the solve launches), status
counts, and a checksum of the rounded objectives that is independent of the sharding (so runs at 1/2/4/8 GPUs can be
compared with each other: the replicas and their solutions are the same whatever the rank that solved them)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch

import yalps_b200
from yalps_b200 import distributed as D
from conftest import load_netlib

name = sys.argv[1] if len(sys.argv) > 1 else "SC105"
total = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
eng = yalps_b200.Engine(local)
g = load_netlib().get(name)
H, W = g["height"], g["width"]
cells = H * W
lo, hi = D.shard_range(total, rank, world)
stream = torch.cuda.current_stream().cuda_stream
d_in = torch.empty(chunk * cells, dtype=torch.float64, device=dev)
d_work = torch.empty(chunk * cells, dtype=torch.float64, device=dev)
d_status = torch.empty(chunk, dtype=torch.int32, device=dev)
d_value = torch.empty(chunk, dtype=torch.float64, device=dev)
d_piv = torch.empty(chunk, 2, dtype=torch.int64, device=dev)
counts = torch.zeros(5, dtype=torch.int64, device=dev)
pivots = torch.zeros(1, dtype=torch.int64, device=dev)
checksum = torch.zeros(1, dtype=torch.float64, device=dev)
vmin = torch.full((1,), float("inf"), dtype=torch.float64, device=dev)
vmax = torch.full((1,), float("-inf"), dtype=torch.float64, device=dev)
solve_ms = 0.0
gen_ms = 0.0
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
first = True
if dist is not None:
    dist.barrier()
torch.cuda.synchronize()
for begin in range(lo, hi, chunk):
    n = min(chunk, hi - begin)
    for rep in range(2 if first else 1):  # the first chunk once untimed (allocations, occupancy queries)
        e0.record()
        eng.generate_replicas_device(begin, n, g["matrix"], H, W, g["row_groups"], d_in.data_ptr(), stream=stream)
        d_work[: n * cells].copy_(d_in[: n * cells])
        e1.record()
        eng.solve_batch_device(n, H, W, d_in.data_ptr(), d_work=d_work.data_ptr(), d_status=d_status.data_ptr(),
                               d_value=d_value.data_ptr(), d_pivots=d_piv.data_ptr(), stream=stream)
        e2.record()
        torch.cuda.synchronize()
    first = False
    gen_ms += e0.elapsed_time(e1)
    solve_ms += e1.elapsed_time(e2)
    counts += torch.bincount(d_status[:n].to(torch.int64), minlength=5)
    pivots += d_piv[:n].sum()
    v = d_value[:n]
    ok = d_status[:n] == 0
    # order-independent checksum: sum of the rounded objectives scaled to integers (exact in float64 at this size)
    checksum += torch.round(v[ok] * 1e6).sum()
    vmin = torch.minimum(vmin, v[ok].min()) if bool(ok.any()) else vmin
    vmax = torch.maximum(vmax, v[ok].max()) if bool(ok.any()) else vmax
t = torch.tensor([solve_ms, gen_ms], dtype=torch.float64, device=dev)
if dist is not None:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(counts)
    dist.all_reduce(pivots)
    dist.all_reduce(checksum)
    dist.all_reduce(vmin, op=dist.ReduceOp.MIN)
    dist.all_reduce(vmax, op=dist.ReduceOp.MAX)
if rank == 0:
    s = float(t[0]) * 1e-3
    dense = 16 * W * H  # dense bytes per pivot
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    print(json.dumps({
        "workload": f"config3: {total} RHS-perturbed (eps=1e-2) replicas of Netlib {name} {H}x{W}, {world} GPU(s), chunks of {chunk}",
        "n_gpus": world, "solve_s": s, "generate_s": float(t[1]) * 1e-3, "lps_per_s": total / s,
        "pivots_per_s": int(pivots.item()) / s, "pivots_per_lp": int(pivots.item()) / total,
        "status_counts": counts.tolist(), "objective_checksum": float(checksum.item()),
        "value_range": [float(vmin.item()), float(vmax.item())], "base_value": g["value"],
        "dense_bytes_per_pivot": dense,
        "hbm_roofline_dense_frac": int(pivots.item()) * dense / s / 1e9 / (hbm * world),
        "note": "dense-equivalent bytes; the kernels skip untouched rows (|coef| <= 1e-16), so > 1 is possible"}))
eng.close()
if dist is not None:
    dist.destroy_process_group()
