#!/usr/bin/env python
"""Branch and cut with the frontier sharded over the GPUs of one box (torchrun, one rank per GPU):
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/milp_sharded.py "<case>"
Every rank uploads the root, evaluates its share of each node wave, all-gathers results over NCCL and
min-allreduces the incumbent; rank 0 checks the result against the committed golden vectors."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist

import yalps_b200
from yalps_b200 import distributed as D, engine as E
from conftest import load_cases

name = sys.argv[1] if len(sys.argv) > 1 else "Large Farm MIP"
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
case = next(c for c in load_cases() if c["name"] == name)
opt = {"precision": 1e-8, "checkCycles": False, "maxPivots": 8192, "tolerance": 0, "timeout": float("inf"),
       "maxIterations": 32768, "includeZeroVariables": False, **case["options"]}
eng = yalps_b200.Engine(local)
tm = yalps_b200.tableau_model(case["model"])
t = tm.tableau
copt = E.make_options(opt["precision"], opt["maxPivots"], opt["checkCycles"], opt["tolerance"], opt["timeout"],
                      opt["maxIterations"])
root = eng.solve_batch(t.matrix, t.height, t.width, copt, want_matrices=True)
eng.bnb_set_root(root["matrices"][0], t.height, t.width, root["pos"][0], root["var"][0], 2 * len(tm.integers))
dist.barrier()
t0 = time.perf_counter()
res = D.branch_and_cut_sharded(D.engine_node_evaluator(eng, copt), root["rhs"][0], root["pos"][0], root["var"][0],
                               t.width, t.height, tm.integers, tm.sign, float(root["value"][0]), opt, wave=64,
                               device=torch.device("cuda", local))
dt = time.perf_counter() - t0
o = case["oracle"]
ok = (res["status"] == o["status"] and res["stats"]["nodes"] == o["nodes"]
      and res["stats"]["node_pivots"] == o["node_pivots"] and np.array_equal(res["pos"], o["final_pos"]))
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"{name}: {res['status']} {-tm.sign * res['result']} nodes={res['stats']['nodes']} waves={res['stats']['waves']} "
          f"device_nodes={res['stats']['device_nodes']} allreduces={res['stats']['allreduces']} world={world} {dt * 1e3:.1f} ms")
    print("PARITY OK" if int(flag.item()) == 1 else "PARITY FAILED")
eng.close()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
