"""Worker for tests/test_dist.py: exercises bench.py's multi-rank plumbing on CPU (gloo)."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402

rank, world, local, barrier, allmax = bench.dist_setup()
barrier()
slowest = allmax(1.0 + rank)              # max over ranks, as the timing rules require
barrier()
bench.cpu_barrier()                      # the CPU-side (gloo) wait used while rank 0 drives all GPUs alone
agreed = bench.bcast_from_rank0(f"decided-by-rank-{rank}")      # rank 0's decision reaches every rank (sharded leg yes/no)
plan = bench.rank_plan("2A", local, world, world)
solo = {c: bench.rank_plan(c, local, world, world) for c in ("1", "2B", "3", "4", "5")}
out = {"rank": rank, "world": world, "local": local, "max": slowest, "gpu_mask": plan["own_mask"],
       "dir": str(bench.shm_base() / plan["tree"]), "seed": plan["seed"], "active_2A": plan["active"],
       "active_solo": {c: p["active"] for c, p in solo.items()}, "all_mask": solo["3"]["all_mask"], "agreed": agreed}
Path(os.environ["VMIG_DIST_OUT"], f"r{rank}.json").write_text(json.dumps(out))
import torch.distributed as dist  # noqa: E402
dist.barrier()
dist.destroy_process_group()
