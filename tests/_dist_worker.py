"""Worker for tests/test_dist.py: exercises bench.py's multi-rank plumbing on CPU (gloo)."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402

rank, world, local, barrier, allmax = bench.dist_setup(int(os.environ["WORLD_SIZE"]))
barrier()
slowest = allmax(1.0 + rank)              # max over ranks, as the timing rules require
barrier()
out = {"rank": rank, "world": world, "local": local, "max": slowest, "gpu_mask": 1 << local,
       "dir": str(bench.shm_base() / f"vmig_bench_r{rank}")}
Path(os.environ["VMIG_DIST_OUT"], f"r{rank}.json").write_text(json.dumps(out))
import torch.distributed as dist  # noqa: E402
dist.barrier()
dist.destroy_process_group()
