"""CPU, world_size 2, gloo: the N>1 path of bench.py (spec ⑤).  The data path shards by
independent trees / blocks with NO collective; what ranks share is only the barrier and the
max-over-ranks of the timed region, which is what this covers."""
import json
import os
import socket
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_barrier_and_max(tmp_path):
    env = dict(os.environ, VMIG_DIST_OUT=str(tmp_path), CUDA_VISIBLE_DEVICES="")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(ROOT / "tests" / "_dist_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    outs = [json.loads((tmp_path / f"r{i}.json").read_text()) for i in range(2)]
    assert [o["rank"] for o in outs] == [0, 1] and all(o["world"] == 2 for o in outs)
    assert all(o["max"] == 2.0 for o in outs)                    # both ranks see the slowest rank's time
    assert outs[0]["gpu_mask"] == 1 and outs[1]["gpu_mask"] == 2  # one GPU per rank
    assert outs[0]["dir"] != outs[1]["dir"] and outs[0]["seed"] != outs[1]["seed"]   # one tree per rank (weak scaling)
    assert all(o["active_2A"] for o in outs)
    # single-call configs (sharded call, caller threads): rank 0 drives every GPU of the box from ONE process
    assert all(outs[0]["active_solo"].values()) and not any(outs[1]["active_solo"].values())
    assert outs[0]["all_mask"] == 0b11
    assert [o["agreed"] for o in outs] == ["decided-by-rank-0"] * 2      # every rank follows rank 0's decision


def test_reference_arm_skips_on_nonzero_rank():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2"], env=env,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
