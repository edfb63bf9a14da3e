"""Small ragged K1 run for compute-sanitizer (memcheck / racecheck)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np
import __graft_entry__ as g
from oracle import oracle as orc
vm = g.load_pkg(); vm.init(1)
rng = np.random.default_rng(5)
lens = np.array([0, 1, 31, 32, 33, 1023, 1024, 1025, 2048, 4097, 65536, 100001, 300000] + list(rng.integers(0, 200000, 40)), dtype=np.uint32)
offs = np.concatenate([[0], np.cumsum(lens[:-1].astype(np.uint64) + 5)]).astype(np.uint64)
buf = rng.integers(0, 256, int(offs[-1] + lens[-1]) + 8, dtype=np.uint8)
got, ms = vm.hash_blocks(buf, offs, lens)
assert (got == orc.hash_blocks(buf, offs, lens)).all()
print("ok", len(lens), "blocks", ms, "ms")
