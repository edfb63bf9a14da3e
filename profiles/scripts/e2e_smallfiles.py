"""BASELINE config 2B: a realistic diff layer -- many files, log-uniform 1 KiB..64 MiB, depth-4 tree, some
symlinks / empty files / one hard-link pair -- migrated by libvmig and by the reference tar pipe.
usage: python profiles/scripts/e2e_smallfiles.py [total_gib=10] [n_files=40960]"""
import os, shutil, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np
import __graft_entry__ as g
from oracle import oracle as orc      # profiling script: the reference side of the comparison

vm = g.load_pkg()
total = float(sys.argv[1]) if len(sys.argv) > 1 else 10
nfiles = int(sys.argv[2]) if len(sys.argv) > 2 else 40960
base = Path("/dev/shm/vmig_small"); shutil.rmtree(base, ignore_errors=True); (base / "src").mkdir(parents=True)
rng = np.random.default_rng(2)
sizes = np.exp(rng.uniform(np.log(1024), np.log(64 << 20), nfiles))
sizes = (sizes * (total * (1 << 30)) / sizes.sum()).astype(np.int64)
sizes[rng.integers(0, nfiles, nfiles // 200)] = 0
pool = orc.splitmix_bytes(2, 128 << 20)
t0 = time.perf_counter()
for i, sz in enumerate(sizes):
    d = base / "src" / f"a{i % 8}" / f"b{(i // 8) % 8}" / f"c{(i // 64) % 8}"
    if i < 512: d.mkdir(parents=True, exist_ok=True)
    p = d / f"f{i}.bin"
    with open(p, "wb") as f:
        off = int(rng.integers(0, (128 << 20) - min(sz, 64 << 20) - 1)) if sz else 0
        f.write(pool[off:off + sz].data)
    if i % 100 == 7: os.symlink(p.name, d / f"l{i}")
os.link(base / "src/a0/b0/c0/f0.bin", base / "src/a1/hard")
nbytes = int(sizes.sum())
print(f"generated {nfiles} files, {nbytes / (1 << 30):.2f} GiB in {time.perf_counter() - t0:.1f} s", flush=True)
vm.init(1)
for rep in range(3):
    shutil.rmtree(base / "dst", ignore_errors=True); (base / "dst").mkdir()
    t0 = time.perf_counter(); st = vm.migrate_tree(base / "src", base / "dst", None, base / "t.vmig"); dt = time.perf_counter() - t0
    print(f"vmig rep {rep}: {nbytes / dt / (1 << 30):.2f} GiB/s ({dt * 1e3:.0f} ms) walk {st['ns_walk'] / 1e6:.0f} plan {st['ns_plan'] / 1e6:.0f} data {st['ns_data'] / 1e6:.0f} meta {st['ns_meta'] / 1e6:.0f} table {st['ns_table'] / 1e6:.0f} ms; launches {st['kernel_launches']}", flush=True)
(base / "ref").mkdir()
t0 = time.perf_counter(); orc.ref_copy(base / "src", base / "ref"); dt = time.perf_counter() - t0
print(f"reference tar|tar: {nbytes / dt / (1 << 30):.2f} GiB/s ({dt * 1e3:.0f} ms)", flush=True)
t0 = time.perf_counter(); diffs = orc.compare_trees(base / "ref", base / "dst", content=False, mtime_ns=True); print("metadata diffs:", len(diffs), diffs[:3], f"({time.perf_counter() - t0:.1f} s)")
shutil.rmtree(base, ignore_errors=True)
