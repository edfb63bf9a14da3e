"""BASELINE config 4 shape at 10 GiB: the destination already holds the prior version; 30 % of the blocks
changed.  Only those travel back over PCIe and get written.
usage: python profiles/scripts/e2e_rollback.py [n_files=10] [changed_percent=30]"""
import shutil, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np
import __graft_entry__ as g

vm = g.load_pkg()
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 10
pct = int(sys.argv[2]) if len(sys.argv) > 2 else 30
bb, fb = 4 << 20, 1 << 30
base = Path("/dev/shm/vmig_rb"); shutil.rmtree(base, ignore_errors=True); base.mkdir(); (base / "dst").mkdir()
vm.init(1)
vm.datagen_files(base / "src", 4, nf, fb, threads=32)
vm.migrate_tree(base / "src", base / "dst", None, base / "v1.vmig")
nblk = nf * fb // bb
changed = np.sort(np.random.default_rng(44).permutation(nblk)[: nblk * pct // 100])
def mutate():
    for gidx in changed:
        with open(base / "src" / f"f{gidx // 256:05d}.bin", "r+b") as f:
            f.seek((gidx % 256) * bb); w = bytes(x ^ 0xFF for x in f.read(8)); f.seek((gidx % 256) * bb); f.write(w)
prior = base / "v1.vmig"
for rep in range(4):
    mutate()                                   # src differs from what dst holds in exactly `changed`
    out = base / f"v{rep + 2}.vmig"
    t0 = time.perf_counter(); st = vm.migrate_tree(base / "src", base / "dst", prior, out); dt = time.perf_counter() - t0
    print(f"rep {rep}: {nf * fb / dt / (1 << 30):6.2f} GiB/s logical ({dt * 1e3:.0f} ms); written {st['blocks_total'] - st['blocks_skipped']}/{st['blocks_total']} blocks, "
          f"D2H {st['bytes_d2h'] / (1 << 30):.2f} GiB of {st['bytes_h2d'] / (1 << 30):.2f} GiB H2D", flush=True)
    prior = out
shutil.rmtree(base, ignore_errors=True)
