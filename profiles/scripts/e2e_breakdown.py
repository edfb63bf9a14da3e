"""Where does end-to-end time go?  Same 10 GiB through progressively more of the host path.
usage: python profiles/scripts/e2e_breakdown.py [gib=10]"""
import os, shutil, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np
import __graft_entry__ as g

vm = g.load_pkg()
gib = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n = gib << 30
vm.init(1)

def rate(label, fn, reps=3):
    best = 0
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); dt = time.perf_counter() - t0
        best = max(best, n / dt / (1 << 30))
    print(f"{label:70s} {best:7.2f} GiB/s (best of {reps})", flush=True)

a, b = vm.PinnedBuffer(n), vm.PinnedBuffer(n)
a.array[:] = 7
rate("pinned buffer -> pinned buffer (DMA direct, no host copies)", lambda: vm.migrate_buffer(a.array, b.array))
rate("pinned buffer -> hash only (H2D + K1)", lambda: vm.migrate_buffer(a.array, None, flags=vm.F_HASH_ONLY))
b.free()
pg = np.empty(n, np.uint8); pg[:] = 1
rate("pinned src -> pageable dst (writers memcpy out of the OUT ring)", lambda: vm.migrate_buffer(a.array, pg))
rate("pageable src -> hash only (readers memcpy into the IN ring)", lambda: vm.migrate_buffer(pg, None, flags=vm.F_HASH_ONLY))
a.free(); del pg
base = Path("/dev/shm/vmig_breakdown"); shutil.rmtree(base, ignore_errors=True); base.mkdir()
vm.datagen_files(base / "src", 2, gib, 1 << 30, threads=32)
rate("tmpfs files -> hash only (pread + H2D + K1)", lambda: vm.hash_tree(base / "src", base / "t.vmig"))
def full():
    shutil.rmtree(base / "dst", ignore_errors=True); (base / "dst").mkdir()
    vm.migrate_tree(base / "src", base / "dst")
rate("tmpfs files -> tmpfs files (full path, includes rm -rf of dst)", full)
shutil.rmtree(base, ignore_errors=True)
