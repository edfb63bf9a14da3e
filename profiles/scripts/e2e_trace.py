"""End-to-end migrate_tree on a synthetic tmpfs tree with the engine's stage trace on (VMIG_TRACE=1).
usage: python profiles/scripts/e2e_trace.py [n_files=10] [file_gib=1] [reps=3]   (env: VMIG_READERS, VMIG_WRITERS, ...)"""
import os
import shutil
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
os.environ.setdefault("VMIG_TRACE", "1")
import __graft_entry__ as g

vm = g.load_pkg()
n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 10
file_gib = float(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
base = Path("/dev/shm/vmig_e2e_trace")
shutil.rmtree(base, ignore_errors=True)
base.mkdir()
vm.init(1)
vm.datagen_files(base / "src", 2, n_files, int(file_gib * (1 << 30)), threads=32)
nbytes = n_files * int(file_gib * (1 << 30))
for i in range(reps):
    shutil.rmtree(base / "dst", ignore_errors=True)
    (base / "dst").mkdir()
    t0 = time.perf_counter()
    st = vm.migrate_tree(base / "src", base / "dst", None, base / "t.vmig")
    dt = time.perf_counter() - t0
    print(f"rep {i}: {nbytes / dt / (1 << 30):.2f} GiB/s ({dt * 1e3:.0f} ms) R={os.environ.get('VMIG_READERS','auto')} W={os.environ.get('VMIG_WRITERS','auto')} "
          f"slots={os.environ.get('VMIG_SLOTS','16')}x{os.environ.get('VMIG_SLOT_MB','32')}MiB", flush=True)
shutil.rmtree(base, ignore_errors=True)
