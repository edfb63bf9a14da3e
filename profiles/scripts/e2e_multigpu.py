"""BASELINE config 3 in miniature: ONE migration whose block list is sharded across the GPUs of the box by
libvmig itself (vmig_opts.gpu_mask), all inside one process -- the way the Go control plane would use it.
usage: python profiles/scripts/e2e_multigpu.py [n_files=20] [file_gib=1] [gpu_counts=1,2,4,8] [reps=3]"""
import shutil, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import __graft_entry__ as g

vm = g.load_pkg()
n_files = int(sys.argv[1]) if len(sys.argv) > 1 else 20
gib = float(sys.argv[2]) if len(sys.argv) > 2 else 1
base = Path("/dev/shm/vmig_mgpu"); shutil.rmtree(base, ignore_errors=True); base.mkdir()
vm.init(0)
ndev = vm.device_count()
vm.datagen_files(base / "src", 3, n_files, int(gib * (1 << 30)), threads=32)
nbytes = n_files * int(gib * (1 << 30))
tables = {}
counts = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2, 4, 8]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
for ngpu in counts:
    if ngpu > ndev: break
    mask = (1 << ngpu) - 1
    for rep in range(reps):
        shutil.rmtree(base / "dst", ignore_errors=True); (base / "dst").mkdir()
        t0 = time.perf_counter(); st = vm.migrate_tree(base / "src", base / "dst", None, base / f"t{ngpu}.vmig", gpu_mask=mask); dt = time.perf_counter() - t0
        print(f"gpu_mask=0x{mask:02x} ({st['gpus_used']} GPUs) rep {rep}: {nbytes / dt / (1 << 30):6.2f} GiB/s  ({dt * 1e3:.0f} ms, kernel sum {st['ms_kernel']:.0f} ms, {st['kernel_launches']} launches)", flush=True)
    tables[ngpu] = (base / f"t{ngpu}.vmig").read_bytes()
assert all(t == tables[counts[0]] for t in tables.values()), "block tables differ between GPU counts"
print("block tables identical for every GPU count")
shutil.rmtree(base, ignore_errors=True)
