"""Round-2 end-to-end sweeps: ONE vmig_migrate_tree call per rep on an n_files x 1 GiB tmpfs tree, each variant in a fresh
process (the staging rings are created once per process, so ring placement / huge pages / thread knobs need one).
  python profiles/scripts/r2_e2e_sweep.py N_FILES REPS GPUS VARIANT[,VARIANT...]
VARIANT = name understood below, optionally suffixed :L<lanes_per_gpu>, e.g.  base  outfar  split:L2  r6w10
internal: ... run N_FILES REPS GPUS LANES (env already set)"""
import os, shutil, subprocess, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
GiB = 1 << 30
BASE = Path(os.environ.get("VMIG_BENCH_DIR", "/dev/shm")) / "vmig_sweep"
VARIANTS = {
    "base": {},
    "trace": {"VMIG_TRACE": "1"},
    "outfar": {"VMIG_RING_OUT_NODE": "1"},                       # OUT ring on the writers' socket
    "outsplit": {"VMIG_RING_OUT_NODE": "2"},
    "split": {"VMIG_RING_IN_NODE": "2", "VMIG_RING_OUT_NODE": "2"},
    "split_unbound": {"VMIG_RING_IN_NODE": "2", "VMIG_RING_OUT_NODE": "2", "VMIG_BIND_IO": "0"},
    "insplit_outfar": {"VMIG_RING_IN_NODE": "2", "VMIG_RING_OUT_NODE": "1"},
    "outfar_wlocal": {"VMIG_RING_OUT_NODE": "1", "VMIG_BIND_WRITERS": "1"},
    "wlocal": {"VMIG_BIND_WRITERS": "1"},
    "unbound": {"VMIG_BIND_IO": "0"},
    "r4w6": {"VMIG_READERS": "4", "VMIG_WRITERS": "6"}, "r5w8": {"VMIG_READERS": "5", "VMIG_WRITERS": "8"},
    "r6w10": {"VMIG_READERS": "6", "VMIG_WRITERS": "10"}, "r8w10": {"VMIG_READERS": "8", "VMIG_WRITERS": "10"},
    "r3w5": {"VMIG_READERS": "3", "VMIG_WRITERS": "5"}, "r2w3": {"VMIG_READERS": "2", "VMIG_WRITERS": "3"},
    "r8w14": {"VMIG_READERS": "8", "VMIG_WRITERS": "14"}, "r9w12": {"VMIG_READERS": "9", "VMIG_WRITERS": "12"},
    "r7w12": {"VMIG_READERS": "7", "VMIG_WRITERS": "12"}, "r8w16": {"VMIG_READERS": "8", "VMIG_WRITERS": "16"},
    "r10w16": {"VMIG_READERS": "10", "VMIG_WRITERS": "16"}, "r12w20": {"VMIG_READERS": "12", "VMIG_WRITERS": "20"},
    "outfar_r10w16": {"VMIG_RING_OUT_NODE": "1", "VMIG_READERS": "10", "VMIG_WRITERS": "16"},
    "direct": {"VMIG_DIRECT_IO": "1"}, "cufile": {"VMIG_CUFILE": "1"}, "direct_cufile": {"VMIG_DIRECT_IO": "1", "VMIG_CUFILE": "1"},
    "outnear": {"VMIG_RING_OUT_NODE": "0"}, "wboth": {"VMIG_BIND_WRITERS": "0"},
    "slots8": {"VMIG_SLOTS": "8"}, "slot16mb": {"VMIG_SLOT_MB": "16", "VMIG_SLOTS": "32"}, "slot8mb": {"VMIG_SLOT_MB": "8", "VMIG_SLOTS": "32"},
}


def run(n_files, reps, gpus, lanes):
    import __graft_entry__ as g
    vm = g.load_pkg()
    mask = (1 << gpus) - 1
    vm.init(mask)
    for rep in range(reps + 1):
        dst = BASE / "dst"; shutil.rmtree(dst, ignore_errors=True); dst.mkdir()
        t0 = time.perf_counter()
        st = vm.migrate_tree(BASE / "src", dst, None, None, gpu_mask=mask, lanes_per_gpu=lanes)
        dt = time.perf_counter() - t0
        print(f"    rep {rep}{' (creates the rings)' if rep == 0 else ''}: {n_files * GiB / dt / GiB:6.2f} GiB/s ({dt * 1e3:.0f} ms) lanes={st['lanes_used']} gpus={st['gpus_used']}", flush=True)
    shutil.rmtree(BASE / "dst", ignore_errors=True)


def main():
    if sys.argv[1] == "run":
        return run(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]))
    n_files, reps, gpus = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    shutil.rmtree(BASE, ignore_errors=True); BASE.mkdir(parents=True)
    subprocess.run([str(ROOT / "tools" / "vmig_datagen"), "files", str(BASE / "src"), "3", str(n_files), str(GiB), "32"], check=True, capture_output=True)
    print(f"tree: {n_files} x 1 GiB under {BASE}/src; gpu_mask 0x{(1 << gpus) - 1:x}", flush=True)
    try:
        for v in sys.argv[4].split(","):
            name, _, l = v.partition(":L")
            env = dict(os.environ); env.update(VARIANTS[name])
            print(f"[{v}] {VARIANTS[name]}", flush=True)
            subprocess.run([sys.executable, __file__, "run", str(n_files), str(reps), str(gpus), l or "0"], env=env, check=False)
    finally:
        shutil.rmtree(BASE, ignore_errors=True)


if __name__ == "__main__":
    main()
