"""bench.py's e2e leg alone, for N concurrent ranks (one process per GPU, one 10 x 1 GiB tree each) without
torchrun: used to study how the host side scales.  Environment knobs (VMIG_*) pass through to every rank.
usage: python profiles/scripts/e2e_ranks.py N [steps=3] [n_files=10] [mode=copy]
mode: copy (vmig_migrate_tree) | hash (VMIG_F_HASH_ONLY: read side + H2D only) | buffer (pinned -> pinned,
no file I/O) | tar (the reference's pipe, no GPU) | mount (copy, each rank on its own tmpfs mount)
VMIG_SWEEP="8:12,5:8,3:5" repeats the timed steps for each readers:writers pair on the same trees."""
import multiprocessing as mp, os, shutil, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]

def rank_main(rank, world, steps, n_files, bar, out, mode):
    sys.path.insert(0, str(ROOT))
    import __graft_entry__ as g
    vm = g.load_pkg()
    os.environ.setdefault("VMIG_IO_SHARE", str(world))
    vm.init(1 << rank)
    import subprocess
    base = Path(f"/dev/shm/vmig_ranks_r{rank}")
    if mode == "mount":
        base = Path(f"/tmp/vmig_mnt_r{rank}"); base.mkdir(exist_ok=True)
        subprocess.run(["mount", "-t", "tmpfs", "-o", "size=40g", "tmpfs", str(base)], check=True)
    else:
        shutil.rmtree(base, ignore_errors=True); base.mkdir()
    times = []
    if mode == "buffer":
        a, b = vm.PinnedBuffer(n_files << 30), vm.PinnedBuffer(n_files << 30)
        a.array[:] = 7; b.array[:] = 0
    else:
        vm.datagen_files(base / "src", 2 + 1000 * rank, n_files, 1 << 30, threads=max(4, 64 // world))
    sweep = [x for x in os.environ.get("VMIG_SWEEP", "").split(",") if x] or [None]
    for i in range((steps + 1) * len(sweep)):
        cfg = sweep[i // (steps + 1)]
        if cfg: os.environ["VMIG_READERS"], os.environ["VMIG_WRITERS"] = cfg.split(":")
        shutil.rmtree(base / "dst", ignore_errors=True); (base / "dst").mkdir()
        bar.wait()
        t0 = time.perf_counter()
        if mode in ("copy", "mount"): vm.migrate_tree(base / "src", base / "dst", None, None, gpu_mask=1 << rank)
        elif mode == "hash": vm.hash_tree(base / "src", base / "t.vmig", gpu_mask=1 << rank)
        elif mode == "buffer": vm.migrate_buffer(a.array, b.array, gpu_mask=1 << rank)
        elif mode == "tar": subprocess.run(["sh", "-c", f"(cd {base}/src; tar c .) | (cd {base}/dst; tar x)"], check=True)
        dt = time.perf_counter() - t0
        bar.wait()
        if i % (steps + 1): times.append(dt)
    out.put((rank, times))
    if mode == "mount": subprocess.run(["umount", str(base)])
    else: shutil.rmtree(base, ignore_errors=True)

if __name__ == "__main__":
    world = int(sys.argv[1]); steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    n_files = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    mode = sys.argv[4] if len(sys.argv) > 4 else "copy"
    ctx = mp.get_context("spawn")
    bar, out = ctx.Barrier(world), ctx.Queue()
    ps = [ctx.Process(target=rank_main, args=(r, world, steps, n_files, bar, out, mode)) for r in range(world)]
    [p.start() for p in ps]
    res = dict(out.get() for _ in ps)
    [p.join() for p in ps]
    knobs = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("VMIG_") and k != "VMIG_SWEEP")
    sweep = [x for x in os.environ.get("VMIG_SWEEP", "").split(",") if x] or [None]
    for ci, cfg in enumerate(sweep):
        sl = slice(ci * steps, (ci + 1) * steps)
        worst = [max(res[r][sl][i] for r in res) for i in range(steps)]
        print(f"N={world} {mode} [{knobs}{' R:W=' + cfg if cfg else ''}] per-step max-over-ranks ms: {[round(1e3 * t) for t in worst]}  -> "
              f"{world * n_files * steps / sum(worst):.2f} GiB/s total; per-rank mean ms: "
              f"{[round(1e3 * sum(res[r][sl]) / steps) for r in sorted(res)]}", flush=True)
