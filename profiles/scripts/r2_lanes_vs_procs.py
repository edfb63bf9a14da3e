"""Round-2 diagnosis of VERDICT r1 item 1: is the in-process multi-lane slowdown a per-process effect?
K lanes inside ONE process (vmig_opts.lanes_per_gpu / gpu_mask) against K single-lane PROCESSES, same trees,
same thread counts per lane, same GPU set.  Runs on a 1-GPU box (lanes share the GPU) or an N-GPU box.

  python profiles/scripts/r2_lanes_vs_procs.py [K=4] [files_per_lane=10] [reps=3] [gpus=1]
internal roles:  ... inproc <K> <gpus> | worker <i> <gpu>
"""
import os, shutil, subprocess, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
GiB = 1 << 30
BASE = Path(os.environ.get("VMIG_BENCH_DIR", "/dev/shm")) / "vmig_lvp"


def role_inproc(k, gpus, fpl, reps):
    import __graft_entry__ as g
    vm = g.load_pkg()
    mask = (1 << gpus) - 1
    vm.init(mask)
    # a tree holding exactly k of the subtrees (hard links: no extra memory)
    src = BASE / f"src_first{k}"
    if not src.exists():
        for i in range(k):
            (src / f"s{i}").mkdir(parents=True)
            for f in sorted((BASE / "src" / f"s{i}").iterdir()):
                os.link(f, src / f"s{i}" / f.name)
    nbytes = k * fpl * GiB
    lpg = max(1, k // gpus)
    for rep in range(reps + 1):
        dst = BASE / "dst_in"; shutil.rmtree(dst, ignore_errors=True); dst.mkdir()
        t0 = time.perf_counter()
        st = vm.migrate_tree(src, dst, None, None, gpu_mask=mask, lanes_per_gpu=lpg)
        dt = time.perf_counter() - t0
        print(f"  inproc lanes={st['lanes_used']} gpus={st['gpus_used']} rep {rep}{' (warm-up: creates the rings)' if rep == 0 else ''}: "
              f"{nbytes / dt / GiB:6.2f} GiB/s ({dt * 1e3:.0f} ms)", flush=True)
    shutil.rmtree(BASE / "dst_in", ignore_errors=True)


def role_worker(i, gpu, fpl, reps):
    import __graft_entry__ as g
    vm = g.load_pkg()
    vm.init(1 << gpu)
    src, dst = BASE / "src" / f"s{i}", BASE / f"dst_w{i}"
    for rep in range(reps + 1):
        shutil.rmtree(dst, ignore_errors=True); dst.mkdir()
        print("READY", flush=True)
        sys.stdin.readline()
        t0 = time.perf_counter()
        vm.migrate_tree(src, dst, None, None, gpu_mask=1 << gpu)
        print(f"DONE {time.perf_counter() - t0:.4f}", flush=True)
    shutil.rmtree(dst, ignore_errors=True)


def run_procs(k, gpus, fpl, reps, env):
    ps = [subprocess.Popen([sys.executable, __file__, "worker", str(i), str(i % gpus), str(fpl), str(reps)], stdin=subprocess.PIPE,
                           stdout=subprocess.PIPE, text=True, env=env) for i in range(k)]
    for rep in range(reps + 1):
        for p in ps:
            assert p.stdout.readline().strip() == "READY"
        t0 = time.perf_counter()
        for p in ps:
            p.stdin.write("go\n"); p.stdin.flush()
        each = [float(p.stdout.readline().split()[1]) for p in ps]
        dt = time.perf_counter() - t0
        print(f"  procs={k} gpus={gpus} rep {rep}{' (warm-up)' if rep == 0 else ''}: {k * fpl * GiB / dt / GiB:6.2f} GiB/s ({dt * 1e3:.0f} ms; per process "
              f"{' '.join(f'{fpl / e:.1f}' for e in each)} GiB/s)", flush=True)
    for p in ps:
        p.wait()


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "inproc":
        return role_inproc(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]))
    if len(sys.argv) > 1 and sys.argv[1] == "worker":
        return role_worker(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]))
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    fpl = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    gpus = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    variants = sys.argv[5].split(",") if len(sys.argv) > 5 else ["base", "mbind", "nobind"]
    import __graft_entry__ as g
    vm = g.load_pkg()
    shutil.rmtree(BASE, ignore_errors=True); (BASE / "src").mkdir(parents=True)
    for i in range(k):
        vm.datagen_files(BASE / "src" / f"s{i}", 300 + i, fpl, GiB, threads=32)
    print(f"trees: {k} x {fpl} x 1 GiB under {BASE}/src; {gpus} GPU(s)", flush=True)
    env0 = dict(os.environ)

    def inproc(kk, env, tag):
        print(f"[in-process, {kk} lane(s){tag}]", flush=True)
        subprocess.run([sys.executable, __file__, "inproc", str(kk), str(min(gpus, kk)), str(fpl), str(reps)], env=env, check=False)

    try:
        if "base" in variants:
            inproc(1, env0, "")
            kk = 2
            while kk <= k:
                inproc(kk, dict(env0, VMIG_TRACE="1") if kk == k else env0, "")
                print(f"[{kk} processes, one lane each; VMIG_IO_SHARE={kk}]", flush=True)
                run_procs(kk, min(gpus, kk), fpl, reps, dict(env0, VMIG_IO_SHARE=str(kk)))
                kk *= 2
        if "konly" in variants:              # just the full lane count: K lanes in one process against K processes
            inproc(k, dict(env0, VMIG_TRACE="1"), "")
            print(f"[{k} processes, one lane each; VMIG_IO_SHARE={k}]", flush=True)
            run_procs(k, min(gpus, k), fpl, reps, dict(env0, VMIG_IO_SHARE=str(k)))
        for v in variants:                   # rNwM: N readers + M writers per lane, in-process
            if v.startswith("r") and "w" in v and v[1].isdigit():
                r, w = v[1:].split("w")
                inproc(k, dict(env0, VMIG_READERS=r, VMIG_WRITERS=w), f", {r} readers + {w} writers per lane")
        if "mbind" in variants:
            inproc(k, dict(env0, VMIG_RING_MBIND="1"), ", VMIG_RING_MBIND=1")
        if "nobind" in variants:
            inproc(k, dict(env0, VMIG_BIND_IO="0"), ", VMIG_BIND_IO=0")
        if "fewthreads" in variants:
            inproc(k, dict(env0, VMIG_READERS="3", VMIG_WRITERS="5"), ", 3 readers + 5 writers per lane")
    finally:
        shutil.rmtree(BASE, ignore_errors=True)


if __name__ == "__main__":
    main()
