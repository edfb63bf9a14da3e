"""Which descriptors does cuFileHandleRegister accept on this box?  (round 2, N4)
  python profiles/scripts/r2_cufile_flags.py [dir ...]      default: /tmp /dev/shm
Runs twice: stock configuration, then CUFILE_ENV_PATH_JSON pointing at a json that forces the compatibility mode."""
import ctypes as C, fcntl, json, os, subprocess, sys, tempfile

def probe(dirs):
    lib = C.CDLL("libcufile.so.0")
    class Err(C.Structure): _fields_ = [("err", C.c_int), ("cu_err", C.c_int)]
    class Descr(C.Structure): _fields_ = [("type", C.c_int), ("_p0", C.c_int), ("fd", C.c_int), ("_p1", C.c_int), ("fs_ops", C.c_void_p)]   # CUfileDescr_t: enum, union{int fd; void*} at offset 8, fs_ops
    lib.cuFileDriverOpen.restype = Err
    lib.cuFileHandleRegister.restype = Err
    lib.cuFileHandleRegister.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Descr)]
    e = lib.cuFileDriverOpen(); print("cuFileDriverOpen:", e.err, e.cu_err, flush=True)
    for d in dirs:
        p = os.path.join(d, "cufile_flag_probe.bin"); open(p, "wb").write(b"x" * 8192)
        fs = subprocess.run(["stat", "-f", "-c", "%T", d], capture_output=True, text=True).stdout.strip()
        for name, fl in {"O_RDONLY": os.O_RDONLY, "O_RDONLY|O_DIRECT": os.O_RDONLY | os.O_DIRECT, "O_RDWR|O_DIRECT": os.O_RDWR | os.O_DIRECT,
                         "O_RDONLY|O_NOFOLLOW": os.O_RDONLY | os.O_NOFOLLOW, "O_WRONLY": os.O_WRONLY}.items():
            try:
                fd = os.open(p, fl)
            except OSError as ex:
                print(f"{d} ({fs}) {name:22s} open failed: {ex}"); continue
            h = C.c_void_p(); ds = Descr(1, 0, fd, 0, None)
            r = lib.cuFileHandleRegister(C.byref(h), C.byref(ds))
            print(f"{d} ({fs}) {name:22s} F_GETFL=0o{fcntl.fcntl(fd, fcntl.F_GETFL):o}  cuFileHandleRegister err={r.err}", flush=True)
            os.close(fd)
        os.unlink(p)

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    probe(sys.argv[2:]); sys.exit(0)
dirs = sys.argv[1:] or ["/tmp", "/dev/shm"]
work = tempfile.mkdtemp(prefix="cufile_probe_")
for label, env in (("stock configuration", {}), ("forced compatibility mode", None)):
    e = dict(os.environ)
    if env is None:
        cfg = {"logging": {"dir": work, "level": "INFO"}, "properties": {"allow_compat_mode": True, "force_compat_mode": True, "use_poll_mode": False},
               "fs": {"generic": {"posix_unaligned_writes": True}}}
        jp = os.path.join(work, "cufile.json"); json.dump(cfg, open(jp, "w"))
        e["CUFILE_ENV_PATH_JSON"] = jp
    print(f"== {label}", flush=True)
    subprocess.run([sys.executable, __file__, "--child"] + dirs, env=e, cwd=work)
for f in sorted(os.listdir(work)):
    if f.startswith("cufile") and f.endswith(".log"):
        print(f"== {f} (tail)"); print("".join(open(os.path.join(work, f), errors="replace").readlines()[-25:]))
# the engine's own attempt (its descriptors are re-opened plain), with libcufile's log
import shutil
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
try:
    import __graft_entry__ as g
    vm = g.load_pkg()
    for d in dirs:
        a, b = os.path.join(d, "cf_src"), os.path.join(d, "cf_dst")
        shutil.rmtree(a, ignore_errors=True); shutil.rmtree(b, ignore_errors=True); os.makedirs(a); os.makedirs(b)
        open(os.path.join(a, "x.bin"), "wb").write(os.urandom(5 << 20))
        os.chdir(work)
        for fl, nm in ((vm.F_CUFILE, "F_CUFILE"), (vm.F_CUFILE | vm.F_DIRECT_IO, "F_CUFILE|F_DIRECT_IO")):
            try:
                st = vm.migrate_tree(a, b, flags=fl | vm.F_VERIFY)
                ok = open(os.path.join(a, "x.bin"), "rb").read() == open(os.path.join(b, "x.bin"), "rb").read()
                print(f"engine {nm} on {d}: ok={ok} bytes={st['bytes_total']} direct={st['files_direct']}")
            except vm.VmigError as ex:
                print(f"engine {nm} on {d}: {ex}")
        shutil.rmtree(a, ignore_errors=True); shutil.rmtree(b, ignore_errors=True)
    for f in sorted(os.listdir(work)):
        if f.startswith("cufile") and f.endswith(".log"):
            tail = [l for l in open(os.path.join(work, f), errors="replace").readlines() if "NUMA" not in l][-12:]
            print(f"== {f} (tail, after the engine runs)"); print("".join(tail))
except Exception as ex:      # noqa: BLE001
    print("engine probe failed:", ex)
print("nvidia-fs module:", "present" if os.path.exists("/proc/driver/nvidia-fs") else "absent", "| lsmod:", subprocess.run("lsmod | grep -c nvidia", shell=True, capture_output=True, text=True).stdout.strip())
