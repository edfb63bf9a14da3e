"""Which open(2) flags does cuFileHandleRegister accept on this box?  (round 2, N4)  python profiles/scripts/r2_cufile_flags.py [dir=/tmp]"""
import ctypes as C, os, sys, tempfile
lib = C.CDLL("libcufile.so.0")
class Err(C.Structure): _fields_ = [("err", C.c_int), ("cu_err", C.c_int)]
class Descr(C.Structure): _fields_ = [("type", C.c_int), ("fd", C.c_int), ("pad", C.c_int), ("fs_ops", C.c_void_p)]
lib.cuFileDriverOpen.restype = Err
lib.cuFileHandleRegister.restype = Err
lib.cuFileHandleRegister.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Descr)]
e = lib.cuFileDriverOpen(); print("cuFileDriverOpen:", e.err, e.cu_err)
d = sys.argv[1] if len(sys.argv) > 1 else "/tmp"
p = os.path.join(d, "cufile_flag_probe.bin"); open(p, "wb").write(b"x" * 8192)
combos = {"O_RDONLY": os.O_RDONLY, "O_RDONLY|O_DIRECT": os.O_RDONLY | os.O_DIRECT, "O_RDONLY|O_NONBLOCK": os.O_RDONLY | os.O_NONBLOCK,
          "O_RDONLY|O_NOFOLLOW": os.O_RDONLY | os.O_NOFOLLOW, "O_RDONLY|O_CLOEXEC": os.O_RDONLY | os.O_CLOEXEC,
          "O_WRONLY": os.O_WRONLY, "O_WRONLY|O_DIRECT": os.O_WRONLY | os.O_DIRECT, "O_RDWR": os.O_RDWR, "O_RDWR|O_DIRECT": os.O_RDWR | os.O_DIRECT}
for name, fl in combos.items():
    try:
        fd = os.open(p, fl)
    except OSError as ex:
        print(f"{name:24s} open failed: {ex}"); continue
    h = C.c_void_p(); ds = Descr(1, fd, 0, None)
    r = lib.cuFileHandleRegister(C.byref(h), C.byref(ds))
    print(f"{name:24s} F_GETFL=0o{__import__('fcntl').fcntl(fd, __import__('fcntl').F_GETFL):o}  cuFileHandleRegister err={r.err}")
    os.close(fd)
os.unlink(p)
