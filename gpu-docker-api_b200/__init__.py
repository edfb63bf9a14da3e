"""gpu-docker-api_b200 -- host-side mirror of the reference's copy interface over libvmig.

The reference (XShengTech/gpu-docker-api, Go) exposes the bulk-copy path as three package-level
functions in ``utils``.  The Go toolchain is absent in this image, so the mirror
that tests and the bench drive is this ctypes binding; the cgo shim a maintainer would add is in
INTEGRATION.md.  Names, argument meaning and error behaviour follow the reference:

    CopyDir(src, dest)                                   utils/copy.go:21-27
    CopyOldMergedToNewContainerMerged(old, new)          utils/copy.go:31-46
    GetContainerMergedLayer(name)                        utils/copy.go:48-54
    CopyOldMountPointToContainerMountPoint(old, new)     utils/copy.go:58-63  (moveVolumeData :74-128)
    GetVolumeMountPoint(name)                            utils/copy.go:65-72
    DirSize(path) / ToBytes("20GB")                      utils/file.go:13-48

Go returns ``error``; here a failure raises :class:`VmigError` (``.code`` is the VMIG_E* value,
the message is libvmig's thread-local detail).  Docker lookups (container -> overlay2 UpperDir,
volume -> Mountpoint) stay in the control plane: they are injected through :func:`set_resolver`.

This package never imports ``oracle/`` and has no CPU fallback: if ``libvmig.so`` is missing the
import fails, and without an sm_100 GPU every data-path call raises VmigError(VMIG_ENOGPU).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Callable, Optional

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libvmig.so"

VMIG_OK, VMIG_EINVAL, VMIG_ENOGPU, VMIG_ECUDA, VMIG_EIO = 0, -1, -2, -3, -4
VMIG_ENOMEM, VMIG_ETABLE, VMIG_EFAULT, VMIG_ENOTDIR, VMIG_ESRCCHANGED, VMIG_EVERIFY = -5, -6, -7, -8, -9, -10
F_MOVE_SRC, F_SKIP_HIDDEN_TOPDIRS, F_MTIME_NS, F_NO_METADATA, F_HASH_ONLY, F_VERIFY = 0x01, 0x02, 0x04, 0x08, 0x10, 0x20
F_PRUNE, F_DIRECT_IO, F_CUFILE = 0x40, 0x80, 0x100
BLOCK_BYTES = 4 << 20
MOVE_DIR_FLAGS = F_MOVE_SRC | F_VERIFY          # what vmig_move_dir passes (include/vmig.h)

EXPORTS = [
    "vmig_init", "vmig_shutdown", "vmig_device_count", "vmig_strerror", "vmig_last_error", "vmig_version",
    "vmig_migrate_tree", "vmig_copy_dir", "vmig_move_dir", "vmig_migrate_buffer", "vmig_host_alloc",
    "vmig_host_free", "vmig_hash_blocks", "vmig_resident_open", "vmig_resident_close", "vmig_resident_fill",
    "vmig_resident_set_len", "vmig_resident_upload", "vmig_resident_download", "vmig_resident_flip",
    "vmig_resident_set_prior", "vmig_resident_pass", "vmig_resident_results", "vmig_table_info_read",
    "vmig_table_hashes", "vmig_dir_size", "vmig_to_bytes", "vmig_datagen_files", "vmig_manifest", "vmig_link_probe", "vmig_thread_plan",
]


class Opts(C.Structure):
    _fields_ = [("gpu_mask", C.c_uint32), ("block_bytes", C.c_uint32), ("streams_per_gpu", C.c_uint32),
                ("flags", C.c_uint32), ("io_threads", C.c_uint32), ("lanes_per_gpu", C.c_uint32),
                ("reserved", C.c_uint32 * 2)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "bytes_total", "bytes_h2d", "bytes_d2h", "bytes_written", "blocks_total", "blocks_skipped", "files", "dirs",
        "symlinks", "hardlinks", "specials", "kernel_launches", "ns_total", "ns_walk", "ns_plan", "ns_data",
        "ns_meta", "ns_table")] + [("ms_kernel", C.c_double), ("gpus_used", C.c_uint32), ("lanes_used", C.c_uint32), ("pruned", C.c_uint64),
                                                    ("files_untrusted", C.c_uint64), ("files_direct", C.c_uint64)]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


class TableInfo(C.Structure):
    _fields_ = [("block_bytes", C.c_uint32), ("algo", C.c_uint32), ("n_files", C.c_uint64),
                ("n_blocks", C.c_uint64), ("bytes_total", C.c_uint64)]


class VmigError(RuntimeError):
    def __init__(self, code: int, what: str):
        self.code = code
        super().__init__(f"{what}: {strerror(code)} ({code}): {last_error()}")


if not LIB_PATH.exists():
    raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(make -C gpu-docker-api_b200/csrc). There is no CPU fallback.")
_lib = C.CDLL(str(LIB_PATH))
_u64p, _u32p, _u8p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
_sig = {
    "vmig_init": (C.c_int, [C.c_uint32]), "vmig_shutdown": (None, []), "vmig_device_count": (C.c_int, []),
    "vmig_strerror": (C.c_char_p, [C.c_int]), "vmig_last_error": (C.c_char_p, []), "vmig_version": (C.c_char_p, []),
    "vmig_migrate_tree": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(Opts), C.POINTER(Stats)]),
    "vmig_copy_dir": (C.c_int, [C.c_char_p, C.c_char_p]), "vmig_move_dir": (C.c_int, [C.c_char_p, C.c_char_p]),
    "vmig_migrate_buffer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.POINTER(Opts), C.POINTER(Stats)]),
    "vmig_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint64]), "vmig_host_free": (None, [C.c_void_p]),
    "vmig_hash_blocks": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p,
                                   C.POINTER(C.c_double)]),
    "vmig_resident_open": (C.c_int, [C.c_int, C.c_uint64, C.c_uint32, C.POINTER(C.c_void_p)]),
    "vmig_resident_close": (None, [C.c_void_p]), "vmig_resident_fill": (C.c_int, [C.c_void_p, C.c_uint64]),
    "vmig_resident_set_len": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32]),
    "vmig_resident_upload": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32]),
    "vmig_resident_download": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32]),
    "vmig_resident_flip": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
    "vmig_resident_set_prior": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "vmig_resident_pass": (C.c_int, [C.c_void_p, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "vmig_resident_results": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]),
    "vmig_table_info_read": (C.c_int, [C.c_char_p, C.POINTER(TableInfo)]),
    "vmig_table_hashes": (C.c_int, [C.c_char_p, C.c_void_p, C.c_uint64]),
    "vmig_dir_size": (C.c_int, [C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_uint64)]),
    "vmig_to_bytes": (C.c_int, [C.c_char_p, C.POINTER(C.c_int64)]),
    "vmig_datagen_files": (C.c_int, [C.c_char_p, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32]),
    "vmig_manifest": (C.c_int, [C.c_char_p, C.c_uint32, C.c_uint32, C.c_char_p, C.POINTER(Stats)]),
    "vmig_link_probe": (C.c_int, [C.c_int, C.c_uint64, C.POINTER(C.c_double)]),
    "vmig_thread_plan": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
}
for _n, (_r, _a) in _sig.items():
    _f = getattr(_lib, _n)
    _f.restype, _f.argtypes = _r, _a


def lib() -> C.CDLL:
    return _lib


def strerror(code: int) -> str:
    return _lib.vmig_strerror(code).decode()


def last_error() -> str:
    return _lib.vmig_last_error().decode(errors="replace")


def version() -> str:
    return _lib.vmig_version().decode()


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise VmigError(rc, what)


def init(gpu_mask: int = 0) -> None:
    _check(_lib.vmig_init(gpu_mask), "vmig_init")


def shutdown() -> None:
    _lib.vmig_shutdown()


def device_count() -> int:
    n = _lib.vmig_device_count()
    if n < 0:
        raise VmigError(n, "vmig_device_count")
    return n


def _b(p) -> Optional[bytes]:
    return None if p is None else os.fsencode(p)


# ------------------------------------------------------------------------------------------------
# the reference's interface
_resolver: dict = {"container": None, "volume": None}


def set_resolver(container_upperdir: Optional[Callable[[str], str]] = None,
                 volume_mountpoint: Optional[Callable[[str], str]] = None) -> None:
    """Inject the Docker lookups the control plane owns (ContainerInspect -> GraphDriver.Data
    ["UpperDir"], VolumeInspect -> Mountpoint: reference utils/copy.go:48-54, 65-72)."""
    _resolver["container"], _resolver["volume"] = container_upperdir, volume_mountpoint


def GetContainerMergedLayer(name: str) -> str:
    """reference utils/copy.go:48-54 -- the overlay2 UpperDir (despite the name)."""
    if _resolver["container"] is None:
        raise VmigError(VMIG_EINVAL, f"docker.ContainerInspect failed, name: {name} (no resolver injected)")
    p = _resolver["container"](name)
    if not p:
        raise VmigError(VMIG_EINVAL, f"docker.ContainerInspect failed, name: {name}")
    return p


def GetVolumeMountPoint(name: str) -> str:
    """reference utils/copy.go:65-72."""
    if _resolver["volume"] is None:
        raise VmigError(VMIG_EINVAL, f"docker.VolumeInspect failed, name: {name} (no resolver injected)")
    p = _resolver["volume"](name)
    if not p:
        raise VmigError(VMIG_EINVAL, f"docker.VolumeInspect failed, name: {name}")
    return p


def CopyDir(src: str, dest: str) -> None:
    """reference utils.CopyDir (utils/copy.go:21-27): `(cd src; tar c .) | (cd dest; tar x)`."""
    _check(_lib.vmig_copy_dir(_b(src), _b(dest)), f"CopyDir failed, src:{src}, dest: {dest}")


def CopyOldMergedToNewContainerMerged(oldContainer: str, newContainer: str) -> None:
    """reference utils/copy.go:31-46 (called at services/replicaset.go:333,421,831)."""
    CopyDir(GetContainerMergedLayer(oldContainer), GetContainerMergedLayer(newContainer))


def CopyOldMountPointToContainerMountPoint(oldVolume: str, newVolume: str) -> None:
    """reference utils/copy.go:58-63 -> moveVolumeData (:74-128), called at services/volume.go:150.
    Move semantics, synchronous, status checked (the reference neither waits nor checks)."""
    src, dst = GetVolumeMountPoint(oldVolume), GetVolumeMountPoint(newVolume)
    _check(_lib.vmig_move_dir(_b(src), _b(dst)), f"moveData failed, src:{src}, dest: {dst}")


def DirSize(path: str) -> int:
    """reference utils/file.go:13-22."""
    out = C.c_int64(0)
    _check(_lib.vmig_dir_size(_b(path), C.byref(out), None), "DirSize")
    return out.value


def ToBytes(origin: str) -> int:
    """reference utils/file.go:24-48."""
    out = C.c_int64(0)
    _check(_lib.vmig_to_bytes(origin.encode(), C.byref(out)), "ToBytes")
    return out.value


# ------------------------------------------------------------------------------------------------
# engine entry points
def migrate_tree(src, dst, prior_table=None, out_table=None, *, gpu_mask: int = 0, flags: int = 0,
                 block_bytes: int = 0, io_threads: int = 0, streams_per_gpu: int = 0, lanes_per_gpu: int = 0) -> dict:
    o = Opts(gpu_mask=gpu_mask, block_bytes=block_bytes, flags=flags, io_threads=io_threads,
             streams_per_gpu=streams_per_gpu, lanes_per_gpu=lanes_per_gpu)
    st = Stats()
    _check(_lib.vmig_migrate_tree(_b(src), _b(dst), _b(prior_table), _b(out_table), C.byref(o), C.byref(st)),
           f"vmig_migrate_tree({src} -> {dst})")
    return st.as_dict()


def hash_tree(src, out_table, *, gpu_mask: int = 0, block_bytes: int = 0) -> dict:
    """Block table of a tree without copying it (VMIG_F_HASH_ONLY)."""
    return migrate_tree(src, None, None, out_table, gpu_mask=gpu_mask, flags=F_HASH_ONLY, block_bytes=block_bytes)


def hash_blocks(buf: np.ndarray, offs, lens, gpu: int = 0):
    """K1 direct: returns (hashes u64[n], kernel_ms)."""
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    offs = np.ascontiguousarray(offs, dtype=np.uint64)
    lens = np.ascontiguousarray(lens, dtype=np.uint32)
    out = np.zeros(len(offs), dtype=np.uint64)
    ms = C.c_double(0)
    _check(_lib.vmig_hash_blocks(gpu, buf.ctypes.data if buf.size else None, offs.ctypes.data, lens.ctypes.data,
                                 len(offs), out.ctypes.data, C.byref(ms)), "vmig_hash_blocks")
    return out, ms.value


class PinnedBuffer:
    """Page-locked host memory from vmig_host_alloc, exposed as a numpy uint8 array."""

    def __init__(self, nbytes: int):
        p = C.c_void_p()
        _check(_lib.vmig_host_alloc(C.byref(p), nbytes), "vmig_host_alloc")
        self.ptr, self.nbytes = p.value, nbytes
        self.array = np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(self.ptr))

    def free(self):
        if self.ptr:
            self.array = None
            _lib.vmig_host_free(self.ptr)
            self.ptr = None


def migrate_buffer(src: np.ndarray, dst: Optional[np.ndarray], prior_hashes=None, prior_valid=None, *,
                   gpu_mask: int = 0, flags: int = 0, block_bytes: int = 0):
    """Host buffer -> host buffer through H2D / hash / diff / D2H.  Returns (hashes, stats)."""
    bb = block_bytes or BLOCK_BYTES
    n = src.size
    nb = (n + bb - 1) // bb
    out = np.zeros(nb, dtype=np.uint64)
    ph = None if prior_hashes is None else np.ascontiguousarray(prior_hashes, dtype=np.uint64)
    pv = None if prior_valid is None else np.ascontiguousarray(prior_valid, dtype=np.uint8)
    o = Opts(gpu_mask=gpu_mask, block_bytes=block_bytes, flags=flags)
    st = Stats()
    _check(_lib.vmig_migrate_buffer(src.ctypes.data, None if dst is None else dst.ctypes.data, n,
                                    None if ph is None else ph.ctypes.data, None if pv is None else pv.ctypes.data,
                                    out.ctypes.data, C.byref(o), C.byref(st)), "vmig_migrate_buffer")
    return out, st.as_dict()


class Resident:
    """HBM-resident batch of equally sized block slots (the "block-hash GB/s" metric)."""

    def __init__(self, n_blocks: int, block_bytes: int = BLOCK_BYTES, gpu: int = 0):
        h = C.c_void_p()
        _check(_lib.vmig_resident_open(gpu, n_blocks, block_bytes, C.byref(h)), "vmig_resident_open")
        self.h, self.n, self.block_bytes = h, n_blocks, block_bytes

    def close(self):
        if self.h:
            _lib.vmig_resident_close(self.h)
            self.h = None

    def fill(self, seed: int):
        _check(_lib.vmig_resident_fill(self.h, seed & ((1 << 64) - 1)), "vmig_resident_fill")

    def set_len(self, block: int, length: int):
        _check(_lib.vmig_resident_set_len(self.h, block, length), "vmig_resident_set_len")

    def upload(self, block: int, data: np.ndarray):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        _check(_lib.vmig_resident_upload(self.h, block, data.ctypes.data if data.size else None, data.size),
               "vmig_resident_upload")

    def download(self, block: int, length: int) -> np.ndarray:
        out = np.empty(length, dtype=np.uint8)
        _check(_lib.vmig_resident_download(self.h, block, out.ctypes.data, length), "vmig_resident_download")
        return out

    def flip(self, blocks):
        b = np.ascontiguousarray(blocks, dtype=np.uint64)
        _check(_lib.vmig_resident_flip(self.h, b.ctypes.data if b.size else None, b.size), "vmig_resident_flip")

    def set_prior(self, hashes=None, valid=None):
        if hashes is None:
            _check(_lib.vmig_resident_set_prior(self.h, None, None), "vmig_resident_set_prior")
            return
        h = np.ascontiguousarray(hashes, dtype=np.uint64)
        v = np.ascontiguousarray(valid, dtype=np.uint8)
        assert h.size == self.n and v.size == self.n
        _check(_lib.vmig_resident_set_prior(self.h, h.ctypes.data, v.ctypes.data), "vmig_resident_set_prior")

    def run(self, iters: int = 1):
        """Returns (ms of the last hash kernel, ms of all iterations) by CUDA events."""
        a, b = C.c_double(0), C.c_double(0)
        _check(_lib.vmig_resident_pass(self.h, iters, C.byref(a), C.byref(b)), "vmig_resident_pass")
        return a.value, b.value

    def results(self):
        hashes = np.empty(self.n, dtype=np.uint64)
        surv = np.empty(self.n, dtype=np.uint32)
        ns = C.c_uint64(0)
        _check(_lib.vmig_resident_results(self.h, hashes.ctypes.data, surv.ctypes.data, C.byref(ns)),
               "vmig_resident_results")
        return hashes, surv[:ns.value].copy()


def link_probe(gpu: int = 0, nbytes: int = 4 << 30) -> dict:
    """Pinned cudaMemcpyAsync sweep: GB/s host->HBM, HBM->host, and both at once (the e2e roofline denominators)."""
    g = (C.c_double * 4)()
    _check(_lib.vmig_link_probe(gpu, nbytes, g), "vmig_link_probe")
    return {"h2d_GBps": g[0], "d2h_GBps": g[1], "duplex_h2d_GBps": g[2], "duplex_d2h_GBps": g[3]}


def thread_plan(lanes: int = 1, n_gpus: int = 1, *, flags: int = 0, has_prior: bool = False):
    """(readers, writers) per lane the engine would use on this box for such a call (no GPU needed)."""
    r, w = C.c_uint32(0), C.c_uint32(0)
    _check(_lib.vmig_thread_plan(lanes, n_gpus, flags, int(has_prior), C.byref(r), C.byref(w)), "vmig_thread_plan")
    return r.value, w.value


def manifest(src, out_table=None, *, flags: int = 0, block_bytes: int = 0) -> dict:
    """The engine's metadata pass alone (no GPU): totals, and optionally the manifest as a zero-hash table."""
    st = Stats()
    _check(_lib.vmig_manifest(_b(src), flags, block_bytes, _b(out_table), C.byref(st)), f"vmig_manifest({src})")
    return st.as_dict()


def table_info(path) -> dict:
    t = TableInfo()
    _check(_lib.vmig_table_info_read(_b(path), C.byref(t)), "vmig_table_info_read")
    return {n: getattr(t, n) for n, _ in t._fields_}


def table_hashes(path) -> np.ndarray:
    n = table_info(path)["n_blocks"]
    out = np.empty(n, dtype=np.uint64)
    _check(_lib.vmig_table_hashes(_b(path), out.ctypes.data if n else None, n), "vmig_table_hashes")
    return out


def datagen_files(dirpath, seed: int, n_files: int, file_bytes: int, threads: int = 16) -> None:
    _check(_lib.vmig_datagen_files(_b(dirpath), seed & ((1 << 64) - 1), n_files, file_bytes, threads),
           "vmig_datagen_files")
