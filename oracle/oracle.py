"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU checker for the volume-migration hot path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module; the product (``libvmig.so`` and the
``gpu-docker-api_b200`` package) never does.

What it restates, and from where
--------------------------------
* ``ref_copy`` / ``ref_move``: the reference's copy engine is a shell string, so it is
  *executed verbatim* rather than restated -- ``utils/copy.go:17-27`` (``tar c | tar x``)
  and ``utils/copy.go:116`` (``find | xargs mv; mv *``).  GNU tar 1.35 / coreutils 9.4
  here play the part the host's tar / the ubuntu:22.04 helper image play in production.
* ``xxh64`` / ``hash_blocks`` / ``hash_file``: canonical XXH64 seed 0 (C restatement in
  ``xxh64_ref.c``).  The reference has NO hashing -> "parity unpinned" by
  the reference; pinned to the public xxHash spec through tests/golden/xxh64_kat.json and
  libxxhash 0.8.2 when present.
* ``tree_manifest`` / ``compare_trees``: what "identical result" means for a tree copy:
  the attributes GNU tar restores as root (type, mode, uid, gid, mtime, symlink target,
  hard-link grouping, device numbers) plus the content of every regular file.
* ``block_table_of_tree`` / ``read_table``: independent restatement of the engine's
  block-table file (layout documented in include/vmig.h).
* ``splitmix_bytes`` / ``fnv1a64``: the synthetic-data generator of BASELINE.md §3.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import stat
import struct
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

BLOCK_BYTES = 4 << 20


def build() -> Path:
    """Compile liboracle.so from xxh64_ref.c (gcc; a few hundred ms)."""
    so = _HERE / "liboracle.so"
    src = _HERE / "xxh64_ref.c"
    if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(str(build()))
        L.oracle_xxh64.restype = ctypes.c_uint64
        L.oracle_xxh64.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64]
        L.oracle_hash_blocks.restype = None
        L.oracle_hash_blocks.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.c_uint64, ctypes.c_void_p]
        L.oracle_hash_file.restype = ctypes.c_int64
        L.oracle_hash_file.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint64]
        L.oracle_splitmix_fill.restype = None
        L.oracle_splitmix_fill.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p]
        L.oracle_fnv1a64.restype = ctypes.c_uint64
        L.oracle_fnv1a64.argtypes = [ctypes.c_char_p, ctypes.c_uint64]
        _LIB = L
    return _LIB


# --------------------------------------------------------------------------- hashing
def xxh64(data, seed: int = 0) -> int:
    b = np.frombuffer(memoryview(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    b = np.ascontiguousarray(b)
    return int(lib().oracle_xxh64(b.ctypes.data if b.size else None, b.size, seed))


def xxh64_py(data: bytes, seed: int = 0) -> int:
    """Pure-Python XXH64 (small inputs only): a second, independent restatement."""
    M = (1 << 64) - 1
    P1, P2, P3, P4, P5 = (0x9E3779B185EBCA87, 0xC2B2AE3D27D4EB4F, 0x165667B19E3779F9,
                          0x85EBCA77C2B2AE63, 0x27D4EB2F165667C5)
    rotl = lambda x, r: ((x << r) | (x >> (64 - r))) & M
    rnd = lambda a, x: (rotl((a + x * P2) & M, 31) * P1) & M
    n, p = len(data), 0
    if n >= 32:
        v = [(seed + P1 + P2) & M, (seed + P2) & M, seed & M, (seed - P1) & M]
        while p + 32 <= n:
            for k in range(4):
                v[k] = rnd(v[k], int.from_bytes(data[p + 8 * k:p + 8 * k + 8], "little"))
            p += 32
        h = (rotl(v[0], 1) + rotl(v[1], 7) + rotl(v[2], 12) + rotl(v[3], 18)) & M
        for k in range(4):
            h = ((h ^ rnd(0, v[k])) * P1 + P4) & M
    else:
        h = (seed + P5) & M
    h = (h + n) & M
    while p + 8 <= n:
        h = (rotl(h ^ rnd(0, int.from_bytes(data[p:p + 8], "little")), 27) * P1 + P4) & M
        p += 8
    if p + 4 <= n:
        h = (rotl(h ^ (int.from_bytes(data[p:p + 4], "little") * P1 & M), 23) * P2 + P3) & M
        p += 4
    while p < n:
        h = (rotl(h ^ (data[p] * P5 & M), 11) * P1) & M
        p += 1
    h ^= h >> 33
    h = h * P2 & M
    h ^= h >> 29
    h = h * P3 & M
    h ^= h >> 32
    return h


def hash_blocks(buf: np.ndarray, offs, lens) -> np.ndarray:
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    offs = np.ascontiguousarray(offs, dtype=np.uint64)
    lens = np.ascontiguousarray(lens, dtype=np.uint32)
    out = np.empty(len(offs), dtype=np.uint64)
    lib().oracle_hash_blocks(buf.ctypes.data, offs.ctypes.data, lens.ctypes.data, len(offs), out.ctypes.data)
    return out


def hash_file(path, block_bytes: int = BLOCK_BYTES) -> np.ndarray:
    size = os.path.getsize(path)
    nb = (size + block_bytes - 1) // block_bytes
    out = np.empty(max(nb, 1), dtype=np.uint64)
    got = lib().oracle_hash_file(os.fsencode(path), block_bytes, out.ctypes.data, len(out))
    if got < 0:
        raise OSError(f"oracle_hash_file failed on {path}")
    assert got == nb, (got, nb)
    return out[:nb].copy()


def sanity_buffer(n: int) -> bytes:
    """xxHash's 'sanity buffer' (SURVEY.md Appendix A)."""
    g, out = 2654435761, bytearray(n)
    for i in range(n):
        out[i] = (g >> 56) & 0xFF
        g = (g * 11400714785074694797) & ((1 << 64) - 1)
    return bytes(out)


# --------------------------------------------------------------------------- synthetic data
def fnv1a64(s: bytes) -> int:
    return int(lib().oracle_fnv1a64(s, len(s)))


def splitmix_bytes(seed: int, nbytes: int, first_word: int = 0) -> np.ndarray:
    """nbytes of the SplitMix64 stream `seed`, starting at 8-byte word `first_word` (numpy,
    vectorised -- independent of the C loop in xxh64_ref.c)."""
    nw = (nbytes + 7) // 8
    with np.errstate(over="ignore"):
        j = np.arange(first_word + 1, first_word + nw + 1, dtype=np.uint64)
        z = np.uint64(seed & ((1 << 64) - 1)) + j * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z.view(np.uint8)[:nbytes]


def file_seed(seed: int, relpath: str) -> int:
    return (seed ^ fnv1a64(relpath.encode())) & ((1 << 64) - 1)


# --------------------------------------------------------------------------- the reference, verbatim
def ref_copy(src, dst, check: bool = True) -> subprocess.CompletedProcess:
    """reference utils.CopyDir (utils/copy.go:21-27): sh -c "(cd SRC; tar c .) | (cd DST; tar x)"."""
    return subprocess.run([str(_HERE / "ref_copy.sh"), str(src), str(dst)], check=check,
                          capture_output=True, text=True)


def ref_move(src, dst, check: bool = False) -> subprocess.CompletedProcess:
    """reference moveVolumeData's exec'd command (utils/copy.go:116) on host paths.  The
    reference never reads the exit status (utils/copy.go:122-127), hence check=False."""
    return subprocess.run([str(_HERE / "ref_move.sh"), str(src), str(dst)], check=check,
                          capture_output=True, text=True)


# --------------------------------------------------------------------------- tree comparison
def _sha(path) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        while True:
            b = f.read(1 << 22)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


def tree_manifest(root, content: bool = True, mtime_ns: bool = False) -> dict:
    """relpath -> attribute tuple, for every entry under root (root itself is '.')."""
    root = os.fspath(root)
    out, inode_first = {}, {}
    for dirpath, dirnames, filenames in os.walk(root):
        entries = [dirpath] if dirpath == root else []
        entries += [os.path.join(dirpath, n) for n in dirnames + filenames]
        for p in entries:
            st = os.lstat(p)
            rel = os.path.relpath(p, root)
            mt = st.st_mtime_ns if mtime_ns else st.st_mtime_ns // 1_000_000_000
            kind = stat.S_IFMT(st.st_mode)
            rec = {"type": kind, "mode": stat.S_IMODE(st.st_mode), "uid": st.st_uid, "gid": st.st_gid}
            if kind == stat.S_IFLNK:
                rec["target"] = os.readlink(p)
                rec["mtime"] = mt
            elif kind == stat.S_IFREG:
                rec["size"] = st.st_size
                rec["mtime"] = mt
                if content:
                    rec["sha256"] = _sha(p)
                if st.st_nlink > 1:
                    key = (st.st_dev, st.st_ino)
                    rec["hardlink_to"] = inode_first.setdefault(key, rel)
            elif kind in (stat.S_IFCHR, stat.S_IFBLK):
                rec["rdev"] = st.st_rdev
                rec["mtime"] = mt
            elif kind == stat.S_IFDIR:
                rec["mtime"] = mt
            else:
                rec["mtime"] = mt
            out[rel] = rec
    # hard-link grouping must not depend on walk order: canonicalise to the sorted-first member
    groups = {}
    for rel, rec in out.items():
        if "hardlink_to" in rec:
            groups.setdefault(rec["hardlink_to"], []).append(rel)
    for members in groups.values():
        canon = min(members)
        for m in members:
            out[m]["hardlink_to"] = canon
    return out


def compare_trees(a, b, content: bool = True, mtime_ns: bool = False, ignore_root_mtime: bool = False) -> list:
    """List of human-readable differences between two trees ([] == identical)."""
    ma, mb = tree_manifest(a, content, mtime_ns), tree_manifest(b, content, mtime_ns)
    diffs = []
    for rel in sorted(set(ma) | set(mb)):
        if rel not in ma:
            diffs.append(f"only in B: {rel}")
        elif rel not in mb:
            diffs.append(f"only in A: {rel}")
        else:
            ra, rb = dict(ma[rel]), dict(mb[rel])
            if ignore_root_mtime and rel == ".":
                ra.pop("mtime", None), rb.pop("mtime", None)
            if ra != rb:
                keys = [k for k in set(ra) | set(rb) if ra.get(k) != rb.get(k)]
                diffs.append(f"{rel}: " + ", ".join(f"{k} {ra.get(k)!r} != {rb.get(k)!r}" for k in sorted(keys)))
    return diffs


# --------------------------------------------------------------------------- block table (restated)
TABLE_MAGIC = b"VMIGBT02"
TABLE_MAGIC_V1 = b"VMIGBT01"


def block_table_of_tree(root, block_bytes: int = BLOCK_BYTES):
    """Oracle block table: ([(relpath, size, first_block)], hashes) with regular files sorted
    bytewise by relative path; hard-linked files appear once per path (each path is a file)."""
    root = os.fspath(root)
    files = []
    for dirpath, _dirnames, filenames in os.walk(root):
        for n in filenames:
            p = os.path.join(dirpath, n)
            st = os.lstat(p)
            if stat.S_ISREG(st.st_mode):
                files.append((os.fsencode(os.path.relpath(p, root)), st.st_size, p))
    files.sort(key=lambda t: t[0])
    entries, hashes, first = [], [], 0
    for rel, size, p in files:
        h = hash_file(p, block_bytes)
        entries.append((rel, size, first))
        hashes.append(h)
        first += len(h)
    return entries, (np.concatenate(hashes) if hashes else np.empty(0, np.uint64))


def read_table(path):
    """Parse a block-table file written by the engine (format: include/vmig.h).  "VMIGBT02" carries the identity
    (inode, ctime_ns) of the file each entry speaks for; "VMIGBT01" does not (identity (0, 0))."""
    raw = Path(path).read_bytes()
    assert raw[:8] in (TABLE_MAGIC, TABLE_MAGIC_V1), raw[:8]
    v2 = raw[:8] == TABLE_MAGIC
    block_bytes, algo, n_files, n_blocks = struct.unpack_from("<IIQQ", raw, 8)
    off = 32
    entries, identity = [], []
    for _ in range(n_files):
        (plen,) = struct.unpack_from("<I", raw, off)
        off += 4
        rel = raw[off:off + plen]
        off += plen
        size, first = struct.unpack_from("<QQ", raw, off)
        off += 16
        if v2:
            identity.append(struct.unpack_from("<Qq", raw, off))
            off += 16
        else:
            identity.append((0, 0))
        entries.append((rel, size, first))
    hashes = np.frombuffer(raw, dtype="<u8", count=n_blocks, offset=off).copy()
    assert off + 8 * n_blocks == len(raw), (off, n_blocks, len(raw))
    return {"block_bytes": block_bytes, "algo": algo, "entries": entries, "identity": identity, "hashes": hashes}
