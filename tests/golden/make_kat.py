"""Regenerates tests/golden/xxh64_kat.json from python-xxhash (XXH 0.8.2) and cross-checks
libxxhash.so.0 when present.  The reference has no hashing, so these canonical
XXH64 known-answer vectors (SURVEY.md Appendix A + extras) are what pins the hash oracle.
Run here (dev container):  python tests/golden/make_kat.py
"""
import ctypes, ctypes.util, json, os, sys
import xxhash
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle.oracle import sanity_buffer

lens = [0, 1, 3, 4, 7, 8, 14, 15, 16, 17, 31, 32, 33, 63, 64, 65, 95, 96, 127, 128, 222, 1023, 1024,
        2047, 2048, 2049, 4095, 4096, 4097, 65535, 65536, 65537, 1 << 20, (4 << 20) - 1, 4 << 20]
buf = sanity_buffer(max(lens))
try:
    L = ctypes.CDLL("libxxhash.so.0"); L.XXH64.restype = ctypes.c_uint64
    L.XXH64.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint64]
except OSError:
    L = None
kat = {"algo": "XXH64", "seed": 0, "source": f"python-xxhash {xxhash.VERSION} / xxHash {xxhash.XXHASH_VERSION}",
       "sanity": {}, "ascii": {}, "zeros_4MiB": None}
for n in lens:
    h = xxhash.xxh64_intdigest(buf[:n], 0)
    if L is not None:
        assert L.XXH64(buf[:n], n, 0) == h
    kat["sanity"][str(n)] = f"0x{h:016X}"
for s in ["", "a", "abc", "message digest", "abcdefghijklmnopqrstuvwxyz"]:
    kat["ascii"][s] = f"0x{xxhash.xxh64_intdigest(s.encode(), 0):016X}"
kat["zeros_4MiB"] = f"0x{xxhash.xxh64_intdigest(bytes(4 << 20), 0):016X}"
json.dump(kat, open(os.path.join(os.path.dirname(__file__), "xxh64_kat.json"), "w"), indent=1)
print("wrote", len(lens), "vectors")
