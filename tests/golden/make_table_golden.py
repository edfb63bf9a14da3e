"""Writes tests/golden/table_v2.vmig + table_v2.json: a hand-assembled block table in format VMIGBT02 (include/vmig.h), so
that both parsers -- libvmig's (vmig_table_info_read / vmig_table_hashes) and the oracle's read_table -- stay pinned to the
committed bytes.  python tests/golden/make_table_golden.py"""
import json, struct
from pathlib import Path
HERE = Path(__file__).resolve().parent
BB = 4 << 20
entries = [(b"a/deep/x.bin", 9 * (1 << 20) + 1), (b"b.bin", 100), (b"empty", 0), ("café.txt".encode(), BB), (b"z/last", BB + 1)]
entries.sort()
ident = [(1000 + i, 1_700_000_000_000_000_000 + 12345 * i) for i in range(len(entries))]
raw, first, hashes = b"", 0, []
for (rel, size), (ino, ct) in zip(entries, ident):
    raw += struct.pack("<I", len(rel)) + rel + struct.pack("<QQQq", size, first, ino, ct)
    nb = (size + BB - 1) // BB
    hashes += [(0x9E3779B97F4A7C15 * (first + k + 1)) & ((1 << 64) - 1) for k in range(nb)]
    first += nb
blob = b"VMIGBT02" + struct.pack("<IIQQ", BB, 1, len(entries), len(hashes)) + raw + b"".join(struct.pack("<Q", h) for h in hashes)
(HERE / "table_v2.vmig").write_bytes(blob)
(HERE / "table_v2.json").write_text(json.dumps({
    "block_bytes": BB, "algo": 1, "n_files": len(entries), "n_blocks": len(hashes), "bytes_total": sum(s for _, s in entries),
    "entries": [[r.decode("utf-8"), s] for r, s in entries], "identity": ident, "hashes": [f"{h:016x}" for h in hashes]}, indent=1))
print(len(blob), "bytes,", len(hashes), "hashes")
