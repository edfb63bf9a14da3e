"""Run the HBM-resident hot path (xxh64_blocks + diff_select) a few times -- the target of the
ncu captures in profiles/.  usage: python profiles/scripts/k1_resident.py [n_blocks=2560] [iters=3]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import __graft_entry__ as g

vm = g.load_pkg()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2560
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
vm.init(1)
r = vm.Resident(n, 4 << 20)
r.fill(0xB200)
for i in range(iters):
    ms_hash, ms_total = r.run(1)
    print(f"n_blocks={n} iter={i} xxh64_blocks {ms_hash:.4f} ms = {n * (4 << 20) / ms_hash / 1e6:.1f} GB/s; pass {ms_total:.4f} ms", flush=True)
r.close()
