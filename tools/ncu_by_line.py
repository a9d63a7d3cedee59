"""Attribute the warp-state samples of an ncu report to CUDA source lines of one file.

usage: python tools/ncu_by_line.py <rep.ncu-rep> <kernel-name-substring (mangled)> <source-file-basename> [top]

ncu's CSV export of the source page only carries SASS rows; nvdisasm -gi gives, per SASS offset,
the inline chain of source locations.  Every SASS row is charged to the OUTERMOST location that
lies in <source-file-basename> (so inlined helpers are charged to their call site in the kernel).
"""
import collections
import csv
import io
import re
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
rep, kern, fname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
import os
INNER = bool(os.environ.get("NCU_INNER"))  # charge to the innermost location in the file instead (inside lambdas)

tmp = Path(tempfile.mkdtemp())
subprocess.run(["cuobjdump", "-xelf", "all", str(ROOT / "kmeans-gpu_b200" / "lib" / "libkmeans_gpu.so")], cwd=tmp,
               stdout=subprocess.DEVNULL, check=True)
cubin = next(p for p in tmp.glob("*.cubin") if "kmg_host" not in p.name)
sass = subprocess.run(["nvdisasm", "-gi", str(cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and kern in l)
loc_re = re.compile(r'//## File "([^"]+)", line (\d+)')
ins_re = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(.*?);")
# a new location block starts after an instruction
off2line = {}
chain = []
after_ins = True
for l in sass[start + 1:]:
    if l.startswith(".text.") or l.startswith(".nv.") or l.startswith(".section"):
        break
    m = loc_re.search(l)
    if m:
        if after_ins:
            chain = []
            after_ins = False
        chain.append((m.group(1), int(m.group(2))))
        continue
    m = ins_re.search(l)
    if m:
        mine = [c for c in chain if c[0].endswith(fname)]
        off2line[int(m.group(1), 16)] = (mine[0 if INNER else -1][1]) if mine else -1
        after_ins = True

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
base = min(int(r[col["Address"]], 16) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
samples = collections.Counter()
insts = collections.Counter()
stall_by = collections.defaultdict(collections.Counter)
for r in data:
    off = int(r[col["Address"]], 16) - base
    ln = off2line.get(off, -2)
    s = float(r[col["# Samples"]] or 0)
    samples[ln] += s
    insts[ln] += float(r[col["Instructions Executed"]] or 0)
    for st in stalls:
        try:
            stall_by[ln][st[6:]] += float(r[col[st]] or 0)
        except ValueError:
            pass
tot = sum(samples.values()) or 1
toti = sum(insts.values()) or 1
text = (ROOT / "kmeans-gpu_b200" / "csrc" / fname).read_text().splitlines()
print(f"total samples {tot:.0f}, warp instructions {toti:.0f}")
for ln, s in samples.most_common(top):
    top_st = ", ".join(f"{k} {100 * v / max(s, 1):.0f}%" for k, v in stall_by[ln].most_common(3))
    code = text[ln - 1].strip()[:70] if 0 < ln <= len(text) else "?"
    print(f"{ln:5d} {100 * s / tot:5.1f}% smp {100 * insts[ln] / toti:5.1f}% ins  [{top_st}]  {code}")
