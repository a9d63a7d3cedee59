"""Opcode histogram of an address range of a kernel's SASS (static count), e.g. the hot loop.
usage: sass_loop_mix.py <lib.so> <mangled-substring> <lo-hex> <hi-hex> [skip_lo-skip_hi ...]"""
import collections
import re
import subprocess
import sys

lib, pat, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
skips = [tuple(int(x, 16) for x in a.split("-")) for a in sys.argv[5:]]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur = None
hist = collections.Counter()
n = 0
for line in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur is None or pat not in cur:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if not m:
        continue
    a = int(m.group(1), 16)
    if a < lo or a > hi or any(s <= a <= e for s, e in skips):
        continue
    t = m.group(2).strip()
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    op = t.split()[0].split(".")[0]
    hist[op] += 1
    n += 1
print("total", n)
for op, c in hist.most_common():
    print(f"{c:5d} {op}")
