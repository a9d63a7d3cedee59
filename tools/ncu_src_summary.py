"""Summarise an `ncu --page source --csv` export: stall-reason totals and the hottest SASS lines."""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for r in data:
    for s in stalls:
        try: tot[s] += float(r[col[s]] or 0)
        except ValueError: pass
allsum = sum(tot.values())
print("stall totals (% of samples):", {k: round(100 * v / allsum, 1) for k, v in tot.most_common(9)})
ins = sum(float(r[col["Instructions Executed"]] or 0) for r in data)
print("instructions executed:", int(ins), " samples:", int(allsum))
key = "# Samples"
data.sort(key=lambda r: -float(r[col[key]] or 0))
for r in data[:top]:
    top_stall = max(stalls, key=lambda s: float(r[col[s]] or 0))
    print(f"{r[col['Address']][-5:]} {float(r[col[key]]):7.0f} {top_stall[6:]:14s} {r[col['Source']][:90]}")

# samples by opcode (where do warps sit)
byop = collections.Counter(); byop_n = collections.Counter()
for r in data:
    src = r[col['Source']].split()
    op = next((t for t in src if t[0].isalpha() and not t.startswith('@')), '?').split('.')[0]
    byop[op] += float(r[col[key]] or 0); byop_n[op] += float(r[col["Instructions Executed"]] or 0)
print("samples by opcode:", [(k, int(v), f"{100*v/allsum:.1f}%", f"n={byop_n[k]/ins*100:.1f}%") for k, v in byop.most_common(16)])
