"""Turn an .ncu-rep into a small tracked summary (profiles/*.md): key raw metrics + stall mix + hottest lines.
usage: python tools/make_profile_summary.py <rep> <out.md> "<title>" "<command>"
"""
import csv, subprocess, sys, collections, io
rep, out, title, cmd = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
lines = [f"# {title}", "", f"Command (under gpurun, one B200): `{cmd}`", "", f"Kernel: `{vals[hdr.index('Kernel Name')]}`" if "Kernel Name" in hdr else "", "",
         "| metric | value | unit |", "|---|---|---|"]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        lines.append(f"| {w} | {vals[i]} | {units[i]} |")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h2 = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(h2)]
col = {h: i for i, h in enumerate(h2)}
stalls = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for r in data:
    for s in stalls:
        try: tot[s] += float(r[col[s]] or 0)
        except ValueError: pass
allsum = sum(tot.values()) or 1
lines += ["", "Warp-state samples: " + ", ".join(f"{k[6:]} {100*v/allsum:.1f}%" for k, v in tot.most_common(8)), ""]
byop = collections.Counter(); ins = 0.0
for r in data:
    t = r[col['Source']].split()
    op = next((x for x in t if x[0].isalpha() and not x.startswith('@')), '?').split('.')[0]
    n = float(r[col["Instructions Executed"]] or 0); byop[op] += n; ins += n
lines += ["Instruction mix (executed warp instructions): " + ", ".join(f"{k} {100*v/ins:.1f}%" for k, v in byop.most_common(14)), ""]
data.sort(key=lambda r: -float(r[col['# Samples']] or 0))
lines += ["Hottest SASS lines (samples, dominant stall):", "", "```"]
for r in data[:12]:
    ts = max(stalls, key=lambda s: float(r[col[s]] or 0))
    lines.append(f"{float(r[col['# Samples']]):7.0f} {ts[6:]:14s} {r[col['Source']][:100]}")
lines += ["```", ""]
open(out, "w").write("\n".join(lines))
print("wrote", out)
