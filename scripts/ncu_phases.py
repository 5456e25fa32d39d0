"""Summarise the SASS page of an ncu report: top stall sites and per-marker-range aggregates.
Usage: python scripts/ncu_phases.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h, data = rows[hi], rows[hi + 1:]
isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[isamp] or 0) for r in data)
print("total samples", tot, "warp-instr", sum(int(r[iex] or 0) for r in data))
order = sorted(range(len(data)), key=lambda i: -int(data[i][isamp] or 0))[:top_n]
for i in sorted(order):
    r = data[i]
    st = sorted(((int(r[j] or 0), c) for j, c in stall), reverse=True)[:2]
    print(f"{i:5d} {r[isamp]:>6s} {r[iex]:>9s}  {r[isrc][:70]:70s} {st}")
print("--- markers")
for i, r in enumerate(data):
    s = r[isrc]
    if any(k in s for k in ["LDTM", "BAR.SYNC", "UTCBAR", "SYNCS.ARRIVE", "EXIT", "STG", "STL", "LDL"]):
        print(i, r[isamp], r[iex], s[:80])
if len(sys.argv) > 3:
    print("--- ranges")
    bounds = [int(x) for x in sys.argv[3].split(",")]
    for a, b in zip(bounds[:-1], bounds[1:]):
        t = sum(int(r[isamp] or 0) for r in data[a:b])
        ex = sum(int(r[iex] or 0) for r in data[a:b])
        st = {}
        for r in data[a:b]:
            for j, c in stall:
                st[c] = st.get(c, 0) + int(r[j] or 0)
        top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
        print(f"[{a:5d},{b:5d}) samples={t:6d} warp-instr={ex:10d} {top}")
