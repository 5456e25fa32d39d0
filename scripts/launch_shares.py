"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import csv
import re
import sys

lines = open(sys.argv[1]).read().splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = list(csv.DictReader(lines[start:]))
seq = []
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    n = r["Kernel Name"]
    k = ("attn" if "attention" in n else "ln" if "layernorm" in n else "resid" if "Resid" in n else
         "gelu" if "EpiBiasBf16<(bool)1" in n else "qkv" if "EpiBiasBf16" in n else
         re.sub(r"\(.*", "", n.replace("void ", "").replace("css::", ""))[:40])
    seq.append((k, float(r["Metric Value"]) / 1e3))
tot, cnt = {}, {}
for k, v in seq:
    tot[k] = tot.get(k, 0) + v
    cnt[k] = cnt.get(k, 0) + 1
T = sum(tot.values())
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k:42s} n={cnt[k]:4d} avg={v / cnt[k]:8.1f} us share={v / T * 100:5.1f}%")
print("total ms", T / 1e3)
print("first launches:", [(k, round(v)) for k, v in seq[:16]])
