"""Sum an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel."""
import csv, sys, json, collections
rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 5]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in rows[1:]:
    if r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)  # -> us
    name = r[ik].split("(")[0]
    tot[name] += v
    cnt[name] += 1
all_us = sum(tot.values())
out = {k: {"launches": cnt[k], "total_us": round(v, 1), "mean_us": round(v / cnt[k], 2), "share": round(v / all_us, 4)}
       for k, v in sorted(tot.items(), key=lambda kv: -kv[1])}
print(json.dumps(out, indent=1))
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
