"""profiles/kernel_traffic.json from an ncu summary (scripts/ncu_summary.py output): DRAM bytes per launch of each kernel of a
steady-state step; bench.py copies the dominant kernel's figure into roofline.traffic (and says that it is a constant).
    python tools/kernel_traffic.py profiles/r02_step_ncu_full.json > profiles/kernel_traffic.json"""
import json
import sys

src = sys.argv[1]
rows = json.load(open(src))


def mb(x):
    v, u = x.split()[:2]
    return float(v) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


# one step = from the first k_pre up to (not including) the next one
names = [r["Kernel Name"].replace("void ", "").split("<")[0].split("(")[0] for r in rows]
first = names.index("k_pre") if "k_pre" in names else 0
last = names.index("k_pre", first + 1) if names.count("k_pre") > 1 else len(rows)
if last - first < 4:   # the capture started mid-step: take the launches before the first k_pre as well
    first, last = 0, (names.index("k_pre") + 1 if "k_pre" in names else len(rows))
rows = rows[first:last]
out, order = {}, []
for r in rows:
    name = r["Kernel Name"].replace("void ", "").split("<")[0].split("(")[0]
    b = mb(r["dram__bytes_read.sum"]) + mb(r["dram__bytes_write.sum"])
    out.setdefault(name, []).append(b)
    order.append(name)
res = {}
for k, v in out.items():
    res[k] = {"dram_bytes_per_launch": sum(v), "launches_per_step": len(v),
              "note": "sum of the launches of one step" if len(v) > 1 else "",
              "source": f"{src} (ncu --set full, the launches of one steady-state step; cold caches under ncu)"}
step = sum(sum(out.get(k, [0])) for k in ("k_pre", "k_step", "k_step_contact", "k_finish"))
res["k_step + k_step_contact"] = {"dram_bytes_per_launch": step, "launches_per_step": 4,
                                  "note": "k_pre + flight + contact + k_finish launches of one step", "source": res[order[0]]["source"]}
print(json.dumps(res, indent=1))
