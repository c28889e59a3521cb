"""Where a tick kernel's time goes, phase by phase: joins an ncu source-page export (per-SASS-instruction stall samples and
execution counts, `ncu --set full --import-source on`) with the line table of the library (nvdisasm -g) and groups the
instructions by the phase of physics_tick they belong to (the source lines of the phase markers in qs_physics.cuh).

    ncu -i X.ncu-rep --page source --csv --print-source sass > /tmp/src.csv
    python tools/phase_profile.py /tmp/src.csv k_settle_slice [launch index, default 0] > profiles/rNN_slice_phase_profile.json
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src_csv, kernel = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
so = os.path.join(ROOT, "quadruped_springs_b200", "csrc", "libqs_b200.so")
phys = os.path.join(ROOT, "quadruped_springs_b200", "csrc", "qs_physics.cuh")

# phase boundaries = the marker comments of physics_tick (first instantiation in the file = the fast tick)
markers = [("ctx+trunk", "__host__ __device__ int physics_tick("), ("pass A (leg dynamics)", "// ---- pass A"), ("base solve", "// ---- base: S = "),
           ("velocity update", "// ---- v += dt a"), ("contact rows", "// ---- contact rows of the active feet"),
           ("PGS sweeps", "// ---- projected Gauss-Seidel"), ("impulse application", "// ---- apply: base twist change"),
           ("integration", "  cs.mask = active;")]
lines = open(phys).read().split("\n")
bounds = []
for name, text in markers:
    ln = next(i + 1 for i, l in enumerate(lines) if text in l)
    bounds.append((ln, name))
end_tick = next(i + 1 for i, l in enumerate(lines) if "return TICK_DONE;" in l)


def phase_of(line):
    if line < bounds[0][0] or line > end_tick:
        return None
    cur = None
    for ln, name in bounds:
        if line >= ln:
            cur = name
    return cur


with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
info, cur, on = {}, None, False
for l in dis.split("\n"):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        on = kernel in m.group(1) and "Lb0" in m.group(1)
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        info[int(m.group(1), 16)] = cur

rows = list(csv.reader(open(src_csv)))
sections, cur_rows, name = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        name = r[1]
        cur_rows = []
        sections.append((name, cur_rows))
    elif cur_rows is not None:
        cur_rows.append(r)
sel = [s for s in sections if kernel in s[0]]
# ncu prints every launch twice (two views); take distinct launches in order
name, body = sel[2 * which] if len(sel) > 2 * which else sel[which]
hdr = body[0]
ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not" not in h]
data = [r for r in body[1:] if len(r) > iex]
base = int(data[0][ia], 16)
agg = collections.OrderedDict()
last = "outside the tick (torques, loop, state i/o)"
tot_s = tot_e = 0
for r in data:
    fl = info.get(int(r[ia], 16) - base)
    if fl and fl[0] == "qs_physics.cuh":
        p = phase_of(fl[1])
        if p:
            last = p
    elif fl and fl[0] == "qs_step_kernels.cuh":
        last = "outside the tick (torques, loop, state i/o)"
    a = agg.setdefault(last, {"static_instructions": 0, "executed": 0, "samples": 0, "stalls": collections.Counter()})
    a["static_instructions"] += 1
    a["executed"] += int(r[iex] or 0)
    a["samples"] += int(r[isamp] or 0)
    for i in stall_cols:
        a["stalls"][hdr[i][6:]] += int(r[i] or 0)
    tot_s += int(r[isamp] or 0)
    tot_e += int(r[iex] or 0)
bar = [int(r[iex]) for r in data if "BAR.SYNC" in r[hdr.index("Source")]]
ticks = max(bar) if bar else 1
out = {"kernel": name, "how": __doc__.split("\n")[0], "warp_ticks": ticks, "instructions_per_warp_tick": round(tot_e / ticks, 1),
       "note": "instructions of inlined helpers count for the phase whose marker line was seen last: the split is approximate",
       "phases": {}}
for k, a in agg.items():
    st = sum(a["stalls"].values()) or 1
    out["phases"][k] = {"static_instructions": a["static_instructions"], "executed_per_warp_tick": round(a["executed"] / ticks, 1),
                        "share_of_executed": round(a["executed"] / tot_e, 4), "share_of_samples": round(a["samples"] / tot_s, 4),
                        "top_stalls_pct": {n: round(100 * v / st, 1) for n, v in a["stalls"].most_common(4)}}
print(json.dumps(out, indent=1))
