"""Opcode histogram of the hot kernels' SASS (static counts): python tools/sass_histogram.py [lib.so] > profiles/rNN_sass_opcodes.json"""
import collections
import json
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "quadruped_springs_b200/csrc/libqs_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
out, cur = {}, None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = collections.Counter()
        out[re.sub(r"\(.*", "", name)] = cur
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
        if m.group(1) in ("LDL", "STL", "LDS", "STS", "LDG", "STG", "BAR", "MUFU"):
            cur[m.group(1) + m.group(2)] += 0  # keep the base bucket only
keep = ["k_settle_slice<false>", "k_step<false>", "k_step_contact<false>", "k_step_slow<false>", "k_pre"]
res = {}
for k in out:
    short = k.replace("void ", "")
    if short in keep:
        c = out[k]
        tot = sum(c.values())
        fp = sum(c[x] for x in ("FFMA", "FFMA2", "FMUL", "FMUL2", "FADD", "FADD2", "FMNMX", "FSEL", "FSET", "FSETP", "FCHK"))
        res[short] = {"total": tot, "fp32_pipe": fp, "fp32_share": round(fp / tot, 3),
                      "top": dict(c.most_common(28)),
                      "named": {x: c.get(x, 0) for x in ("FFMA", "FFMA2", "FMUL", "FADD", "MUFU", "LDL", "STL", "LDS", "STS", "LDG", "STG", "BAR", "HMMA", "UTCHMMA")}}
print(json.dumps(res, indent=1))
