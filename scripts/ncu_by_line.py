#!/usr/bin/env python
"""Per-source-line dynamic instruction counts and stall samples for one kernel of an .ncu-rep.

ncu's CSV source page is SASS-only; this joins it (by instruction order) with `nvdisasm -g` line info of the
same build:  python scripts/ncu_by_line.py gpurun_out/prof.ncu-rep k_simplexILi1ELi1ELb1 <pivots>
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, mangled, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "yalps_b200", "libyalps_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
dis = ""  # the library has one cubin per translation unit: search them all
for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
    dis += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout + "\n"
lines, cur, on = [], None, False
for ln in dis.split("\n"):
    if ln.startswith("//---") and ".text." in ln:
        on = mangled in ln
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ii, si, wi = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared")
data = []
for r in rows[2:]:
    try:
        data.append((int(r[ii] or 0), int(r[si] or 0), int(r[wi] or 0)))
    except (ValueError, IndexError):
        pass
assert len(data) == len(lines), (len(data), len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0])
for ln, (ie, sm, wf) in zip(lines, data):
    a = agg[ln]
    a[0] += ie
    a[1] += sm
    a[2] += wf
ti, ts, tw = (sum(a[k] for a in agg.values()) for k in range(3))
print(f"total: {ti / units:.1f} instr/unit, {tw / units:.1f} smem wavefronts/unit, {ts} samples")
src = {}
for (f, l), (ie, sm, wf) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if sm < ts * 0.004 and ie < ti * 0.004:
        continue
    if f not in src:
        p = os.path.join(root, "yalps_b200", "csrc", f)
        src[f] = open(p).read().split("\n") if os.path.exists(p) else []
    text = src[f][l - 1].strip()[:90] if l - 1 < len(src[f]) else ""
    print(f"{100 * sm / ts:5.1f}% smp {ie / units:7.1f} ins {wf / units:6.1f} wf  {f}:{l}  {text}")
