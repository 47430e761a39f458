#!/usr/bin/env python
"""Dynamic instruction / stall-sample breakdown of one kernel of an ncu report by SOURCE REGION.
Joins `ncu --page source --csv` (per-SASS-instruction counters, in program order) with `nvdisasm -gi`
of the shipped cubin (per-instruction inline chains; needs -lineinfo).  Each instruction is attributed
to (kernel line, line inside the first-level callee) so that a fully inlined pipeline still splits by stage.
usage: ncu_lines.py report.ncu-rep kernel-regex mangled-substring [depth]"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict

rep, kre, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
depth = int(sys.argv[4]) if len(sys.argv) > 4 else 2
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("ILQG_LIB") or os.path.join(root, "ilqg-mujoco_b200", "libilqg_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.startswith("ilqg.") and f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
# locate the function's text section
start = next(i for i, l in enumerate(dis) if l.startswith("\t.section\t.text.") and mangled in l)
chains, chain, cur = [], [], []
fre = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
ire = re.compile(r'^\s+/\*([0-9a-f]+)\*/\s+(\S.*);')
for l in dis[start + 1:]:
    if l.startswith("\t.section") or l.startswith("//-----"):
        break
    m = fre.search(l)
    if m:
        cur.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = ire.match(l)
    if m:
        if cur:
            chain = cur
        cur = []
        chains.append((int(m.group(1), 16), m.group(2), chain))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hi]
body = []
for r in rows[hi + 1:]:
    if r and r[0] == "Kernel Name":
        break   # first matching launch only
    if len(r) >= len(H) - 2:
        body.append(r)
ci = {k: H.index(k) for k in ("Instructions Executed", "# Samples", "stall_long_sb", "stall_wait", "stall_no_inst", "stall_short_sb", "stall_math", "stall_barrier", "stall_branch_resolving")}
n = min(len(body), len(chains))
if len(body) != len(chains):
    print(f"# warning: {len(body)} profiled instructions vs {len(chains)} disassembled", file=sys.stderr)
agg = defaultdict(lambda: defaultdict(float))
tot = defaultdict(float)
for (off, ins, ch), r in zip(chains[:n], body[:n]):
    key = tuple(reversed(ch[-depth:])) if ch else (("?", 0),)
    isf64 = ins.split()[0].startswith(("DFMA", "DADD", "DMUL", "DSETP", "MUFU.RCP64H", "MUFU.RSQ64H")) or (ins.startswith("@") and len(ins.split()) > 1 and ins.split()[1].startswith(("DFMA", "DADD", "DMUL")))
    for k, i in ci.items():
        try:
            v = float(r[i])
        except ValueError:
            v = 0
        agg[key][k] += v
        tot[k] += v
    if isf64:
        agg[key]["f64"] += float(r[ci["Instructions Executed"]] or 0)
        tot["f64"] += float(r[ci["Instructions Executed"]] or 0)
    agg[key]["static"] += 1
src = {}
def line_text(f, ln):
    p = os.path.join(root, "ilqg-mujoco_b200", "csrc", f)
    if p not in src:
        try:
            src[p] = open(p).read().split("\n")
        except OSError:
            src[p] = []
    return src[p][ln - 1].strip()[:70] if 0 < ln <= len(src[p]) else ""
print(f"kernel {kre}: {int(tot['Instructions Executed'])} warp instructions, {int(tot['# Samples'])} samples, fp64 share {tot['f64'] / max(1, tot['Instructions Executed']):.2f}\n")
print("| region (outermost lines) | static | inst % | fp64 % of region | samples % | long_sb | wait | no_inst | source |")
print("|---|---|---|---|---|---|---|---|---|")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"]):
    if a["# Samples"] < 0.004 * tot["# Samples"] and a["Instructions Executed"] < 0.004 * tot["Instructions Executed"]:
        continue
    name = " < ".join(f"{f}:{ln}" for f, ln in key)
    s = max(1.0, a["# Samples"])
    print(f"| {name} | {int(a['static'])} | {100 * a['Instructions Executed'] / tot['Instructions Executed']:.1f} | {100 * a['f64'] / max(1, a['Instructions Executed']):.0f} | "
          f"{100 * a['# Samples'] / tot['# Samples']:.1f} | {100 * a['stall_long_sb'] / s:.0f} | {100 * a['stall_wait'] / s:.0f} | {100 * a['stall_no_inst'] / s:.0f} | `{line_text(*key[-1])}` |")
