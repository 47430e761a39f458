#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` report of tools/prof_split.py: DRAM bytes (read + write) per knot and per launch of
each FD kernel — what bench.py's `roofline.traffic` is scaled from.   usage: ncu_traffic.py report.ncu-rep nknots out.json "source note" """
import csv, io, json, subprocess, sys
rep, nknots, out, note = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U = rows[0], rows[1]
kn, rd, wr = H.index("Kernel Name"), H.index("dram__bytes_read.sum"), H.index("dram__bytes_write.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
per = {}
for r in rows[2:]:
    if len(r) < len(H):
        continue
    name = r[kn].split("<")[0].split("(")[0].replace("void ", "").replace("ilqg::", "").strip()
    b = float(r[rd]) * scale[U[rd]] + float(r[wr]) * scale[U[wr]]
    per.setdefault(name, []).append(b)
doc = {"dram_bytes_per_knot": {k: sum(v) / len(v) / nknots for k, v in per.items()}, "knots_in_capture": nknots, "source": note}
json.dump(doc, open(out, "w"), indent=1)
print(json.dumps(doc, indent=1))
