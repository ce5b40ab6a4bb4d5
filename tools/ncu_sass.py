"""Dump the SASS of one launch from an ncu report with per-instruction samples (address order).
usage: python tools/ncu_sass.py rep.ncu-rep <launch> [min_samples_to_mark]"""
import csv, io, subprocess, sys
rep, sel = sys.argv[1], int(sys.argv[2])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kern = []
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        kern.append({"name": r[1], "rows": []}); continue
    if r and r[0] == "Address":
        kern[-1]["hdr"] = r; continue
    if kern:
        kern[-1]["rows"].append(r)
k = kern[sel]; h = k["hdr"]; si = h.index("# Samples")
for i, r in enumerate(k["rows"]):
    print(f"{i:5d} {int(r[si]):6d}  {r[1].strip()}")
