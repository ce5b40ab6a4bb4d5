"""Per-instruction stall samples in address order (development aid). usage: ncu_regions.py rep launch_idx [min_samples]"""
import csv, io, subprocess, sys
rep, sel = sys.argv[1], int(sys.argv[2]); mn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
kern = []
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name": kern.append({"name": r[1], "rows": []}); continue
    if r and r[0] == "Address": kern[-1]["hdr"] = r; continue
    if kern: kern[-1]["rows"].append(r)
k = kern[sel]; h = k["hdr"]; si = h.index("# Samples"); ie = h.index("Instructions Executed")
cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
tot = sum(int(r[si]) for r in k["rows"]); print(k["name"][:70], "samples", tot, "warp-instr", sum(int(r[ie]) for r in k["rows"]))
marks = ("LDTM", "UTCHMMA", "UTMALDG", "UBLKCP", "HFMA2", "HMUL2", "BAR.SYNC", "SYNCS", "STG", "STS", "LDG", "EXIT", "UTCBAR", "LDS", "F2FP")
run = 0; runi = 0
for n, r in enumerate(k["rows"]):
    s = int(r[si]); ins = r[1].strip()
    run += s; runi += int(r[ie])
    if s >= mn or any(m in ins for m in ("LDTM", "UTCHMMA", "UTMALDG", "EXIT", "BAR.SYNC", "UBLKCP")):
        st = {h[i][6:]: int(r[i]) for i in cols if int(r[i]) > 0}
        print(f"{n:5d} cum={run:6d} cumI={runi:9d} s={s:5d} ex={r[ie]:>8} {ins[:60]:60s}", sorted(st.items(), key=lambda x: -x[1])[:2])
