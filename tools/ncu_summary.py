"""Summarise an ncu report: per-launch key metrics + stall breakdown of one kernel (development aid).
usage: python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep [launch index for the stall page]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = [("us", "gpu__time_duration.sum"), ("rdMB", "dram__bytes_read.sum"), ("wrMB", "dram__bytes_write.sum"),
        ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l2%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size"), ("smem", "launch__shared_mem_per_block_dynamic")]
ki = hdr.index("Kernel Name")
print("| # | kernel | " + " | ".join(w[0] for w in want) + " |")
print("|---|---|" + "---|" * len(want))
units = rows[1]
SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,                      # -> us
         "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}        # -> MB
for n, r in enumerate(rows[2:]):
    name = r[ki].split("::")[-1].split("(")[0][:28]
    if name.startswith("GemmParams") or name.startswith("DftParams") or name.startswith("FusedParams"):
        name = r[ki].split("(")[0].split("::")[-1][:28]
    vals = []
    for _, m in want:
        try:
            i = hdr.index(m)
            vals.append("%.4g" % (float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)))
        except Exception:
            vals.append("-")
    print(f"| {n} | {name} | " + " | ".join(vals) + " |")
if len(sys.argv) > 2:
    sel = int(sys.argv[2])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    kern = []
    for r in csv.reader(io.StringIO(src)):
        if r and r[0] == "Kernel Name":
            kern.append({"name": r[1], "rows": []}); continue
        if r and r[0] == "Address":
            kern[-1]["hdr"] = r; continue
        if kern:
            kern[-1]["rows"].append(r)
    k = kern[sel]; h = k["hdr"]
    si = h.index("# Samples")
    cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    tot = sum(int(r[si]) for r in k["rows"])
    print("\nstalls of launch", sel, k["name"][:60], "samples", tot)
    agg = {h[i]: sum(int(r[i]) for r in k["rows"]) for i in cols}
    print(", ".join(f"{a[6:]} {100*b/tot:.1f}%" for a, b in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    for r in sorted(k["rows"], key=lambda r: -int(r[si]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
        st = {h[i][6:]: int(r[i]) for i in cols if int(r[i]) > 0}
        print(r[si], r[1].strip()[:64], sorted(st.items(), key=lambda x: -x[1])[:2])
