"""Decode a KWS_TRACE event log (tc_net.cu trace_ev) and print per-role interval statistics and a timeline excerpt."""
import sys
import numpy as np
EV = 4096
ROLES = {0: "epi", 1: "mma", 2: "raw", 3: "prod"}
NAMES = {0: {1: "wait_acc_full", 2: "got_acc_full", 3: "done"},
         1: {1: "wait_acc_empty", 2: "got_acc_empty", 3: "got_a_full", 4: "issued", 5: "fenced", 6: "mma_issued", 7: "committed"},
         2: {1: "wait_stage_empty", 2: "got_stage_empty"},
         3: {1: "wait_raw_full", 2: "got_raw_full", 3: "fir_done", 4: "got_a_empty", 5: "arrived"}}
d = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, EV)
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (200, 260)
t0 = None
allev = []
for r, name in ROLES.items():
    x = d[r]; x = x[x != 0]
    ev = (x >> np.uint64(56)).astype(int); idx = ((x >> np.uint64(40)) & np.uint64(0xffff)).astype(int)
    clk = (x & np.uint64((1 << 40) - 1)).astype(np.int64)
    if len(clk) == 0: continue
    print(f"== {name}: {len(clk)} events, span {clk[-1]-clk[0]} clk")
    # mean interval from each event to the next one of the same role, keyed by (ev -> next ev)
    stats = {}
    for i in range(len(clk) - 1):
        if i < 50: continue            # skip the ramp-up
        k = (ev[i], ev[i + 1]); stats.setdefault(k, []).append(clk[i + 1] - clk[i])
    for k, v in sorted(stats.items()):
        v = np.array(v)
        print(f"   {NAMES[r].get(k[0],k[0]):>18s} -> {NAMES[r].get(k[1],k[1]):<18s} n={len(v):5d} mean={v.mean():8.1f} med={np.median(v):8.1f} p90={np.percentile(v,90):8.1f}")
    for e, i, c in zip(ev, idx, clk): allev.append((c, name, NAMES[r].get(e, e), i))
allev.sort()
base = allev[0][0]
print("== timeline excerpt")
sel = [a for a in allev if a[1] == "mma" and a[2] == "got_a_full"]
if len(sel) > hi:
    c0, c1 = sel[lo][0], sel[hi][0]
    for c, nm, e, i in allev:
        if c0 <= c <= c1: print(f"{c-c0:8d} {nm:6s} {e:18s} {i}")
