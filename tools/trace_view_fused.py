"""Decode a KWS_TRACE=0 event log of conv1_block1_kernel (tc_net.cu) and print per-role interval statistics."""
import sys
import numpy as np
EV = 4096
ROLES = {0: "mid", 1: "out", 2: "mma"}
NAMES = {0: {1: "wait_a2_empty", 2: "got_a2_empty", 3: "packed", 4: "bar1", 5: "fir_loaded", 6: "bar2", 7: "arrived", 8: "wait_acc1", 9: "got_acc1"},
         1: {1: "wait_acc2_full", 2: "got_acc2_full", 3: "done"},
         2: {1: "wait_acc2_empty", 2: "got_acc2_empty", 3: "got_a2_full"}}
d = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, EV)
for r, name in ROLES.items():
    x = d[r]; x = x[x != 0]
    ev = (x >> np.uint64(56)).astype(int); clk = (x & np.uint64((1 << 40) - 1)).astype(np.int64)
    if len(clk) == 0: continue
    print(f"== {name}: {len(clk)} events, span {clk[-1]-clk[0]} clk")
    stats = {}
    for i in range(50, len(clk) - 1):
        stats.setdefault((ev[i], ev[i + 1]), []).append(clk[i + 1] - clk[i])
    for k, v in sorted(stats.items()):
        v = np.array(v)
        print(f"   {NAMES[r].get(k[0],k[0]):>16s} -> {NAMES[r].get(k[1],k[1]):<16s} n={len(v):5d} mean={v.mean():8.1f} med={np.median(v):8.1f} p90={np.percentile(v,90):8.1f}")
