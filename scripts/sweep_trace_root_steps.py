"""time line of the root-level units of a sweep trace, grouped by task: python scripts/sweep_trace_root_steps.py trace.npz f|b"""
import numpy as np, sys
d = np.load(sys.argv[1]); nm = sys.argv[2]
tr = d["trace_" + nm].astype(np.int64); un = d["units_" + nm]; depth = d["sn_depth"]
g0 = tr[:, 0]; t0 = g0.min()
start = (g0 - t0) / 1e3; ready = start + (tr[:, 2] - tr[:, 1]) / 1965.0; end = start + (tr[:, 3] - tr[:, 1]) / 1965.0
names = ["FWD_GATHER", "FWD_SMALL", "GEMV", "GEMVT", "BWD_SMALL", "BATCH", "GEMV8", "GEMVT8", "COMBINE"]
sel = np.nonzero(depth[un[:, 4]] == 0)[0]
# group consecutive units with the same (kind, task)
i = 0; rows = 0
while i < len(sel) and rows < 40:
    j = i
    while j + 1 < len(sel) and un[sel[j + 1], 0] == un[sel[i], 0] and un[sel[j + 1], 1] == un[sel[i], 1]: j += 1
    s = sel[i:j + 1]
    print(f"{names[un[s[0],0]]:8s} task {un[s[0],1]:6d} n={len(s):4d} idx {s[0]:7d}  grab {start[s].min():8.1f}..{start[s].max():8.1f}  ready {ready[s].min():8.1f}..{ready[s].max():8.1f}  end {end[s].min():8.1f}..{end[s].max():8.1f}")
    i = j + 1; rows += 1
