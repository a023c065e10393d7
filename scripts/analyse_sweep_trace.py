"""per-unit trace of the persistent sweeps (scripts/gpu_sweep_trace.py -> .npz): unit counts, waiting / working time per unit kind, first start .. last end per tree level, units at work per 250 us.  usage: python scripts/analyse_sweep_trace.py trace.npz"""
import numpy as np, sys
d = np.load(sys.argv[1])
names = ["FWD_GATHER", "FWD_SMALL", "GEMV", "GEMVT", "BWD_SMALL", "BATCH", "GEMV8", "GEMVT8", "COMBINE"]
depth = d["sn_depth"]
for nm in ("f", "b"):
    if "trace_" + nm in d: tr = d["trace_" + nm].astype(np.int64); un = d["units_" + nm]
    else: tr = np.concatenate([d["trace_" + nm + "c"], d["trace_" + nm + "b"]]).astype(np.int64); un = np.concatenate([d["units_" + nm + "c"], d["units_" + nm + "b"]])
    g0 = tr[:, 0]; cg, cr, cd = tr[:, 1], tr[:, 2], tr[:, 3]
    t0 = g0.min()
    start = (g0 - t0) / 1e3     # us
    dur_wait = (cr - cg) / 1965.0; dur_body = (cd - cr) / 1965.0
    end = start + (cd - cg) / 1965.0
    print(f"== sweep {nm}: {len(un)} units, span {end.max():.0f} us; sum wait {dur_wait.sum()/1e3:.1f} ms, sum body {dur_body.sum()/1e3:.1f} ms  (592 CTA x span = {592*end.max()/1e3:.1f} ms)")
    for k in range(9):
        sel = un[:, 0] == k
        if sel.any(): print(f"   {names[k]:10s} n={sel.sum():7d} wait avg {dur_wait[sel].mean():7.2f} us  body avg {dur_body[sel].mean():7.2f} med {np.median(dur_body[sel]):7.2f} max {dur_body[sel].max():8.1f}  sum body {dur_body[sel].sum()/1e3:7.2f} ms sum wait {dur_wait[sel].sum()/1e3:7.2f} ms")
    # timeline by depth: first start and last end of units per depth
    dp = depth[un[:, 4]]
    print("   depth: first start .. last end (us), units, body ms, wait ms")
    for dd in (sorted(set(dp.tolist()), reverse=(nm == "f"))):
        s = dp == dd
        print(f"     d={dd:2d} {start[s].min():8.0f} .. {end[s].max():8.0f}  n={s.sum():6d} body {dur_body[s].sum()/1e3:6.2f} wait {dur_wait[s].sum()/1e3:6.2f}")
    # concurrency over time: number of units in body per 250us bucket
    T = end.max(); nb = int(T // 250) + 1
    busy = np.zeros(nb)
    bs = start + dur_wait
    for i in range(len(un)):
        a, b = bs[i], end[i]
        ia, ib = int(a // 250), int(b // 250)
        if ia == ib: busy[ia] += b - a
        else:
            busy[ia] += (ia + 1) * 250 - a; busy[ib] += b - ib * 250
            if ib > ia + 1: busy[ia + 1: ib] += 250
    print("   avg units in body per 250us bucket:", " ".join(f"{x/250:.0f}" for x in busy))
