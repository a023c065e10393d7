"""debug: where do non-finite values appear in x (elimination order) for growing poisson3d grids"""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g
m = g.load_package()
from meagre_crowd_b200 import stencils
for grid in [int(a) for a in sys.argv[1:]] or [64, 96, 128]:
    n, cp, ri, v = stencils.poisson3d_upper(grid)
    s = m.Solver(); S = s.analyze(n, cp, ri, 0)
    s.L.b200_set_refinement(s.ctx, 0)
    r = s.factor(v); b = np.arange(n, dtype=np.float64)
    k_ = np.diff(m.sym_array(s.S, "sn_start", S.ns + 1, np.int32)).astype(np.int64)
    sz = (k_ // 256) * 65536 + (k_ % 256) ** 2; off = np.concatenate([[0], np.cumsum(sz)])
    for which in (2, 3, 4):
        tot = int(off[-1]) if which > 2 else int(((k_ // 64) * 4096 + (k_ % 64) ** 2).sum())
        buf = np.empty(tot); s.L.b200_debug_read(s.ctx, which, buf.ctypes.data, 0, tot)
        bad = np.where(~np.isfinite(buf))[0]
        print("  array", which, "size", tot, "nonfinite", len(bad), "absmax", np.nanmax(np.abs(buf)) if tot else 0)
        if len(bad) and which > 2:
            sn = np.searchsorted(off, bad[0], side="right") - 1
            print("   first bad at", bad[0], "supernode", sn, "k", k_[sn], "local", bad[0] - off[sn], "block", (bad[0] - off[sn]) // 65536)
    x = s.solve(b)
    perm = m.sym_array(s.S, "perm", n, np.int32); ns = S.ns
    st = m.sym_array(s.S, "sn_start", ns + 1, np.int32); dep = m.sym_array(s.S, "sn_depth", ns, np.int32); rp = m.sym_array(s.S, "rowptr", ns + 1, np.int64)
    xp = x[perm]; bad = np.where(~np.isfinite(xp))[0]
    A = stencils.to_scipy(n, cp, ri, v, True)
    berr = np.abs(A @ x - b).max() / (12 * np.abs(x).max() + np.abs(b).max())
    print(grid, "rc", r, "nonfinite", len(bad), "berr", berr, flush=True)
    if len(bad):
        sn = np.searchsorted(st, bad, side="right") - 1
        u, cnt = np.unique(sn, return_counts=True)
        for a, c in list(zip(u, cnt))[:12] + list(zip(u, cnt))[-12:]:
            print("  sn", a, "depth", dep[a], "k", st[a + 1] - st[a], "f", rp[a + 1] - rp[a], "bad cols", c, "first bad local col", bad[sn == a][0] - st[a], "last", bad[sn == a][-1] - st[a])
        print("  distinct supernodes with bad:", len(u), "depths", np.unique(dep[u]))
    s.close()
