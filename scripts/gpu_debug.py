import importlib.util, sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("meagre_crowd_b200", os.path.join(ROOT, "meagre-crowd_b200", "__init__.py"), submodule_search_locations=[os.path.join(ROOT, "meagre-crowd_b200")])
m = importlib.util.module_from_spec(spec); sys.modules["meagre_crowd_b200"] = m; spec.loader.exec_module(m)
from meagre_crowd_b200 import stencils
import scipy.linalg as sl
for name, gen in [("p2d-12", stencils.poisson2d_upper(12)), ("p2d-40", stencils.poisson2d_upper(40)), ("p3d-12", stencils.poisson3d_upper(12))]:
    n, cp, ri, v = gen
    A = stencils.to_scipy(n, cp, ri, v, True).toarray()
    s = m.Solver(); S = s.analyze(n, cp, ri, 0); r = s.factor(v)
    perm = m.sym_array(s.S, "perm", n, np.int32)
    PA = A[np.ix_(perm, perm)]
    Lref = np.linalg.cholesky(PA)
    Lg = s.dense_L()
    err = np.abs(Lg - Lref)
    print(name, "factor rc", r, "max|L-Lref|", err.max(), "ns", S.ns)
    if err.max() > 1e-10:
        st = m.sym_array(s.S, "sn_start", S.ns + 1, np.int32); rp = m.sym_array(s.S, "rowptr", S.ns + 1, np.int64); dep = m.sym_array(s.S, "sn_depth", S.ns, np.int32)
        par = m.sym_array(s.S, "sn_parent", S.ns, np.int32)
        bad = 0
        for sn in range(S.ns):
            cols = np.arange(st[sn], st[sn + 1])
            e = err[:, cols]
            if e.max() > 1e-10:
                rr, cc = np.unravel_index(np.argmax(e > 1e-10), e.shape)
                rows_bad = np.unique(np.nonzero(e > 1e-10)[0])
                print("  sn", sn, "depth", dep[sn], "k", len(cols), "f", rp[sn + 1] - rp[sn], "parent", par[sn], "nchild", int((par == sn).sum()), "maxerr", e.max(), "first bad local col", np.nonzero((e > 1e-10).any(axis=0))[0][:5], "bad rows (front-local)", [int(np.searchsorted(np.concatenate([cols, []]), 0))] and rows_bad[:8])
                bad += 1
                if bad > 6: break
    # solve check using reference L on host to isolate the solve kernels
    b = np.arange(n, dtype=float)
    x = s.solve(b)
    xr = np.linalg.solve(A, b)
    print("  solve max|x-xref|", np.abs(x - xr).max(), "rel", np.abs(x - xr).max() / np.abs(xr).max())
    s.close()
