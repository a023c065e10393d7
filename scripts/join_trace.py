"""join a per-launch GPU trace (B200_TRACE_ALL=1 python scripts/gpu_trace.py grid 2> trace.log) with the device-free schedule of the same
problem (K and tile count of every update launch): where does the factorization time go, by K class, by tile count, by tree level.
usage: python scripts/join_trace.py trace.log [grid]"""
import sys, os, re, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
m = g.load_package()
import sched_sim
from meagre_crowd_b200 import stencils

log = sys.argv[1]; grid = int(sys.argv[2]) if len(sys.argv) > 2 else 128
tr = []
for l in open(log):
    mt = re.match(r"launch (\S+) depth=(\d+) count=(\d+) stream=(\d+) ms=([\d.]+) flops=([\d.e+]+)", l)
    if mt: tr.append((mt.group(1), int(mt.group(2)), int(mt.group(3)), float(mt.group(5)), float(mt.group(6))))
n, cp, ri, v = stencils.poisson3d_upper(grid)
S = m.analyze(n, cp, ri)
ls, ms = sched_sim.export(m, S, 1, 0)
ls = [l for l in ls]
assert len(ls) == len(tr), (len(ls), len(tr))
tot = sum(t[3] for t in tr)
print(f"{len(tr)} launches, {tot:.1f} ms serialised")
byk = collections.defaultdict(lambda: [0, 0.0, 0.0])
bylev = collections.defaultdict(lambda: collections.defaultdict(float))
for l, t in zip(ls, tr):
    assert l["name"] == t[0] and l["count"] == t[2], (l, t)
    bylev[t[1]][t[0]] += t[3]
    if t[0] in ("gemm_dmma", "trsm_dmma"):
        K = l["kavg"]; kc = 64 if K <= 64 else 128 if K <= 128 else 192 if K <= 192 else 256 if K <= 256 else 384 if K <= 384 else 512
        tc = "<148" if t[2] < 148 else "<444" if t[2] < 444 else "<1332" if t[2] < 1332 else "<4440" if t[2] < 4440 else ">=4440"
        a = byk[(t[0], kc, tc)]; a[0] += 1; a[1] += t[3]; a[2] += l["flops"]
print("update launches by (kind, K class, tile count class): launches, ms, stored-tile TFLOP/s")
for k in sorted(byk):
    a = byk[k]; print(f"  {k[0]:10s} K<={k[1]:3d} tiles{k[2]:7s} {a[0]:5d} {a[1]:8.2f} ms {a[2] / (a[1] * 1e-3) / 1e12 if a[1] > 0 else 0:6.2f}")
print("per level (depth: total ms | kinds)")
for d in sorted(bylev, reverse=True):
    x = bylev[d]; print(f"  d={d:2d} {sum(x.values()):8.2f} | " + " ".join(f"{k}={v_:.2f}" for k, v_ in sorted(x.items(), key=lambda z: -z[1])))
