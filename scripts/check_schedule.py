"""CPU check of the multi-GPU factor schedule of a poisson3d grid^3 problem: build every rank's schedule without a device
and replay all ranks together; then check that every message is ordered against the launches touching the same data (tests/sched_sim.py).  usage: python scripts/check_schedule.py [grid] [world ...]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
m = g.load_package()
import sched_sim
from meagre_crowd_b200 import stencils
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 64
worlds = [int(a) for a in sys.argv[2:]] or [2, 4, 8]
n, cp, ri, v = stencils.poisson3d_upper(grid)
t = time.time(); S = m.analyze(n, cp, ri); s = S.contents
print(f"grid {grid}: n={n} F={s.flops:.4e} nnzL={s.nnzL:.4e} supernodes={s.ns} levels={s.maxdepth + 1} maxfront={s.maxfront} analysis {time.time() - t:.1f}s", flush=True)
for w in worlds:
    t = time.time()
    full = [sched_sim.export(m, S, w, r, with_accesses=True) for r in range(w)]
    sch = [(a, b) for a, b, c in full]
    st = sched_sim.check(sch)
    pairs = sum(sched_sim.check_hazards(ls, ms, acc, r) for r, (ls, ms, acc) in enumerate(full))
    pl = m.lib().b200_plan_create(S, w, 2048); P = pl.contents
    nd = sum(P.dist[i] for i in range(P.ns))
    print(f"world {w}: OK  launches/rank {st['launches']}  messages/rank {st['messages']}  NCCL {st['nccl_bytes'] / 1e9:.2f} GB  distributed fronts {nd}  hazard pairs checked {pairs}  ({time.time() - t:.1f}s)", flush=True)
    m.lib().b200_plan_free(pl)
