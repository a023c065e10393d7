"""multi-GPU check, run under torchrun (one rank per GPU):
   factor a poisson3d grid^3 system across the ranks, compare the factor bit for bit with a single-GPU
   factorization done by rank 0 on its own, solve on every rank, report times.
   usage: torchrun --nproc-per-node N scripts/gpu_mg_check.py [grid] [reps]"""
import os, sys, time, ctypes as C, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as g
import torch, torch.distributed as dist
m = g.load_package()
from meagre_crowd_b200 import stencils
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n, cp, ri, v = stencils.poisson3d_upper(grid)
s = m.Solver(device=local); s.join(rank, world, dist=dist)
S = s.analyze(n, cp, ri, 0)
b = np.arange(n, dtype=np.float64)
for it in range(reps):
    dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter(); rc = s.factor(v); tf = time.perf_counter() - t
    assert rc == 0, s.err()
    st = s.stats()
    x = s.solve(b) if rank == 0 else b; st2 = s.stats()      # the finished factor is collected on rank 0
    tt = torch.tensor([st.factor_ms], device="cuda", dtype=torch.float64); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"[world {world}] grid={grid} factor max-over-ranks={tt.item():.2f} ms ({S.flops / tt.item() / 1e9:.2f} TF/s) wall={tf * 1e3:.1f} ms solve={st2.solve_ms:.2f} ms launches={st.launches}", flush=True)
cb = C.create_string_buffer(8192); s.L.b200_get_profile(s.ctx, cb, 8192)
lv = [l for l in cb.value.decode().splitlines() if l.startswith("levels_ms")]
for r_ in range(world):
    dist.barrier()
    if r_ == rank and lv:
        t = lv[0].split(":")[1].split(); print(f"rank {rank} level ms (deepest first, root last): " + " ".join(t) + f" | sum {sum(float(x) for x in t):.1f}", flush=True)
A = stencils.to_scipy(n, cp, ri, v, True)
berr = float(np.abs(A @ x - b).max() / (12.0 * np.abs(x).max() + np.abs(b).max())) if rank == 0 else 0.0
Lmg = s.read_factor(0)
print(f"rank {rank}: backward error {berr:.2e}", flush=True)
ok = berr <= 1e-12
if rank == 0 and world > 1:
    s1 = m.Solver(device=local); s1.analyze(n, cp, ri, 0); assert s1.factor(v) == 0
    L1 = s1.read_factor(0); x1 = s1.solve(b)
    # compare the stored lower trapezoids only (entries above the diagonal of a pivot block are scratch)
    ns = S.ns; stt = m.sym_array(s.S, "sn_start", ns + 1, np.int32); rp = m.sym_array(s.S, "rowptr", ns + 1, np.int64); lp = m.sym_array(s.S, "lptr", ns + 1, np.int64)
    nbad = 0; maxd = 0.0
    for q in range(ns):
        k = stt[q + 1] - stt[q]; f = rp[q + 1] - rp[q]
        a = m.panel_view(Lmg, lp[q], f, k); c = m.panel_view(L1, lp[q], f, k)
        tri = np.tril(np.ones((f, k), dtype=bool))
        dd = np.abs(a - c)[tri]
        if dd.size and dd.max() > 0: nbad += 1; maxd = max(maxd, float(dd.max()))
    print(f"factor vs single GPU: supernodes that differ {nbad}/{ns}, max |diff| {maxd:.3e}; x identical: {np.array_equal(x, x1)} max|dx| {np.abs(x - x1).max():.3e}", flush=True)
    ok = ok and nbad == 0
    s1.close()
s.close()
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
