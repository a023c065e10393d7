"""Host-side logic of the multi-GPU factorization (csrc/host/b200_plan.c), CPU only:
the plan every rank derives on its own must be a partition of the work, groups must nest along the
tree, and the contribution-block transfers must pair up.  A world-size-2 gloo run checks that two
PROCESSES derive bit-identical plans and matching send/receive lists without talking to each other
(the property the NCCL schedule relies on: src/solver_mumps.c:44-51 leaves all of this to MUMPS)."""
import ctypes as C
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest
from helpers import sym_dict
from meagre_crowd_b200 import stencils

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def plan_arrays(mc, S, world, min_front=64):
    L = mc.lib()
    pl = L.b200_plan_create(S, world, min_front)
    P = pl.contents; ns = P.ns
    out = {k: np.ctypeslib.as_array(getattr(P, k), (ns,)).copy() for k in ("grp_first", "grp_size", "dist", "work")}
    out["level_width"] = np.ctypeslib.as_array(P.level_width, (P.nlevels,)).copy()
    return pl, out


def transfers(mc, pl, S, depth):
    L = mc.lib()
    n = L.b200_plan_cb_transfers(pl, S, depth, None, 0)
    buf = (mc.b200_xfer_t * max(n, 1))()
    L.b200_plan_cb_transfers(pl, S, depth, buf, n)
    return [(buf[i].src, buf[i].dst, buf[i].sn, buf[i].offset, buf[i].count) for i in range(n)]


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("gen", [lambda: stencils.poisson3d_upper(14), lambda: stencils.poisson2d_upper(40, 23)])
def test_plan_is_a_nested_partition(mc, world, gen):
    n, cp, ri, v = gen()
    S = mc.analyze(n, cp, ri)
    d = sym_dict(mc, S); ns = d["ns"]
    pl, P = plan_arrays(mc, S, world)
    gf, gs, di = P["grp_first"], P["grp_size"], P["dist"]
    assert np.all(gs >= 1) and np.all(gf >= 0) and np.all(gf + gs <= world)
    par = d["sn_parent"]
    for s in range(ns):                       # groups nest: a child's group lies inside its parent's
        if par[s] >= 0:
            assert gf[par[s]] <= gf[s] and gf[s] + gs[s] <= gf[par[s]] + gs[par[s]]
        else:
            assert world == 1 or gs[s] >= 1
        if gs[s] == 1:
            assert di[s] == 0
    roots = np.where(par < 0)[0]
    if len(roots) == 1:
        assert gf[roots[0]] == 0 and gs[roots[0]] == world
    # subtree work is the sum over the subtree
    k = np.diff(d["sn_start"]).astype(np.float64); f = np.diff(d["rowptr"]).astype(np.float64)
    own = k * f * f - k * k * f + k ** 3 / 3.0
    acc = own.copy()
    for s in range(ns):
        if par[s] >= 0:
            acc[par[s]] += acc[s]
    assert np.allclose(acc, P["work"], rtol=1e-12)
    # every front column has exactly one owner, and that owner is a member of the supernode's group
    L = mc.lib()
    for s in list(range(0, ns, max(1, ns // 40))) + [ns - 1]:
        fs = int(f[s])
        owners = np.array([L.b200_plan_col_owner(pl, S, s, c) for c in range(fs)])
        assert np.all((owners >= gf[s]) & (owners < gf[s] + gs[s]))
        members = [r for r in range(world) if L.b200_plan_member(pl, s, r)]
        assert set(owners.tolist()) <= set(members)
        if di[s]:
            W = P["level_width"][d["sn_depth"][s]]
            assert np.array_equal(owners, gf[s] + (np.arange(fs) // W) % gs[s])
        else:
            assert members == [gf[s]]
    L.b200_plan_free(pl); L.b200_symbolic_free(S)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_contribution_block_transfers_cover_exactly_what_is_missing(mc, world):
    n, cp, ri, v = stencils.poisson3d_upper(16)
    S = mc.analyze(n, cp, ri)
    d = sym_dict(mc, S); ns = d["ns"]; L = mc.lib()
    pl, P = plan_arrays(mc, S, world)
    st, rp, par, dep, rel, cbp = d["sn_start"], d["rowptr"], d["sn_parent"], d["sn_depth"], d["relidx"], d["cbptr"]
    total = 0
    for depth in range(d["maxdepth"] + 1):
        xs = transfers(mc, pl, S, depth)
        got = {}
        for (src, dst, sn, off, cnt) in xs:
            assert src != dst and par[sn] >= 0 and dep[par[sn]] == depth
            kc = st[sn + 1] - st[sn]; uc = (rp[sn + 1] - rp[sn]) - kc; ldu = (uc + 1) & ~1
            assert (off - cbp[sn]) % ldu == 0 and cnt % ldu == 0
            j0 = (off - cbp[sn]) // ldu
            for j in range(j0, j0 + cnt // ldu):
                assert (sn, j) not in got
                got[(sn, j)] = (src, dst)
            total += cnt
        # reference: every child column whose computing rank differs from the rank that assembles it
        for c in range(ns):
            p = par[c]
            if p < 0 or dep[p] != depth:
                continue
            kc = st[c + 1] - st[c]; uc = (rp[c + 1] - rp[c]) - kc
            for j in range(uc):
                src = L.b200_plan_col_owner(pl, S, c, int(kc + j)); dst = L.b200_plan_col_owner(pl, S, int(p), int(rel[rp[c] + kc + j]))
                if src != dst:
                    assert got.pop((c, j)) == (src, dst)
        assert not got
    assert total > 0
    L.b200_plan_free(pl); L.b200_symbolic_free(S)


WORKER = r'''
import os, sys, hashlib, numpy as np, ctypes as C
sys.path.insert(0, {root!r})
import torch.distributed as dist, torch
import __graft_entry__ as g
m = g.load_package(); L = m.lib()
from meagre_crowd_b200 import stencils
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n, cp, ri, v = stencils.poisson3d_upper(16)
S = m.analyze(n, cp, ri, threads=1 + rank)          # different host thread counts: the analysis must not care
pl = L.b200_plan_create(S, world, 64); P = pl.contents
h = hashlib.sha256()
for k in ("grp_first", "grp_size", "dist"):
    h.update(np.ctypeslib.as_array(getattr(P, k), (P.ns,)).tobytes())
sends, recvs = [], []
for depth in range(S.contents.maxdepth + 1):
    nx = L.b200_plan_cb_transfers(pl, S, depth, None, 0)
    buf = (m.b200_xfer_t * max(nx, 1))(); L.b200_plan_cb_transfers(pl, S, depth, buf, nx)
    for i in range(nx):
        t = (depth, buf[i].src, buf[i].dst, buf[i].offset, buf[i].count)
        if buf[i].src == rank: sends.append(t)
        if buf[i].dst == rank: recvs.append(t)
digest = np.frombuffer(h.digest(), dtype=np.uint8).copy()
all_d = [torch.zeros(32, dtype=torch.uint8) for _ in range(world)]
dist.all_gather(all_d, torch.from_numpy(digest))
assert all(torch.equal(all_d[0], x) for x in all_d), "ranks derived different plans"
# my sends to the peer must be the peer's receives from me, in the same order
peer = 1 - rank
mine = [t for t in sends if t[2] == peer]
obj = [None, None]; dist.all_gather_object(obj, [t for t in recvs if t[1] == peer])
assert obj[peer] == mine, "send/recv lists do not pair up"
assert len(mine) > 0
print("rank", rank, "ok", len(mine), "messages to the peer")
dist.destroy_process_group()
'''


def test_two_processes_derive_the_same_plan_and_matching_messages(mc, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


WORKER2 = r'''
import os, sys, hashlib, numpy as np
sys.path.insert(0, {root!r})
import torch.distributed as dist, torch
import __graft_entry__ as g
m = g.load_package(); L = m.lib()
from meagre_crowd_b200 import stencils
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n, cp, ri, v = stencils.poisson3d_upper(20)
perm = m.shared_ordering(n, cp, ri, m.KIND_CHOLESKY, rank, world, dist, threads=2)
assert sorted(perm.tolist()) == list(range(n))
S = m.analyze(n, cp, ri, perm=perm, threads=1 + rank)
s = S.contents
h = hashlib.sha256()
for name, ln, dt in (("perm", n, np.int32), ("sn_start", s.ns + 1, np.int32), ("sn_parent", s.ns, np.int32), ("rowptr", s.ns + 1, np.int64), ("lptr", s.ns + 1, np.int64), ("cbptr", s.ns, np.int64)):
    h.update(m.sym_array(S, name, ln, dt).tobytes())
h.update(m.sym_array(S, "rowidx", int(m.sym_array(S, "rowptr", s.ns + 1, np.int64)[-1]), np.int32).tobytes())
d = [torch.zeros(32, dtype=torch.uint8) for _ in range(world)]
dist.all_gather(d, torch.from_numpy(np.frombuffer(h.digest(), dtype=np.uint8).copy()))
assert all(torch.equal(d[0], x) for x in d), "ranks built different symbolic structures from the shared ordering"
print("rank", rank, "ok", s.ns, s.nnzL)
dist.destroy_process_group()
'''


def test_shared_ordering_gives_identical_analyses_on_every_rank(mc, tmp_path):
    """bench.py at N > 1: rank 0 orders, broadcasts the permutation, every rank analyses the given order"""
    script = tmp_path / "worker2.py"
    script.write_text(WORKER2.format(root=ROOT))
    env = dict(os.environ, OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29535", str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2
