"""Deadlock-freedom and consistency of the multi-GPU factor schedule, on the CPU: every rank's static schedule is built
without a device (b200_ctx_create_dry) and all ranks are replayed together (tests/sched_sim.py)."""
import os

import pytest
import sched_sim
from meagre_crowd_b200 import stencils


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("name,gen,min_front", [("p3d-18", lambda: stencils.poisson3d_upper(18), 96), ("p3d-26", lambda: stencils.poisson3d_upper(26), 256),
                                                ("p2d-90x60", lambda: stencils.poisson2d_upper(90, 60), 48), ("p3d-30x14x9", lambda: stencils.poisson3d_upper(30, 14, 9), 64)])
def test_schedules_of_all_ranks_replay_without_deadlock(mc, world, name, gen, min_front):
    n, cp, ri, v = gen()
    S = mc.analyze(n, cp, ri)
    old = os.environ.get("B200_MIN_DIST_FRONT")
    os.environ["B200_MIN_DIST_FRONT"] = str(min_front)     # small fronts distributed too, so that every message kind occurs
    try:
        sch = [sched_sim.export(mc, S, world, r) for r in range(world)]
    finally:
        if old is None:
            del os.environ["B200_MIN_DIST_FRONT"]
        else:
            os.environ["B200_MIN_DIST_FRONT"] = old
    st = sched_sim.check(sch)
    assert st["nccl_bytes"] > 0
    kinds = {l["name"] for ls, _ in sch for l in ls}
    assert "nccl_factor_gather" in kinds and "nccl_cb_columns" in kinds
    mc.lib().b200_symbolic_free(S)


def test_replay_detects_a_broken_schedule(mc):
    """the checker itself: swapping two messages on one side, or dropping a record, must be caught"""
    n, cp, ri, v = stencils.poisson3d_upper(18)
    S = mc.analyze(n, cp, ri)
    os.environ["B200_MIN_DIST_FRONT"] = "96"
    try:
        sch = [sched_sim.export(mc, S, 2, r) for r in range(2)]
    finally:
        del os.environ["B200_MIN_DIST_FRONT"]
    sched_sim.check(sch)
    import copy
    bad = copy.deepcopy(sch)
    ls, ms = bad[1]
    li = next(i for i, l in enumerate(ls) if len(l["msgs"]) >= 2 and ms[l["msgs"][0]][2:] != ms[l["msgs"][1]][2:])
    a, b = ls[li]["msgs"][0], ls[li]["msgs"][1]; ms[a], ms[b] = ms[b], ms[a]
    with pytest.raises(AssertionError):
        sched_sim.check(bad)
    bad = copy.deepcopy(sch)
    ls, ms = bad[0]
    li = next(i for i, l in enumerate(ls) if l["rec"] >= 0 and any(k["wait"] == l["rec"] for k in ls))
    ls[li]["rec"] = -1
    with pytest.raises(AssertionError):
        sched_sim.check(bad)
    mc.lib().b200_symbolic_free(S)


@pytest.mark.parametrize("replicate", [False, True])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_every_rank_ends_up_with_the_whole_factor(mc, world, replicate, monkeypatch):
    """the sweeps run on rank 0 (default) or replicated (B200_REPLICATE_FACTOR=1), so after the factorization rank 0 / every rank must
    hold every column of L and every 64x64 inverse block: either it computed them (b200_plan_col_owner) or a message delivered them.
    Checked from the exported messages; without replication no other rank may receive bulk factor traffic it does not need."""
    if replicate:
        monkeypatch.setenv("B200_REPLICATE_FACTOR", "1")
    else:
        monkeypatch.delenv("B200_REPLICATE_FACTOR", raising=False)
    import numpy as np
    from helpers import sym_dict
    n, cp, ri, v = stencils.poisson3d_upper(22)
    S = mc.analyze(n, cp, ri)
    d = sym_dict(mc, S); ns = d["ns"]; L = mc.lib()
    os.environ["B200_MIN_DIST_FRONT"] = "128"
    try:
        sch = [sched_sim.export(mc, S, world, r) for r in range(world)]
        pl = L.b200_plan_create(S, world, 128)
    finally:
        del os.environ["B200_MIN_DIST_FRONT"]
    sched_sim.check(sch)
    st, rp, lp = d["sn_start"], d["rowptr"], d["lptr"]
    k = np.diff(st); f = np.diff(rp); ld = (f + 1) & ~1
    nblk = (k + 63) // 64; blk0 = np.concatenate([[0], np.cumsum(nblk)])
    invsz = (k // 64) * 4096 + (k % 64) ** 2; invp = np.concatenate([[0], np.cumsum(invsz)])
    assert np.array_equal(lp[1:] - lp[:-1], ld * k)
    for r in (range(world) if replicate else [0]):
        have = np.zeros(n, dtype=bool); have_inv = np.zeros(blk0[-1], dtype=bool)
        for s in range(ns):
            for c in range(k[s]):
                if L.b200_plan_col_owner(pl, S, s, int(c)) == r:
                    have[st[s] + c] = True; have_inv[blk0[s] + c // 64] = True
        for (src, dst, which, off, cnt) in sch[r][1]:
            if dst != r or which > 1:
                continue
            if which == 0:
                s = int(np.searchsorted(lp, off, side="right") - 1)
                while cnt > 0:
                    rel = off - lp[s]
                    assert rel % ld[s] == 0, "L message does not start at a column"
                    c0 = rel // ld[s]; nc = min(k[s] - c0, cnt // ld[s])
                    assert nc > 0 and (cnt >= ld[s] * nc)
                    have[st[s] + c0: st[s] + c0 + nc] = True
                    off += nc * ld[s]; cnt -= nc * ld[s]; s += 1
            else:
                s = int(np.searchsorted(invp, off, side="right") - 1)
                while cnt > 0:
                    rel = off - invp[s]
                    assert rel % 4096 == 0, "inverse message does not start at a block"
                    b0 = rel // 4096
                    take = min(invsz[s] - rel, cnt)
                    nb = (take + 4095) // 4096
                    have_inv[blk0[s] + b0: blk0[s] + b0 + nb] = True
                    off += take; cnt -= take; s += 1
        assert have.all(), f"rank {r} lacks {np.count_nonzero(~have)} columns of L"
        assert have_inv.all(), f"rank {r} lacks {np.count_nonzero(~have_inv)} inverse blocks"
    L.b200_plan_free(pl); L.b200_symbolic_free(S)


def test_multi_gpu_lu_is_refused_not_silently_wrong(mc):
    """the distributed extend-add of the U' part is not built: the shim must say so instead of producing a schedule"""
    n, cp, ri, v = stencils.convdiff27_full(8)
    S = mc.analyze(n, cp, ri, mc.KIND_LU)
    L = mc.lib()
    c = L.b200_ctx_create_dry(0, 2)
    assert L.b200_upload_symbolic(c, S) == mc.B200_ERR_ARG and b"Cholesky only" in L.b200_last_error()
    L.b200_ctx_destroy(c)
    c = L.b200_ctx_create_dry(0, 1)            # one rank: LU schedules fine
    assert L.b200_upload_symbolic(c, S) == mc.B200_OK
    assert L.b200_factor_host(c, v.ctypes.data) == mc.B200_ERR_STATE      # and a schedule-only context never computes
    L.b200_ctx_destroy(c)
    L.b200_symbolic_free(S)


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("gen,min_front", [(lambda: stencils.poisson3d_upper(22), 128), (lambda: stencils.poisson2d_upper(90, 60), 48)])
def test_messages_are_ordered_against_the_launches_that_touch_the_same_data(mc, world, gen, min_front):
    """no race between NCCL traffic and compute: happens-before (stream order + events) must relate every message to every
    launch of its rank whose conservative access intervals overlap the message's range; arriving data precedes its readers"""
    n, cp, ri, v = gen()
    S = mc.analyze(n, cp, ri)
    os.environ["B200_MIN_DIST_FRONT"] = str(min_front)
    try:
        sch = [sched_sim.export(mc, S, world, r, with_accesses=True) for r in range(world)]
    finally:
        del os.environ["B200_MIN_DIST_FRONT"]
    sched_sim.check([(a, b) for a, b, c in sch])
    total = sum(sched_sim.check_hazards(ls, ms, acc, r) for r, (ls, ms, acc) in enumerate(sch))
    assert total > 0
    mc.lib().b200_symbolic_free(S)


def test_hazard_check_notices_a_missing_wait(mc):
    """the checker itself: dropping the wait that orders a trailing update after the panel's arrival must be reported"""
    import copy
    n, cp, ri, v = stencils.poisson3d_upper(22)
    S = mc.analyze(n, cp, ri)
    os.environ["B200_MIN_DIST_FRONT"] = "128"
    try:
        sch = [sched_sim.export(mc, S, 2, r, with_accesses=True) for r in range(2)]
    finally:
        del os.environ["B200_MIN_DIST_FRONT"]
    caught = 0; tried = 0
    for r, (ls, ms, acc) in enumerate(sch):
        recv_events = {l["rec"] for l in ls if l["msgs"] and any(ms[q][1] == r for q in l["msgs"]) and l["rec"] >= 0}
        for i, l in enumerate(ls):
            if l["wait"] in recv_events and l["stream"] in (0, 1):
                bad = copy.deepcopy(ls); bad[i]["wait"] = -1; tried += 1
                try:
                    sched_sim.check_hazards(bad, ms, acc, r)
                except AssertionError:
                    caught += 1
    assert tried > 0 and caught > 0, (tried, caught)
    mc.lib().b200_symbolic_free(S)
