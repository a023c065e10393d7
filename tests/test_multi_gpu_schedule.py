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
