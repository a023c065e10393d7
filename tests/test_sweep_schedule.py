"""The persistent-sweep schedules (p <= 4 right-hand sides; csrc/cuda/b200_sweep.cuh) checked without a GPU: tests/sweep_sim.py proves on
the exported unit lists that no unit waits for a later one (no deadlock), that every counter wait can only be satisfied by the
signals it is meant for, and that every conflicting pair of accesses to y / z / w is ordered by the wait/signal relation (no race).
A removed wait must be noticed (the checker is itself checked)."""
import os
import random

import numpy as np
import pytest

import sweep_sim
from meagre_crowd_b200 import stencils

CASES = [
    ("poisson3d_12", lambda: stencils.poisson3d_upper(12), 0),
    ("poisson3d_24", lambda: stencils.poisson3d_upper(24), 0),
    ("poisson3d_20x13x9", lambda: stencils.poisson3d_upper(20, 13, 9), 0),
    ("poisson2d_300", lambda: stencils.poisson2d_upper(300), 0),
    ("poisson2d_130x7", lambda: stencils.poisson2d_upper(130, 7), 0),
    ("convdiff27_14_lu", lambda: stencils.convdiff27_full(14), 1),
    ("single_unknown", lambda: stencils.poisson2d_upper(1, 1), 0),
    ("poisson3d_34", lambda: stencils.poisson3d_upper(34), 0),
]


def _analyze(m, gen, kind):
    n, cp, ri, v = gen()
    return m.analyze(n, cp, ri, kind) if kind else m.analyze(n, cp, ri)


@pytest.mark.parametrize("name,gen,kind", CASES, ids=[c[0] for c in CASES])
def test_sweep_units_are_deadlock_free_and_race_free(mc, name, gen, kind):
    m = mc
    S = _analyze(m, gen, kind)
    uf, ub, tk, ncnt = sweep_sim.export(m, S)
    sy = sweep_sim.Sym(m, S)
    sweep_sim.coverage(uf, ub, tk, sy)
    rf = sweep_sim.check(uf, tk, sy, ncnt, "forward")
    rb = sweep_sim.check(ub, tk, sy, ncnt, "backward")
    assert rf["units"] == len(uf) and rb["units"] == len(ub)
    m.lib().b200_symbolic_free(S)


def test_wide_solve_blocks_schedule(mc, monkeypatch):
    m = mc
    """the 1024-column solve blocks of the big fronts, forced onto a small problem"""
    monkeypatch.setenv("B200_SOLVE_WIDE_MIN", "512")
    S = _analyze(m, lambda: stencils.poisson3d_upper(30), 0)
    uf, ub, tk, ncnt = sweep_sim.export(m, S)
    sy = sweep_sim.Sym(m, S)
    assert sy.k.max() >= 512
    sweep_sim.coverage(uf, ub, tk, sy)
    sweep_sim.check(uf, tk, sy, ncnt, "forward")
    sweep_sim.check(ub, tk, sy, ncnt, "backward")
    m.lib().b200_symbolic_free(S)


def test_checker_notices_a_removed_wait(mc):
    m = mc
    S = _analyze(m, lambda: stencils.poisson3d_upper(24), 0)
    uf, ub, tk, ncnt = sweep_sim.export(m, S)
    sy = sweep_sim.Sym(m, S)
    rng = random.Random(7)
    for units, d in ((uf, "forward"), (ub, "backward")):
        cand = [i for i in range(len(units)) if (units[i][6] > 0 and units[i][7] > 0) or units[i][8] >= 0]
        for i in rng.sample(cand, 12):
            u2 = units.copy()
            if u2[i][8] >= 0 and (u2[i][6] == 0 or rng.random() < 0.5):
                u2[i][8] = -1
            else:
                u2[i][6] = 0
            with pytest.raises(AssertionError):
                sweep_sim.check(u2, tk, sy, ncnt, d)
    m.lib().b200_symbolic_free(S)
