"""CPU check of the persistent-sweep schedules (csrc/cuda/b200_sweep.cuh): the unit lists of a device-free context are replayed with
exactly what the kernel guarantees -- units are handed out in list order, a unit starts when its counters have reached the
values it waits for, signals happen when it ends -- and three properties are proved for a given problem:

  1. no deadlock: every counter a unit waits for is raised only by units that precede it in the list (so whichever CTAs
     are resident, the oldest unfinished unit can always run);
  2. every wait means what the builder intends: "cnt[c] >= v" is only ever satisfied after the FIRST v signallers of c (in
     list order) have all finished -- either they are all the signallers, or every later signaller happens-after them;
  3. no data race: for every row of y / z / w, every access that conflicts with an earlier one (in list order) happens-after
     it through the wait/signal relation (transitively).

Happens-before sets are Python integers used as bit sets, so this is for test-size problems (a few 10^4 units).
Test infrastructure only."""
import ctypes as C
import numpy as np

UK_FWD_GATHER, UK_FWD_SMALL, UK_GEMV, UK_GEMVT, UK_BWD_SMALL = range(5)
FIELDS = ("kind", "task", "o0", "aux", "sn", "w1", "n1", "v1", "w2", "v2", "sig", "up")


def export(mc, S):
    """unit lists (forward, backward) and task table of the persistent sweeps of a symbolic analysis, from a schedule-only context"""
    L = mc.lib()
    c = L.b200_ctx_create_dry(0, 1)
    assert c
    assert L.b200_upload_symbolic(c, S) == 0, L.b200_last_error()
    out = []
    for d in (0, 1):
        n = L.b200_debug_sweep_units(c, d, None, 0)
        a = np.zeros((max(n, 1), 16), dtype=np.int32)
        L.b200_debug_sweep_units(c, d, a.ctypes.data_as(C.c_void_p), n)
        out.append(a[:n])
    nt = L.b200_debug_sweep_tasks(c, None, 0)
    t = np.zeros((max(nt, 1), 8), dtype=np.int64)
    L.b200_debug_sweep_tasks(c, t.ctypes.data_as(C.c_void_p), nt)
    ncnt = L.b200_debug_sweep_counters(c)
    L.b200_ctx_destroy(c)
    return out[0], out[1], t[:nt], ncnt


class Sym:
    def __init__(self, mc, S):
        s = S.contents
        self.n, self.ns = s.n, s.ns
        self.start = mc.sym_array(S, "sn_start", s.ns + 1, np.int32).astype(np.int64)
        self.parent = mc.sym_array(S, "sn_parent", s.ns, np.int32)
        self.rowptr = mc.sym_array(S, "rowptr", s.ns + 1, np.int64)
        self.rowidx = mc.sym_array(S, "rowidx", int(self.rowptr[-1]), np.int32)
        self.relidx = mc.sym_array(S, "relidx", int(self.rowptr[-1]), np.int32)
        self.k = np.diff(self.start); self.f = np.diff(self.rowptr); self.u = self.f - self.k
        self.wgl = np.concatenate([[0], np.cumsum(self.u)])[:-1]
        self.children = [[] for _ in range(s.ns)]
        for c in range(s.ns):
            if self.parent[c] >= 0: self.children[self.parent[c]].append(c)


BASE = 1 << 40


def accesses(u, tasks, sy):
    """(reads, writes): arrays of location ids (base * 2^40 + row; bases 0 y, 1 z, 2 w) of one unit"""
    kind, task, o0, aux, sn = (int(u[i]) for i in range(5))
    k, f, c0 = int(sy.k[sn]), int(sy.f[sn]), int(sy.start[sn])
    if kind == UK_GEMV:
        vsel, voff, K, dsel, doff, N, tri, mode = (int(x) for x in tasks[task])
        rows = np.arange(o0, min(o0 + 32, N))
        k_lo = (o0 & ~63) if tri == 2 else 0
        k_hi = min(K, ((o0 + 31) | 63) + 1) if tri == 1 else K
        rd = [vsel * BASE + voff + np.arange(k_lo, k_hi)]
        wr = dsel * BASE + doff + rows
        if mode == 0: rd.append(wr)
        return np.concatenate(rd), wr
    if kind == UK_GEMVT:
        vsel, voff, K, dsel, doff, N, tri, mode = (int(x) for x in tasks[task])
        cols = np.arange(o0, min(o0 + 64, N))
        if aux:
            ri = sy.rowidx[sy.rowptr[sn] + k + voff: sy.rowptr[sn] + k + voff + K].astype(np.int64)
            assert len(ri) == K
            rd = vsel * BASE + ri
        else:
            rd = vsel * BASE + voff + np.arange(K)
        wr = dsel * BASE + doff + cols
        return np.concatenate([rd, wr]), wr
    if kind in (UK_FWD_GATHER, UK_FWD_SMALL):
        lo, hi = (o0, aux) if kind == UK_FWD_GATHER else (0, f)
        rd, wr = [], []
        for c in sy.children[sn]:
            kc, uc = int(sy.k[c]), int(sy.u[c])
            rel = sy.relidx[sy.rowptr[c] + kc: sy.rowptr[c] + kc + uc].astype(np.int64)
            sel = np.nonzero((rel >= lo) & (rel < hi))[0]
            rd.append(2 * BASE + sy.wgl[c] + sel)
            d = rel[sel]
            dst = np.where(d < k, 0 * BASE + c0 + d, 2 * BASE + sy.wgl[sn] + (d - k))
            wr.append(dst)
        if kind == UK_FWD_SMALL:
            rd.append(0 * BASE + c0 + np.arange(k))                    # y of the pivots
            wr.append(1 * BASE + c0 + np.arange(k))                    # z
            wr.append(2 * BASE + sy.wgl[sn] + np.arange(f - k))        # w -= L21 z
            # the gathered pivot values stay in shared memory: y itself is not written by a small unit
            wr = [w for w in wr]
            w_all = np.concatenate(wr) if wr else np.zeros(0, np.int64)
            w_all = w_all[(w_all // BASE) != 0]
            return np.concatenate(rd + [w_all]), w_all
        w_all = np.concatenate(wr) if wr else np.zeros(0, np.int64)
        return np.concatenate(rd + [w_all]) if rd else w_all, w_all
    if kind == UK_BWD_SMALL:
        ri = sy.rowidx[sy.rowptr[sn] + k: sy.rowptr[sn] + f].astype(np.int64)
        rd = np.concatenate([1 * BASE + c0 + np.arange(k), 0 * BASE + ri])
        return rd, 0 * BASE + c0 + np.arange(k)
    raise AssertionError("unknown unit kind")


def check(units, tasks, sy, ncnt, direction):
    """the three properties for one sweep; raises AssertionError with the offending unit; returns some statistics"""
    N = len(units)
    U = [dict(zip(FIELDS, (int(x) for x in units[i][:12]))) for i in range(N)]
    # signallers of every counter, in list order: a unit adds 1 to cnt[sig] (a tile / product counter) and 1 to cnt[up] (a completion counter)
    sig = {}
    for i, u in enumerate(U):
        if u["sig"] >= 0: sig.setdefault(u["sig"], []).append((i, 1))
        if u["up"] >= 0: sig.setdefault(u["up"], []).append((i, 1))
    for c in sig:
        assert 0 <= c < ncnt
    # pass 1: happens-before sets under the intended meaning of every wait: cnt[c] >= v stands for the shortest prefix (in list order) of c's
    # signallers whose contributions add up to v -- and they must add up to v EXACTLY, so that no other subset can reach it without them
    HB = {}
    waits = [[] for _ in range(N)]          # per unit: list of (counter, v)
    prefix_cache = {}

    def preds_of(i, c, v):
        if (c, v) not in prefix_cache:
            lst = sig.get(c, [])
            tot = 0; m = 0
            while m < len(lst) and tot < v:
                tot += lst[m][1]; m += 1
            assert tot == v, f"unit {i} waits for cnt[{c}] >= {v}: the signals of its first {m} signallers add up to {tot}, not {v}"
            prefix_cache[(c, v)] = m
        return [d for d, w in sig[c][:prefix_cache[(c, v)]]]
    for i, u in enumerate(U):
        hb = 0
        ws = [(u["w1"] + j, u["v1"]) for j in range(u["n1"])] + ([(u["w2"], u["v2"])] if u["w2"] >= 0 else [])
        assert all(w[1] >= 0 for w in ws), f"unit {i}: a wait value was never filled in"
        waits[i] = [w for w in ws if w[1] > 0]
        for c, v in waits[i]:
            for d in preds_of(i, c, v):
                assert d < i, f"unit {i} waits for unit {d}, which follows it in the list (deadlock)"
                hb |= HB[d] | (1 << d)
        HB[i] = hb
    # pass 2: a wait for a proper prefix of the signallers is sound only if every later signaller happens-after the whole prefix
    seen = set()
    for i in range(N):
        for c, v in waits[i]:
            if (c, v) in seen: continue
            seen.add((c, v))
            lst = sig[c]; m = prefix_cache[(c, v)]
            if m == len(lst): continue
            first = 0
            for d, w in lst[:m]: first |= 1 << d
            for d, w in lst[m:]:
                assert d in HB and (HB[d] & first) == first, f"cnt[{c}] >= {v} (unit {i}) can be satisfied by signaller {d} before the first {m} signallers are done"
    # pass 3: data races
    lastw = {}; readers = {}
    nconf = 0
    for i, u in enumerate(units):
        rd, wr = accesses(u, tasks, sy)
        hb = HB[i]
        wset = set(int(x) for x in wr)
        for loc in set(int(x) for x in rd) - wset:
            w = lastw.get(loc, -1)
            if w >= 0:
                nconf += 1
                assert (hb >> w) & 1, f"{direction}: unit {i} {U[i]} reads location {loc // BASE}:{loc % BASE} written by unit {w} {U[w]} without happening after it"
            readers.setdefault(loc, []).append(i)
        for loc in wset:
            w = lastw.get(loc, -1)
            if w >= 0:
                nconf += 1
                assert (hb >> w) & 1, f"{direction}: unit {i} {U[i]} overwrites location {loc // BASE}:{loc % BASE} last written by unit {w} {U[w]} without happening after it"
            for r in readers.get(loc, ()):
                nconf += 1
                assert (hb >> r) & 1, f"{direction}: unit {i} {U[i]} overwrites location {loc // BASE}:{loc % BASE} that unit {r} {U[r]} reads, without happening after it"
            lastw[loc] = i; readers[loc] = []
    return dict(units=N, counters=len(sig), ordered_conflicts=nconf)


def coverage(units_f, units_b, tasks, sy):
    """every pivot gets its z in the forward sweep and its x in the backward sweep exactly once"""
    for units, base in ((units_f, 1), (units_b, 0)):
        cnt = np.zeros(sy.n, dtype=np.int64)
        for u in units:
            kind, task, o0, aux, sn = (int(u[i]) for i in range(5))
            if kind in (UK_FWD_SMALL, UK_BWD_SMALL):
                cnt[int(sy.start[sn]): int(sy.start[sn + 1])] += 1
            elif kind == UK_GEMV:
                vsel, voff, K, dsel, doff, N, tri, mode = (int(x) for x in tasks[task])
                if mode == 1 and dsel == base: cnt[doff + o0: doff + min(o0 + 32, N)] += 1
        assert (cnt == 1).all(), f"base {base}: {int((cnt != 1).sum())} pivots are not produced exactly once"
