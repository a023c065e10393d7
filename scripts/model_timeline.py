"""Timeline model of the (multi-GPU) factorization from the device-free schedules: every launch gets a cost from a few
measured constants, the streams / events / messages of all ranks are replayed in time, the makespan is the predicted factor
time.  A planning aid only -- calibrated against the measured 1/2/4/8-GPU times of the mid-round ordering
(B200_ND_STARTS=1 B200_ND_UB=1.20: 678 / 408 / 254 / 212 ms); nothing here is a measurement.
usage: python scripts/model_timeline.py [grid] [world ...]"""
import sys, os, time, heapq
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
m = g.load_package()
import sched_sim
from meagre_crowd_b200 import stencils

SLOTS = 148 * 3
P = dict(launch_us=7.0, gemm_c=90.0, gemm_floor_us=6.0, kstep_us=0.33, diag_us=42.0, diag_small_us=14.0, small_front_us=15.0, trsm_rows_us=22.0,
         asm_tbs=2.7, memset_tbs=6.8, msg_us=50.0, p2p_gbs=225.0, event_us=12.0)     # fitted by hand to 678 / 408 / 254 / 212 ms
for kv in os.environ.get("MODEL_P", "").split(","):
    if "=" in kv: P[kv.split("=")[0]] = float(kv.split("=")[1])


def cost_us(l, ms, rank):
    k = l["name"]
    if k in ("gemm_dmma", "trsm_dmma"):
        K = max(l["kavg"], 16.0); rate = 37e12 * K / (K + P["gemm_c"])
        waves = max(1.0, l["count"] / SLOTS)
        return max(P["gemm_floor_us"] + K / 16 * P["kstep_us"] * min(waves, 1.0) * 3, l["flops"] / rate * 1e6) + P["launch_us"]
    if k == "diag_block": return P["diag_us"] * max(1.0, l["count"] / SLOTS) + P["launch_us"]
    if k == "diag_block_small": return P["diag_small_us"] * max(1.0, l["count"] / SLOTS) + P["launch_us"]
    if k == "small_front": return P["small_front_us"] * max(1.0, l["count"] / (148 * 4)) + P["launch_us"]
    if k == "trsm_rows": return P["trsm_rows_us"] * max(1.0, l["count"] / (148 * 2)) + P["launch_us"]
    if k == "extend_add": return l["bytes"] / (P["asm_tbs"] * 1e12) * 1e6 + 8.0
    if k == "memset_cb": return l["bytes"] / (P["memset_tbs"] * 1e12) * 1e6 + 3.0
    if k.startswith("nccl"):
        snd = sum(ms[q][4] for q in l["msgs"] if ms[q][0] == rank) * 8; rcv = sum(ms[q][4] for q in l["msgs"] if ms[q][1] == rank) * 8
        return P["msg_us"] + max(snd, rcv) / (P["p2p_gbs"] * 1e9) * 1e6
    return 1.0   # nop


def simulate(sch):
    """event-driven replay: a launch starts when its stream is free, its event has fired and (NCCL) all peer launches holding the
    other ends of its messages have started; it ends cost later.  Returns per-rank finish times (ms) and the busy time per kind."""
    world = len(sch)
    seq = {}
    for r, (ls, ms) in enumerate(sch):
        for i, l in enumerate(ls):
            for q in l["msgs"]:
                s, d = ms[q][0], ms[q][1]
                seq.setdefault((s, d, l["stream"]), [[], []])[0 if r == s else 1].append((i, q))
    peers = {}
    for (s, d, st), (snd, rcv) in seq.items():
        for a, b in zip(snd, rcv):
            peers.setdefault((s, a[0]), set()).add((d, b[0])); peers.setdefault((d, b[0]), set()).add((s, a[0]))
    start = [dict() for _ in range(world)]; end = [dict() for _ in range(world)]; evt = [dict() for _ in range(world)]
    streams = [dict() for _ in range(world)]
    for r, (ls, _) in enumerate(sch):
        for i, l in enumerate(ls): streams[r].setdefault(l["stream"], []).append(i)
    pos = [{st: 0 for st in streams[r]} for r in range(world)]
    sfree = [{st: 0.0 for st in streams[r]} for r in range(world)]
    ready = [dict() for _ in range(world)]       # comm launches whose local conditions hold: time they could start
    busy = {}
    progress = True
    while progress:
        progress = False
        for r in range(world):
            ls, ms = sch[r]
            for st, lst in streams[r].items():
                while pos[r][st] < len(lst):
                    i = lst[pos[r][st]]; l = ls[i]
                    if l["wait"] >= 0 and l["wait"] not in evt[r]: break
                    t0 = max(sfree[r][st], evt[r][l["wait"]] + P["event_us"] if l["wait"] >= 0 else 0.0)
                    if l["msgs"]:
                        ready[r][i] = t0
                        ps = peers.get((r, i), ())
                        if not all(pi in ready[pr] for pr, pi in ps): break
                        t0 = max([t0] + [ready[pr][pi] for pr, pi in ps])
                    c = cost_us(l, ms, r)
                    start[r][i] = t0; end[r][i] = t0 + c; sfree[r][st] = t0 + c
                    busy[l["name"]] = busy.get(l["name"], 0.0) + c / world
                    if l["rec"] >= 0: evt[r][l["rec"]] = t0 + c
                    pos[r][st] += 1; progress = True
    assert all(pos[r][st] == len(streams[r][st]) for r in range(world) for st in streams[r]), "model replay stuck"
    return [max(end[r].values()) / 1e3 for r in range(world)], busy


if __name__ == "__main__":
    grid = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    worlds = [int(a) for a in sys.argv[2:]] or [1, 2, 4, 8]
    n, cp, ri, v = stencils.poisson3d_upper(grid)
    S = m.analyze(n, cp, ri); s = S.contents
    print(f"grid {grid}: F={s.flops:.4e} nnzL={s.nnzL:.4e} maxfront={s.maxfront}", flush=True)
    for w in worlds:
        sch = [sched_sim.export(m, S, w, r) for r in range(w)]
        fin, busy = simulate(sch)
        top = sorted(busy.items(), key=lambda x: -x[1])[:6]
        print(f"world {w}: predicted factor {max(fin):7.1f} ms  ({s.flops / max(fin) / 1e9:.1f} TFLOP/s)  per rank {[round(x) for x in fin]}  busy/rank: " + ", ".join(f"{k} {v_ / 1e3:.0f}" for k, v_ in top), flush=True)
