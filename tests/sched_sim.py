"""CPU replay of the static multi-GPU factor schedules of all ranks (built without a device by b200_ctx_create_dry).

Model (what CUDA and NCCL guarantee, nothing more):
  * a rank issues its launches in list order; each stream executes its own launches in that order;
  * a launch with wait_ev >= 0 waits for that event; the event must be recorded by a launch issued EARLIER on the same
    rank (cudaStreamWaitEvent on an event whose record has not been issued yet does not wait at all -> a race);
  * a compute launch finishes as soon as it may start; it then records its event;
  * an NCCL launch (a group of point-to-point messages) becomes ACTIVE when it may start; a message completes when the
    launches holding its two ends are both active; the launch finishes when all its messages have completed;
  * messages between one (src, dst) pair on one communicator match in issue order, so both sides must list the same
    (buffer, offset, count) sequence.
If the replay stops with unfinished launches, the schedule can deadlock."""
import collections
import ctypes as C

COMM_STREAMS = (2, 3)


def export(mc, S, world, rank, with_accesses=False):
    L = mc.lib()
    c = L.b200_ctx_create_dry(rank, world)
    assert c, "dry context"
    rc = L.b200_upload_symbolic(c, S)
    assert rc == 0, L.b200_last_error()
    nl = L.b200_debug_schedule(c, None, 0); nm = L.b200_debug_messages(c, None, 0)
    le = (mc.b200_sched_entry_t * max(nl, 1))(); me = (mc.b200_msg_t * max(nm, 1))()
    L.b200_debug_schedule(c, le, nl); L.b200_debug_messages(c, me, nm)
    launches = [dict(kind=le[i].kind, stream=le[i].stream, wait=le[i].wait_ev, rec=le[i].record_ev, depth=le[i].depth,
                     msgs=list(range(le[i].first_msg, le[i].first_msg + le[i].nmsg)) if le[i].nmsg > 0 and le[i].first_msg >= 0 else [],
                     name=L.b200_debug_kind_name(le[i].kind).decode(), count=le[i].count, flops=le[i].flops, kavg=le[i].kavg, bytes=le[i].bytes) for i in range(nl)]
    msgs = [(me[i].src, me[i].dst, me[i].which, me[i].offset, me[i].count) for i in range(nm)]
    if with_accesses:
        import numpy as np
        na = L.b200_debug_accesses(c, None, 0)
        ae = (mc.b200_access_t * max(na, 1))(); L.b200_debug_accesses(c, ae, na)
        acc = np.frombuffer(ae, dtype=np.dtype([("launch", "<i4"), ("which", "<i4"), ("write", "<i4"), ("pad", "<i4"), ("offset", "<i8"), ("count", "<i8")]), count=na).copy()
        L.b200_ctx_destroy(c)
        return launches, msgs, acc
    L.b200_ctx_destroy(c)
    return launches, msgs


def check(schedules):
    """schedules[r] = (launches, msgs) of rank r.  Returns statistics; raises AssertionError on any violation."""
    world = len(schedules)
    # ---- static checks
    for r, (ls, ms) in enumerate(schedules):
        rec_at = {}
        for i, l in enumerate(ls):
            if l["wait"] >= 0:
                assert l["wait"] in rec_at, f"rank {r}: launch {i} ({l['name']}) waits for event {l['wait']} whose record has not been issued yet"
            if l["rec"] >= 0:
                assert l["rec"] not in rec_at, f"rank {r}: event {l['rec']} recorded twice"
                rec_at[l["rec"]] = i
            for q in l["msgs"]:
                s, d = ms[q][0], ms[q][1]
                assert r in (s, d) and s != d, f"rank {r}: message {ms[q]} does not involve this rank"
                assert l["stream"] in COMM_STREAMS
    # per (src, dst, comm) sequences must agree on both sides
    seq = collections.defaultdict(lambda: [[], []])      # key -> [sender's list, receiver's list] of (launch idx, msg idx, payload)
    for r, (ls, ms) in enumerate(schedules):
        for i, l in enumerate(ls):
            for q in l["msgs"]:
                s, d, which, off, cnt = ms[q]
                seq[(s, d, l["stream"])][0 if r == s else 1].append((i, q, (which, off, cnt)))
    nbytes = 0
    for key, (snd, rcv) in seq.items():
        assert [x[2] for x in snd] == [x[2] for x in rcv], f"pair {key}: send and receive sequences differ"
        nbytes += sum(x[2][2] for x in snd) * 8
    # ---- dynamic replay
    peer_of = {}                                          # (rank, launch, msg) -> (peer rank, peer launch)
    for (s, d, st), (snd, rcv) in seq.items():
        for a, b in zip(snd, rcv):
            peer_of[(s, a[0], a[1])] = (d, b[0]); peer_of[(d, b[0], b[1])] = (s, a[0])
    by_stream = [collections.defaultdict(collections.deque) for _ in range(world)]
    for r, (ls, _) in enumerate(schedules):
        for i, l in enumerate(ls):
            by_stream[r][l["stream"]].append(i)
    recorded = [set() for _ in range(world)]
    active = [set() for _ in range(world)]
    done = [set() for _ in range(world)]
    pending = [{i: set(l["msgs"]) for i, l in enumerate(ls) if l["msgs"]} for ls, _ in schedules]
    progress = True; rounds = 0
    while progress:
        progress = False; rounds += 1
        for r in range(world):
            ls = schedules[r][0]
            for st, dq in by_stream[r].items():
                while dq:
                    i = dq[0]; l = ls[i]
                    if i not in active[r]:
                        if l["wait"] >= 0 and l["wait"] not in recorded[r]:
                            break
                        active[r].add(i); progress = True
                    if l["msgs"]:
                        for q in list(pending[r][i]):
                            pr, pi = peer_of[(r, i, q)]
                            if pi in active[pr]:
                                pending[r][i].discard(q); progress = True
                        if pending[r][i]:
                            break
                    done[r].add(i); dq.popleft(); progress = True
                    if l["rec"] >= 0:
                        recorded[r].add(l["rec"])
    stuck = {r: [(i, schedules[r][0][i]["name"], schedules[r][0][i]["stream"], schedules[r][0][i]["depth"]) for st, dq in by_stream[r].items() for i in list(dq)[:1]] for r in range(world) if any(by_stream[r].values())}
    assert not stuck, f"schedule deadlocks; heads of the blocked streams: {stuck}"
    return dict(launches=[len(s[0]) for s in schedules], messages=[len(s[1]) for s in schedules], nccl_bytes=nbytes, rounds=rounds)


def check_hazards(launches, msgs, acc, rank):
    """Every message must be ORDERED (stream order / events) against every compute launch of the same rank that touches the
    same memory: a received range against its readers and writers, a sent range against its writers.  Accesses are the
    conservative intervals of b200_debug_accesses.  Returns the number of (message, access) pairs examined."""
    import numpy as np
    m = len(launches)
    anc = [0] * m                      # bit j of anc[i]: launch j happens before launch i
    last_in_stream = {}; rec_at = {}
    for i, l in enumerate(launches):
        a = 0
        p = last_in_stream.get(l["stream"])
        if p is not None:
            a |= anc[p] | (1 << p)
        if l["wait"] >= 0:
            q = rec_at[l["wait"]]; a |= anc[q] | (1 << q)
        anc[i] = a
        last_in_stream[l["stream"]] = i
        if l["rec"] >= 0:
            rec_at[l["rec"]] = i
    by_which = {w: acc[acc["which"] == w] for w in range(4)}
    pairs = 0
    for ci, l in enumerate(launches):
        for q in l["msgs"]:
            src, dst, which, off, cnt = msgs[q]
            a = by_which[which]
            if len(a) == 0:
                continue
            hit = a[(a["offset"] < off + cnt) & (a["offset"] + a["count"] > off)]
            if src == rank:
                hit = hit[hit["write"] == 1]           # a send only conflicts with writers
            for li in np.unique(hit["launch"]):
                li = int(li); pairs += 1
                ordered = (anc[li] >> ci) & 1 or (anc[ci] >> li) & 1
                assert ordered, (f"rank {rank}: {'send' if src == rank else 'receive'} of buffer {which} [{off}, {off + cnt}) in launch {ci} ({l['name']}, depth {l['depth']}) "
                                 f"is not ordered against launch {li} ({launches[li]['name']}, stream {launches[li]['stream']}, depth {launches[li]['depth']}) that touches the same range")
                # data that arrives must arrive BEFORE it is used.  L and the inverse blocks are written once per factorization;
                # the contribution-block arenas are reused by every second tree level, so there the rule binds the readers of the
                # SAME level only (an earlier level's reader must merely be ordered: it read the slot's previous content)
                if dst == rank and (which < 2 or launches[li]["depth"] == l["depth"]):
                    rd = hit[(hit["launch"] == li) & (hit["write"] == 0)]
                    if len(rd):
                        assert (anc[li] >> ci) & 1, f"rank {rank}: launch {li} ({launches[li]['name']}) reads [{off}, {off + cnt}) of buffer {which} before the message in launch {ci} delivers it"
    return pairs
