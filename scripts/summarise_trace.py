"""compact per-kind summary of a B200_TRACE_ALL log: top-of-tree launches (depth <= 2) and the rest"""
import re, sys, collections
agg = collections.defaultdict(lambda: [0, 0.0, 0, 0.0])
for l in open(sys.argv[1]):
    m = re.match(r"launch (\S+) depth=(\d+) count=(\d+) stream=(\d) ms=([\d.]+)", l)
    if not m: continue
    k, d, ms = m.group(1), int(m.group(2)), float(m.group(5))
    a = agg[k]
    if d <= 2: a[0] += 1; a[1] += ms
    else: a[2] += 1; a[3] += ms
for k, (n0, m0, n1, m1) in sorted(agg.items(), key=lambda x: -(x[1][1] + x[1][3])):
    print(f"{k:20s} depth<=2: n={n0:5d} {m0:9.3f} ms (avg {m0 / max(n0, 1) * 1e3:7.1f} us) | deeper: n={n1:5d} {m1:9.3f} ms")
