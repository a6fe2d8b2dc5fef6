"""Read a SRGP_TIMELINE file (capi.cu: 'class start_ms end_ms stream' per accounted kernel bracket, '# reset' between
evaluations) and print the last evaluation per stream: runs of consecutive brackets of one class with the gaps between them.
   SRGP_TIMELINE=/tmp/tl.txt python tools/run_vi.py 125000 1024 8 6 ; python tools/timeline.py /tmp/tl.txt"""
import sys

NAMES = {0: "assemble", 1: "gen", 2: "gram", 3: "km", 4: "dense", 5: "reduce", 6: "comm", 7: "other"}
STREAMS = {1: "main", 2: "side (stream2)", 3: "generators (stream3)", 4: "low priority (stream4)"}
evals, cur = [], []
for line in open(sys.argv[1]):
    if line.startswith("#"):
        if cur:
            evals.append(cur)
        cur = []
        continue
    f = line.split()
    cur.append((float(f[1]), float(f[2]), int(f[0]), int(f[3]) if len(f) > 3 else 0))
if cur:
    evals.append(cur)
ev = sorted(evals[-1])
print("evaluation: %d brackets, last end %.3f ms" % (len(ev), max(b for _, b, _, _ in ev)))
for st in sorted({t for _, _, _, t in ev}):
    print("-- %s" % STREAMS.get(st, str(st)))
    runs = []
    for a, b, c, t in ev:
        if t != st:
            continue
        if runs and runs[-1][2] == c and a - runs[-1][1] < 0.03:
            runs[-1][1] = max(runs[-1][1], b)
            runs[-1][3] += 1
            runs[-1][4] += b - a
        else:
            runs.append([a, b, c, 1, b - a])
    prev = 0.0
    for a, b, c, k, busy in runs:
        print("%8.3f .. %8.3f  %-8s x%-3d  span %.3f  busy %.3f   gap before %.3f" % (a, b, NAMES.get(c, str(c)), k, b - a, busy, a - prev))
        prev = max(prev, b)
