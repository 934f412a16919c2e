"""Timeline of the last call in a DY4_TRACE=1 log (development tool): python tools/trace_view.py bench.err"""
import re
import sys

names = {0: 'FE', 1: 'BPF', 2: 'LOOP', 3: 'AUD', 4: 'tail', 5: 'rBPF', 6: 'rPLL', 7: 'rBB', 8: 'aux'}
rows = []
for l in open(sys.argv[1]):
    m = re.match(r'dy4-trace k=(\d+) start=([\d.]+) end=([\d.]+)', l)
    if m:
        rows.append((int(m[1]), float(m[2]), float(m[3])))
starts = [i for i, (k, a, b) in enumerate(rows) if k == 0 and (i == 0 or rows[i - 1][0] == 4)]   # a call starts with a front-end launch that follows the previous call's closing tails
calls = [rows[a:b] for a, b in zip(starts, starts[1:] + [len(rows)])]
last = [c for c in calls if len(c) >= 20][-1]        # the last pipelined call (single-launch calls of the isolated pass are short)
t0 = last[0][1]
for k, a, b in sorted(last, key=lambda r: r[1]):
    if k != 4 or '-t' in sys.argv:
        print("  %-5s %7.3f -> %7.3f  (%.3f)" % (names.get(k, k), a - t0, b - t0, b - a))
