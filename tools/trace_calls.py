"""Timeline of several consecutive calls in a DY4_TRACE=1 log (development tool): python tools/trace_calls.py bench.err [first_call [n_calls]]"""
import re
import sys

names = {0: 'FE', 1: 'BPF', 2: 'LOOP', 3: 'AUD', 4: 'tail', 5: 'rBPF', 6: 'rPLL', 7: 'rBB', 8: 'aux'}
rows = []
for l in open(sys.argv[1]):
    m = re.match(r'dy4-trace k=(\d+) start=([\d.]+) end=([\d.]+)', l)
    if m:
        rows.append((int(m[1]), float(m[2]), float(m[3])))
call, lab = -1, []
for i, (k, a, b) in enumerate(rows):                 # records are in host queue order: a call opens with a front-end launch not preceded by a BPF
    if k == 0 and (i == 0 or rows[i - 1][0] != 1):
        call += 1
    lab.append(call)
first = int(sys.argv[2]) if len(sys.argv) > 2 else max(0, call - 4)
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2
sel = [(a, b, k, c) for (k, a, b), c in zip(rows, lab) if first <= c < first + n and k != 4]
t0 = min(x[0] for x in sel)
for a, b, k, c in sorted(sel):
    print("  call %d %-5s %7.3f -> %7.3f  (%.3f)" % (c, names.get(k, k), a - t0, b - t0, b - a))
