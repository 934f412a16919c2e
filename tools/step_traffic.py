"""Per-kernel launches, time and DRAM bytes of the TIMED steps in an ncu launch list of `bench.py --steps K --warmup W`
(tools/prof_r2.sh writes the list):  python tools/step_traffic.py gpurun_out/r2_step_launches.csv [W=3] [K=2] [pairs_per_step]
Calls are told apart by their first front-end launch (the one not preceded by a band-pass / tails launch)."""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
W = int(sys.argv[2]) if len(sys.argv) > 2 else 3
K = int(sys.argv[3]) if len(sys.argv) > 3 else 2
pairs = float(sys.argv[4]) if len(sys.argv) > 4 else 256 * 48 * 51200.0
lines = [l for l in open(path) if l.startswith('"')]
launch = OrderedDict()
for r in csv.DictReader(lines):
    d = launch.setdefault(int(r["ID"]), {"name": re.search(r"(k_\w+)", r["Kernel Name"]).group(1)})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(unit, 1)
    d[r["Metric Name"]] = v * scale
seq = list(launch.values())
call, calls = -1, []
for i, d in enumerate(seq):
    if d["name"] == "k_frontend_stream" and (i == 0 or seq[i - 1]["name"] not in ("k_bpf_mixed", "k_twin_bpf")):
        call += 1
    calls.append(call)
rows = OrderedDict()
for d, c in zip(seq, calls):
    if W <= c < W + K:
        r = rows.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
        r[0] += 1
        r[1] += d.get("gpu__time_duration.sum", 0.0)
        r[2] += d.get("dram__bytes_read.sum", 0.0)
        r[3] += d.get("dram__bytes_write.sum", 0.0)
print("calls in the list: %d; rows below: calls %d..%d (the timed steps)" % (call + 1, W, W + K - 1))
print("| kernel | launches/step | ms/step (serialised, cold) | DRAM read B/pair | DRAM write B/pair | total B/pair |")
print("|---|---|---|---|---|---|")
tot_ms = tot_b = 0.0
for k, (n, ms, rd, wr) in sorted(rows.items(), key=lambda kv: -(kv[1][2] + kv[1][3])):
    print("| %s | %g | %.3f | %.2f | %.2f | %.2f |" % (k, n / K, ms / K, rd / K / pairs, wr / K / pairs, (rd + wr) / K / pairs))
    tot_ms += ms / K
    tot_b += (rd + wr) / K / pairs
print("| **sum** | | %.2f | | | **%.2f** |" % (tot_ms, tot_b))
