#!/bin/bash
# One bench line per BASELINE.json config on one GPU, kept under profiles/ (evidence, not the headline): usage tools/configs_bench.sh r2
R=${1:-r2}
run() { name=$1; shift; python bench.py "$@" $EXTRA --steps ${STEPS:-3} --warmup 3 --no-cpu > gpurun_out/${R}_bench_$name.json 2> gpurun_out/${R}_bench_$name.err || echo "FAILED $name"; python - <<P
import json
d=json.loads(open("gpurun_out/${R}_bench_$name.json").read().strip().splitlines()[-1])
print("%-8s value %9.0f M/s  e2e %8.0f  ms/step %7.3f  clocks %s  %s" % ("$name", d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], {k: v["avg_ms"] for k, v in d["kernels"].items()}))
P
}
run mono --mono
run m1 --mode 1 --streams 256 --blocks 36
run m2 --mode 2 --streams 1024 --blocks 12
run m3 --mode 3 --streams 1024 --blocks 12
run rds --rds --streams 4096 --blocks 6
run s1024 --streams 1024 --blocks 24
run s4096 --streams 4096 --blocks 6
run s16384 --streams 16384 --blocks 3
run s65536 --streams 65536 --blocks 1
