#!/bin/bash
# A/B of the SM partition (DY4_LOOP_SMS: SMs set aside for the PLL's serial loops; 0 = none) over stream counts
cd "$(dirname "$0")/.."
for s in ${STREAMS:-256}; do
 for l in ${LOOPS:-0 32 64}; do
  echo "== streams=$s DY4_LOOP_SMS=$l"
  DY4_LOOP_SMS=$l DY4_TRACE=1 python bench.py --steps 5 --warmup 3 --no-cpu --streams $s 2> gpurun_out/part_${s}_$l.err | tee gpurun_out/part_${s}_$l.json | python tools/bench_summary.py | grep -E "value|pll:" | cut -c1-150
 done
done
