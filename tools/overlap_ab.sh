#!/bin/bash
# A/B of pipelined calls (DY4_FLAG_PIPELINED: steps overlap on the device) and the SM partition that goes with them (DY4_LOOP_SMS),
# on the default bench workload; DY4_TRACE timelines land in gpurun_out/ov_*.err (tools/trace_calls.py)
cd "$(dirname "$0")/.."
run() { echo "== $EXTRA $*"; env "$@" DY4_TRACE=1 python bench.py --steps ${STEPS:-8} --warmup 3 --no-cpu $EXTRA 2> gpurun_out/ov_$TAG.err | tee gpurun_out/ov_$TAG.json | python tools/bench_summary.py | grep -E "value|pll:|frontend" | cut -c1-160; }
TAG=off EXTRA=--no-overlap run A=1
TAG=on_nopart EXTRA= run DY4_LOOP_SMS=0
TAG=on_32 EXTRA= run A=1
TAG=on_64 EXTRA= run DY4_LOOP_SMS=64
