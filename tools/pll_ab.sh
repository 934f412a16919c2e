#!/bin/bash
# A/B of the serial PLL loop variants on one GPU (development aid): bench.py per setting, key numbers only.
# usage: tools/pll_ab.sh "ENV=VAL ENV2=VAL2" "..." ...
for cfg in "$@"; do
env $cfg DY4_PLL_STATS=1 python bench.py --steps 3 --warmup 3 --no-cpu 2>/tmp/pll_ab.err | env CFG="$cfg" python -c "
import json,sys,os
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']
print('%-40s value %8.0f ms/step %6.3f  pll ms/step %6.3f ns/sample %6.2f | aux %s | fe %s bpf %s audio %s | pcm_ok %s' % (
  os.environ['CFG'], d['value'], d['ms_per_step'], d['pll']['ms_per_step'], d['pll']['ns_per_sample_per_stream'],
  k['pll_aux']['avg_ms'], k['frontend']['avg_ms'], k['twin_bpf']['avg_ms'], k['audio']['avg_ms'], d['e2e']['pcm_equals_device_path']))"
grep 'dy4 pll stats' /tmp/pll_ab.err | tail -1
done
