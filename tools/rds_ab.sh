for t in 32 128; do for cfg in "--rds --streams 4096 --blocks 6" "--streams 4096 --blocks 6" "--streams 256 --blocks 47"; do
DY4_PLL_THREADS=$t python bench.py $cfg --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('threads=$t', '$cfg', d['value'], d['e2e']['value'], d['ms_per_step'], {k:round(v['avg_ms'],3) for k,v in d['kernels'].items()})"
done; done
