for cfg in "--streams 1024 --blocks 24" "--streams 4096 --blocks 6" "--streams 16384 --blocks 2" "--streams 65536 --blocks 1"; do
python bench.py $cfg --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=d['kernels']; fir=sum(v['share'] for n,v in k.items() if n in ('frontend','twin_bpf','audio','tails')); print('$cfg', d['value'], d['e2e']['value'], d['ms_per_step'], 'pll share', k['pll']['share'], 'fir share', round(fir,3), {n:round(v['avg_ms'],3) for n,v in k.items()})"
done
