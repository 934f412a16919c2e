for cfg in "--mode 2 --streams 1024 --blocks 12" "--mode 3 --streams 1024 --blocks 12" "--mode 1 --streams 256 --blocks 36"; do
python bench.py $cfg --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$cfg', d['value'], d['e2e']['value'], d['ms_per_step'], {k:(round(v['avg_ms'],3),v['launches']) for k,v in d['kernels'].items()}, {k:v['avg_ms'] for k,v in d['roofline']['whole_job_launches'].items()})"
done
