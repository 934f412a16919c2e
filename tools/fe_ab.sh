# A/B of front-end kernel variants: parity subset + bench (whole-job front-end launch time in roofline.whole_job_launches)
for v in "$@"; do
  DY4_FRONTEND=$v timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_stereo or batch_against or ragged" > gpurun_out/fe_${v}_tests.log 2>&1; echo "$v tests: $(tail -1 gpurun_out/fe_${v}_tests.log)"
  DY4_FRONTEND=$v timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/fe_${v}_bench.json 2> gpurun_out/fe_${v}_bench.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/fe_${v}_bench.json").read().strip().splitlines()[-1])
    print("${v}", d["value"], d["e2e"]["value"], {k:(x["avg_ms"],x["frac_of_fma_peak"]) for k,x in d["roofline"]["whole_job_launches"].items()}, d["kernels"]["frontend"]["avg_ms"])
except Exception as e:
    print("${v} failed", e)
PY
done
