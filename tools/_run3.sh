timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu4.log 2>&1; tail -5 gpurun_out/pytest_gpu4.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_r01e.json 2> gpurun_out/bench_r01e.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r01e.json'))
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['frac_of_r_rand'])
print(d['regex'])
PY
