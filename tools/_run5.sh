timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu6.log 2>&1; tail -3 gpurun_out/pytest_gpu6.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_r01g.json 2> gpurun_out/bench_r01g.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r01g.json'))
print(d['value'], d['e2e']['value'])
print(d['regex'])
PY
