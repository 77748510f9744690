timeout 600 python -m pytest tests -m gpu -x -q -k "regex or thompson or concurr or degenerate" > gpurun_out/pytest_gpu5.log 2>&1; tail -3 gpurun_out/pytest_gpu5.log
FMX_TRACE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r01f.json 2> gpurun_out/bench_r01f.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r01f.json'))
print(d['value'], d['e2e']['value'])
print(d['regex'])
PY
grep "regex_set_search" gpurun_out/bench_r01f.err | tail -12
