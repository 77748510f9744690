timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu8.log 2>&1; tail -12 gpurun_out/pytest_gpu8.log
timeout 900 python bench.py --workload cfg5 --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_cfg5_n1_r01i.json 2> gpurun_out/bench_cfg5_n1_r01i.err; tail -4 gpurun_out/bench_cfg5_n1_r01i.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_cfg5_n1_r01i.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'], d['config'])
PY
