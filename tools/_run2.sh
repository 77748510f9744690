timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu3.log 2>&1; tail -3 gpurun_out/pytest_gpu3.log
python tools/exp_count.py --configs auto --lanes 0,1,2 --out gpurun_out/exp_v3.jsonl 2>&1 | grep '"count"\|"open"\|stats'
for mb in 5 7 8; do echo MINB=$mb; FMX_MINB=$mb python tools/exp_count.py --configs auto --lanes 1 --out gpurun_out/exp_v3_minb$mb.jsonl 2>&1 | grep '"count"'; done
