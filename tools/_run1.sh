set -x
python bench.py --steps 200 --warmup 5 > gpurun_out/bench_r01d.json 2> gpurun_out/bench_r01d.err; tail -c 3000 gpurun_out/bench_r01d.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 3 --warmup 3 --no-cpu --regexes 0 > gpurun_out/ncu_launch4.log 2>&1
SKIP=$(python tools/ncu_skip.py gpurun_out/launches_r01d.csv 10000000 2); echo skip=$SKIP
ncu --set full --clock-control none --import-source on -k regex:count_fixed_kernel -s $SKIP -c 2 -o gpurun_out/prof_count_r01d -f python bench.py --steps 3 --warmup 3 --no-cpu --regexes 0 > gpurun_out/ncu_full4.log 2>&1; tail -3 gpurun_out/ncu_full4.log
FMX_MINB=6 python tools/exp_count.py --configs auto --lanes 2,1 --out gpurun_out/exp_minb6.jsonl 2>&1 | grep '"count"'
FMX_MINB=4 python tools/exp_count.py --configs auto --lanes 2,1 --out gpurun_out/exp_minb4.jsonl 2>&1 | grep '"count"'
FMX_PERSISTENT=1 python tools/exp_count.py --configs auto --lanes 2 --out gpurun_out/exp_persist.jsonl 2>&1 | grep '"count"'
python tools/exp_count.py --n 500000000 --configs auto --lanes 2 --out gpurun_out/exp_half.jsonl 2>&1 | grep '"count"\|"open"\|stats'
