set -x
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r01h.json 2> gpurun_out/bench_ref_r01h.err; tail -c 600 gpurun_out/bench_ref_r01h.json
python bench.py --steps 200 --warmup 5 > gpurun_out/bench_r01h.json 2> gpurun_out/bench_r01h.err; head -c 300 gpurun_out/bench_r01h.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r01h.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_launch5.log 2>&1
SKIP=$(python tools/ncu_skip.py gpurun_out/launches_r01h.csv 10000000 1); echo skip=$SKIP
ncu --set full --clock-control none --import-source on -k regex:count_fixed_kernel -s $SKIP -c 2 -o gpurun_out/prof_count_r01h -f python bench.py --steps 3 --warmup 3 --no-cpu --regexes 0 > gpurun_out/ncu_full5.log 2>&1; tail -2 gpurun_out/ncu_full5.log
ncu --set full --clock-control none --import-source on -k regex:regex_search_kernel -s 3 -c 1 -o gpurun_out/prof_regex_r01h -f python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_full6.log 2>&1; tail -2 gpurun_out/ncu_full6.log
timeout 900 python tools/bench_configs.py --out gpurun_out/configs_r01h.jsonl > gpurun_out/configs_r01h.log 2>&1; tail -12 gpurun_out/configs_r01h.log | cut -c 1-600
