#!/bin/bash
# tools/ncu_capture.sh — ncu --set full captures of the four hot kernels on their BASELINE workloads, reduced to text on the GPU box
# (gpurun brings back at most 64 MiB; a report with sources is ~28 MB).  Run under gpurun from the repo root.
#   details csv  = every section's metrics     raw csv = all raw metrics (dram__bytes_*, stalls, ...)
set -u
N="ncu --set full --clock-control none --import-source on"
cap() {   # name kernel-regex launch-skip what
  timeout 600 $N -k regex:$2 -c 1 --launch-skip $3 -f -o /tmp/$1 python tools/profile_legs.py --what $4 --reps 3 2>&1 | grep -v "^==PROF==" | tail -2
  ncu -i /tmp/$1.ncu-rep --page details --csv > gpurun_out/$1_details.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ls -la /tmp/$1.ncu-rep
}
timeout 400 python tools/profile_legs.py --what regex --reps 5 --rx-local 0,16,32,64,128,256 2>&1 | grep what > gpurun_out/r02_regex_local_keep_sweep.jsonl
cat gpurun_out/r02_regex_local_keep_sweep.jsonl
cap r02_regex_queue_english regex_queue_kernel 1 regex
cp /tmp/r02_regex_queue_english.ncu-rep gpurun_out/ 2>/dev/null
cap r02_locate_english locate_kernel 1 locate
cap r02_count_english_len12_pass1 count_dict_first_kernel 2 count3
cap r02_count_english_len12_pass2 count_list_kernel 2 count3
cap r02_count_cfg2_len16 count_fixed_kernel 2 count2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_bench_launches_gpu_time.csv \
    python bench.py --steps 20 --warmup 5 --legs "" --no-cpu > gpurun_out/r02_bench_under_ncu.json 2> gpurun_out/r02_bench_under_ncu.err
echo ncu-bench rc=$?
du -sh gpurun_out
