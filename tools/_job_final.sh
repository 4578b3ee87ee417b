#!/bin/bash
# final 1-GPU job of the round: GPU tests, default bench (all extras), reference arm, launch list, sanitizer
mkdir -p gpurun_out
(time timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/gputest_final.log 2>&1
(time python bench.py --steps 20 --warmup 5) > gpurun_out/bench_final_1gpu.json 2> gpurun_out/bench_final_1gpu.err
(time python bench.py --impl reference --steps 3 --warmup 1) > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 3 --ecs-per-gpu 3000000 --no-cpu-baseline --extras none > gpurun_out/ncu_launches_final.log 2>&1
timeout 600 bash tools/_job_san.sh > gpurun_out/sanitizer_final.txt 2>&1
tail -3 gpurun_out/gputest_final.log
