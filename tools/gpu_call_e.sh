#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/e_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/e_bench_under_ncu.log 2>&1
tail -3 gpurun_out/e_bench_under_ncu.log | cut -c1-300
