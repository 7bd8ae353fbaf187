#!/bin/bash
mkdir -p gpurun_out
( timeout 100 python bench.py --steps 50 --warmup 5 --cpu-steps 8 ) > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
tail -c 600 gpurun_out/final_bench.json
timeout 75 ncu --metrics gpu__time_duration.sum --clock-control none -c 140 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/final_ncu.log 2>&1
wc -l gpurun_out/final_launches.csv
