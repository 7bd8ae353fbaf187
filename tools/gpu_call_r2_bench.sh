#!/bin/bash
# round-2 bench lines at HEAD: the contract line (config 3, north star) and the other BASELINE configs, plus the reference arm
mkdir -p gpurun_out
for c in 3 2 4 5; do
  timeout 300 python bench.py --config $c > gpurun_out/r2final_bench_config$c.json 2> gpurun_out/r2final_bench_config$c.err
  cut -c1-420 gpurun_out/r2final_bench_config$c.json
done
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2final_bench_reference.json 2> gpurun_out/r2final_bench_reference.err
cut -c1-600 gpurun_out/r2final_bench_reference.json
