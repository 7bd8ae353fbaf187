#!/bin/bash
# round-2 evidence at HEAD: ncu --set full of the three tile kernels inside a step (third eager step, caches as the step leaves
# them), the launch list of one bench step, and a back-to-back timing
mkdir -p gpurun_out
timeout 100 python tools/ab_bench.py "" 2>&1 | grep us/step
timeout 400 ncu --set full --clock-control none --cache-control none --import-source on \
    -k regex:"k_layer_rowred_tc|k_layer_bwd_tc|k_chain_fwd_tc" --launch-skip 22 --launch-count 11 \
    -f -o gpurun_out/r2final_big3 python tools/step_only.py > gpurun_out/r2final_big3.log 2>&1
tail -3 gpurun_out/r2final_big3.log
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r2final_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2final_ncu_launches.log 2>&1
wc -l gpurun_out/r2final_launches.csv
ls -la gpurun_out/r2final_big3.ncu-rep
