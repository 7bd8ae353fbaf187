#!/bin/bash
mkdir -p gpurun_out
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_prepA --launch-skip 2 --launch-count 1 -o gpurun_out/i_prep_algo1 -f python tools/prep_only.py prep_algo=1 > gpurun_out/i1.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_prepA --launch-skip 2 --launch-count 1 -o gpurun_out/i_prep_algo2 -f python tools/prep_only.py prep_algo=2 > gpurun_out/i2.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/i1.log gpurun_out/i2.log
