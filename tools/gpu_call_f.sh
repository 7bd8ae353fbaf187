#!/bin/bash
mkdir -p gpurun_out
( timeout 200 python tools/ab_bench.py "prep_algo=1,fin_per_layer=0" "prep_algo=1" "prep_algo=2,prep_threads=512" "prep_algo=2,prep_threads=256" ) > gpurun_out/f_ab.log 2>&1
cat gpurun_out/f_ab.log
( timeout 300 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider ) > gpurun_out/f_tests.log 2>&1
tail -15 gpurun_out/f_tests.log
