#!/bin/bash
mkdir -p gpurun_out
( timeout 200 python tools/ab_bench.py "prep_algo=1" "prep_algo=2,prep_threads=512" "prep_algo=2,prep_threads=256" "prep_algo=2,prep_threads=1024" ) > gpurun_out/g_ab.log 2>&1
cat gpurun_out/g_ab.log
( timeout 300 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider ) > gpurun_out/g_tests.log 2>&1
tail -8 gpurun_out/g_tests.log
