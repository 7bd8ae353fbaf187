#!/bin/bash
mkdir -p gpurun_out
( timeout 200 python tools/ab_bench.py "" "overlap=0" ) > gpurun_out/k_ab.log 2>&1
cat gpurun_out/k_ab.log
( timeout 300 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider ) > gpurun_out/k_tests.log 2>&1
tail -8 gpurun_out/k_tests.log
