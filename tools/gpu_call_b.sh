#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python bench.py --steps 50 --warmup 5 ) > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
( timeout 120 python tools/dbg_stamps.py 2 102 ) > gpurun_out/b_stamps.out 2> gpurun_out/b_stamps.err
tail -c 3000 gpurun_out/b_bench.json; tail -60 gpurun_out/b_stamps.err
