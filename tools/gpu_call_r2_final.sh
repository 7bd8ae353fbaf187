#!/bin/bash
# final verification of the round at HEAD: smoke, the GPU suite, a memcheck pass over one tcgen05-path gradient test (tile
# hand-over included), the contract bench line and the launch list of one bench step
mkdir -p gpurun_out
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 )
( timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 )
( timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_gpu_parity.py::test_gradients_match_autograd" -k "4-False-1 or 4-True-1" -q -x -p no:cacheprovider > gpurun_out/r2final_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2final_memcheck.log | tail -3 )
timeout 300 python bench.py > gpurun_out/r2final_bench_config3.json 2> gpurun_out/r2final_bench_config3.err
cut -c1-300 gpurun_out/r2final_bench_config3.json
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r2final_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2final_ncu_launches.log 2>&1
wc -l gpurun_out/r2final_launches.csv
