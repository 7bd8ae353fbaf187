#!/bin/bash
mkdir -p gpurun_out
# step 3 (after 2 warm-up steps): kernels of interest in launch order: chain fwd, bwd(L5), rowred(L5), bwd(L4), rowred(L4)
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"k_layer_rowred_tc|k_layer_bwd_tc|k_chain_fwd_tc" --launch-skip 22 --launch-count 5 -o gpurun_out/j_big3 -f python tools/step_only.py > gpurun_out/j.log 2>&1
ls -la gpurun_out/j_big3.ncu-rep; tail -n 3 gpurun_out/j.log
