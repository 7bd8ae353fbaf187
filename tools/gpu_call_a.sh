#!/bin/bash
# GPU call A (this session): first exposure of the new CUDA paths (natgrad, predict epilogues, full_cov, DGP_Quad weights,
# configs 4/5) + the existing parity suite (the likelihood kernels and run_step were touched).  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/a_env.txt 2>&1
rm -f gpurun_out/parity_errors.jsonl
( time timeout 400 python -m pytest tests/test_gpu_natgrad.py tests/test_gpu_predict.py tests/test_gpu_full_cov.py tests/test_gpu_quad.py tests/test_gpu_configs45.py -m gpu -q -rA --tb=short --durations=15 -p no:cacheprovider ) > gpurun_out/a_new_tests.log 2>&1
echo "new tests exit: $?" >> gpurun_out/a_new_tests.log
( time timeout 420 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_multi.py -m gpu -q --tb=short --durations=10 -p no:cacheprovider ) > gpurun_out/a_old_tests.log 2>&1
echo "old tests exit: $?" >> gpurun_out/a_old_tests.log
( time timeout 150 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_gpu_full_cov.py::test_full_cov_propagate_matches_oracle[0-False]" "tests/test_gpu_natgrad.py::test_small_gamma_every_layer_multi_output[False-0]" "tests/test_gpu_predict.py::test_predict_y_and_density[multiclass-0]" -m gpu -q --tb=short -p no:cacheprovider ) > gpurun_out/a_sanitizer.log 2>&1
echo "sanitizer exit: $?" >> gpurun_out/a_sanitizer.log
tail -5 gpurun_out/a_new_tests.log; tail -5 gpurun_out/a_old_tests.log; tail -5 gpurun_out/a_sanitizer.log
