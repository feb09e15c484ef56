#!/bin/bash
# compute-sanitizer passes over the kernel tests (the reference has no race detection at all).
# memcheck on the comm + BN + elementwise kernels, racecheck on the smem-using BN reductions, synccheck on the tcgen05 kernels.
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $CS --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "sgd_flat or push_reduce or batchnorm_eval or maxpool or softmax or k_of_p" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/sanitize_memcheck.log
timeout 600 $CS --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "batchnorm_eval or softmax" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/sanitize_racecheck.log
timeout 600 $CS --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "tc_gemm and 128-64-64" > gpurun_out/sanitize_tc.log 2>&1; echo "tc memcheck rc=$?" | tee -a gpurun_out/sanitize_tc.log
grep -hE "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_*.log | tail -12
