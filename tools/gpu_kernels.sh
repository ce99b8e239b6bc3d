#!/bin/bash
# Runs the per-kernel GPU parity tests in separate processes (a trapped kernel poisons its CUDA context).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; timeout 900 python -m pytest "$@" -m gpu -q -s --tb=short --timeout=300 -p no:cacheprovider > gpurun_out/k_$name.log 2>&1; echo "$name exit=$?"; tail -3 gpurun_out/k_$name.log; }
run simple tests/test_gpu_simple_kernels.py
run fprop_auto tests/test_gpu_conv.py -k "test_conv_fprop and auto"
run fprop_gather tests/test_gpu_conv.py -k "test_conv_fprop and gather"
run epilogue tests/test_gpu_conv.py -k "test_conv_fprop_epilogue"
run dgrad tests/test_gpu_conv.py -k "test_conv_dgrad"
run wgrad_auto tests/test_gpu_conv.py -k "test_conv_wgrad and auto"
run wgrad_gather tests/test_gpu_conv.py -k "test_conv_wgrad and gather"
run linear tests/test_gpu_conv.py -k "test_linear"
