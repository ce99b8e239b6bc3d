#!/bin/bash
# Round-2 ncu --set full captures (one GPU): the kernels that changed this round + the default bench line.
mkdir -p gpurun_out
timeout 400 python bench.py --steps 20 --warmup 5 --layers-out gpurun_out/layers_r2.json 2> gpurun_out/bench_r2.err | tail -1 > gpurun_out/bench_r2_n1.json
cut -c1-200 gpurun_out/bench_r2_n1.json; grep "bench " gpurun_out/bench_r2.err | tail -3
N="--set full --clock-control none --import-source on"
timeout 200 ncu $N -k regex:gemm_fused_kernel -s 9 -c 1 -o gpurun_out/ncu_r2_dgrad_resid_64_56 python tools/time_dgrad_resid.py 64 56 512 > /dev/null 2>&1
timeout 200 ncu $N -k regex:bn_bwd_reduce_fixed -s 2 -c 1 -o gpurun_out/ncu_r2_bn_bwd_reduce python tools/time_fused.py 64 56 512 > /dev/null 2>&1
timeout 200 ncu $N -k regex:bn_apply_fixed -s 2 -c 1 -o gpurun_out/ncu_r2_bn_apply python tools/time_fused.py 64 56 512 > /dev/null 2>&1
timeout 200 ncu $N -k regex:conv_igemm_kernel -s 2 -c 1 -o gpurun_out/ncu_r2_fprop_1x1_64_256_56 python tools/time_fused.py 64 56 512 > /dev/null 2>&1
timeout 200 ncu $N -k regex:xchg_sum -c 1 -o gpurun_out/ncu_r2_dummy python -c "print(1)" > /dev/null 2>&1
ls -la gpurun_out/ncu_r2_*.ncu-rep
