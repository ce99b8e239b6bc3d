#!/bin/bash
# Round-2 evidence run (one GPU).  Outputs land in gpurun_out/ (summaries are copied into profiles/ afterwards).
#   1. default bench line (CUDA graphs on) with the per-layer table and the CPU arm
#   2. the other BASELINE.json configurations that fit one GPU (fp32-accurate RN50 b256, RN50 @384 b256, RN200 b256)
#   3. ncu launch list of ONE eager step of the bench command with per-launch DRAM bytes (roofline.traffic)
mkdir -p gpurun_out
timeout 400 python bench.py --steps 20 --warmup 5 --layers-out gpurun_out/layers_r2.json 2> gpurun_out/bench_r2.err | tail -1 > gpurun_out/bench_r2_n1.json
cut -c1-260 gpurun_out/bench_r2_n1.json
timeout 300 python bench.py --steps 5 --warmup 3 --no-layers --no-cpu-baseline --precision fp32 --batch-per-gpu 256 2>/dev/null | tail -1 > gpurun_out/bench_r2_fp32_b256.json
cut -c1-200 gpurun_out/bench_r2_fp32_b256.json
timeout 300 python bench.py --steps 5 --warmup 3 --no-layers --no-cpu-baseline --image-size 384 --batch-per-gpu 256 2>/dev/null | tail -1 > gpurun_out/bench_r2_rn50_384_b256.json
cut -c1-200 gpurun_out/bench_r2_rn50_384_b256.json
timeout 300 python bench.py --steps 5 --warmup 3 --no-layers --no-cpu-baseline --arch resnet200 --batch-per-gpu 256 2>/dev/null | tail -1 > gpurun_out/bench_r2_rn200_b256.json
cut -c1-200 gpurun_out/bench_r2_rn200_b256.json
# one eager step under ncu: skip construction + the warm-up step, capture the timed step (+ the start of the e2e loop)
BYOL_B200_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -s 1100 -c 1150 --csv --log-file gpurun_out/launches_r2.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-layers > gpurun_out/bench_under_ncu_r2.log 2>&1
wc -l gpurun_out/launches_r2.csv
