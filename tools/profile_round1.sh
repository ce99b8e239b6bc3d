#!/bin/bash
# Round-1 evidence run (one GPU): default bench line, ncu launch list of the same command, full ncu captures of the
# two tcgen05 kernels.  Outputs land in gpurun_out/ (copied into profiles/ afterwards).
mkdir -p gpurun_out
python bench.py --steps 6 --warmup 3 --layers-out gpurun_out/layers_r1.json 2> gpurun_out/bench_r1.err | tail -1 > gpurun_out/bench_r1_n1.json
cut -c1-300 gpurun_out/bench_r1_n1.json
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_r1_reference.json
cut -c1-200 gpurun_out/bench_r1_reference.json
# launch list of the bench command (1 warm-up + 1 timed step under ncu; skip model construction + warm-up launches)
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1400 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-layers > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches_r1.csv
# full captures: 3x3 fprop (gather path), 1x1 fprop (TMA path), 3x3 wgrad, 1x1 wgrad at the bench batch
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 2 -c 1 -o gpurun_out/ncu_fprop_3x3_128_28 \
    python tools/conv_case.py 128 128 3 1 28 512 fprop > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 2 -c 1 -o gpurun_out/ncu_fprop_1x1_256_1024_14 \
    python tools/conv_case.py 256 1024 1 1 14 512 fprop > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -s 2 -c 1 -o gpurun_out/ncu_wgrad_3x3_128_28 \
    python tools/conv_case.py 128 128 3 1 28 512 wgrad > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 2 -c 1 -o gpurun_out/ncu_dgrad_3x3_128_28 \
    python tools/conv_case.py 128 128 3 1 28 512 dgrad > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
