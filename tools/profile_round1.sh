#!/bin/bash
# Round-1 evidence run (one GPU): default bench line, reference arm, ncu launch list of the bench command, full ncu
# captures of the tcgen05 kernels.  Outputs land in gpurun_out/ (copied into profiles/ afterwards).
mkdir -p gpurun_out
python bench.py --steps 6 --warmup 3 --layers-out gpurun_out/layers_r1.json 2> gpurun_out/bench_r1.err | tail -1 > gpurun_out/bench_r1_n1.json
cut -c1-300 gpurun_out/bench_r1_n1.json
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_r1_reference.json
cut -c1-200 gpurun_out/bench_r1_reference.json
# launch list of the bench command (1 warm-up + 1 timed step under ncu; skip model construction + warm-up launches)
ncu --metrics gpu__time_duration.sum --clock-control none -s 1120 -c 1100 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-layers > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches_r1.csv
for spec in "fprop_3x3_patch_128_28:128 128 3 1 28 512 fprop:conv3x3_patch" "fprop_1x1_256_1024_14:256 1024 1 1 14 512 fprop:conv_igemm" \
            "wgrad_3x3_128_28:128 128 3 1 28 512 wgrad:conv3x3_wgrad" "dgrad_1x1_1024_256_14:256 1024 1 1 14 512 dgrad:conv_igemm"; do
  name=${spec%%:*}; rest=${spec#*:}; args=${rest%%:*}; kern=${rest##*:}
  ncu --set full --clock-control none --import-source on -k regex:$kern -s 2 -c 1 -o gpurun_out/ncu_$name \
      python tools/conv_case.py $args > /dev/null 2>&1
done
for k in stem_fprop stem_wgrad; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o gpurun_out/ncu_$k python tools/time_stem.py 512 > /dev/null 2>&1
done
ls -la gpurun_out/ncu_*.ncu-rep
