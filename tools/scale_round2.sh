#!/bin/bash
# Round-2 multi-GPU evidence on ONE 8 x B200 lease: byol_b200 at N = 8 / 4 / 2 (weak scaling, 512 images/GPU, SyncBN
# statistics over peer memory, CUDA graphs) and the UNMODIFIED reference at N = 8 (DDP + SyncBatchNorm, autocast bf16).
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
for n in 8 4 2; do
  timeout 240 bash -c "$(declare -f run); run $n $((29600 + n)) bench.py --gpus $n --steps 10 --warmup 3" \
      2> gpurun_out/scale_r2_n$n.err | tail -1 > gpurun_out/scale_r2_n$n.json
  cut -c1-200 gpurun_out/scale_r2_n$n.json
done
rm -f gpurun_out/ref_gpu_r2_n8.jsonl
timeout 400 bash -c "$(declare -f run); run 8 29650 tools/ref_gpu_baseline.py --mode bf16 --batch-per-gpu 512 --sync-bn --steps 5 --warmup 2 --out gpurun_out/ref_gpu_r2_n8.jsonl" > gpurun_out/ref_gpu_r2_n8.log 2>&1
if [ ! -s gpurun_out/ref_gpu_r2_n8.jsonl ]; then
  timeout 300 bash -c "$(declare -f run); run 8 29651 tools/ref_gpu_baseline.py --mode bf16 --batch-per-gpu 256 --sync-bn --steps 5 --warmup 2 --out gpurun_out/ref_gpu_r2_n8.jsonl" >> gpurun_out/ref_gpu_r2_n8.log 2>&1
fi
cut -c1-400 gpurun_out/ref_gpu_r2_n8.jsonl; tail -2 gpurun_out/ref_gpu_r2_n8.log | cut -c1-300
