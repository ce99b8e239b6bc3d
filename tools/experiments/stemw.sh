V=none
for v in 0 1; do
  if BYOL_STEM_WGRAD_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_conv.py -m gpu -x -q --tb=line -p no:cacheprovider -k "stem4_wgrad" > gpurun_out/stemw_v$v.log 2>&1; then V=$v; break; fi
done
echo "passing variant: $V"; tail -3 gpurun_out/stemw_v0.log | cut -c1-300
if [ "$V" != none ]; then
  export BYOL_STEM_WGRAD_VARIANT=$V
  timeout 120 python tools/time_stem.py 2>&1 | grep -v Warn
  timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_blocks.py tests/test_gpu_step.py -m gpu -x -q --tb=short -p no:cacheprovider 2>&1 | grep -v Warn | tail -6
  timeout 300 python bench.py --steps 6 --warmup 3 --no-layers --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_b512_v22.json
  python -c "
import json; b=json.load(open('gpurun_out/bench_b512_v22.json')); print(round(b['value']), round(b['ms_per_step'],2), round(b['e2e']['value']), b['gpu_launches'])"
fi
