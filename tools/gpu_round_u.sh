#!/bin/bash
mkdir -p gpurun_out/ru; O=gpurun_out/ru
echo "== pytest -m gpu (encoder, API)"; timeout 1500 python -m pytest tests/test_gpu_encoder.py tests/test_wels_api.py tests/test_gpu_decoder.py -m gpu -q -x 2>&1 | tail -4 | tee $O/gpu_tests.txt
for v in resident serial resident2 serial2; do
  unset B2H264_NO_RESIDENT_DEBLOCK; case $v in serial*) export B2H264_NO_RESIDENT_DEBLOCK=1;; esac
  timeout 1200 python bench.py --steps 10 --warmup 3 --no-hard --no-api --no-cpu-baseline --no-decode > $O/bench_$v.json 2> $O/bench_$v.err
  python -c "
import json; d=json.load(open('$O/bench_$v.json')); print('$v: value %.0f e2e_l2 %.0f parity %s'%(d['value'], d['e2e_layer2']['value'], d.get('parity_checked')), d['breakdown_ms_per_step'])" || tail -5 $O/bench_$v.err
done
unset B2H264_NO_RESIDENT_DEBLOCK
echo "== sanitizers"; bash tools/sanitize.sh 2>&1 | grep -E "SUMMARY|exit=|ok:" | head -12
