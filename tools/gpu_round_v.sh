#!/bin/bash
# every step is bounded tightly: a hang must not eat the GPU budget
mkdir -p gpurun_out/rv; O=gpurun_out/rv
echo "== small encode + decode (resident deblocking CTAs on)"; timeout 120 python tools/sanitize_small.py 2>&1 | tail -2 | tee $O/small.txt
grep -q "sanitize_small ok" $O/small.txt || { echo "ABORT: the small case failed or hung"; exit 1; }
echo "== the same under compute-sanitizer memcheck (kernels serialised: the resident CTAs must step aside)"; timeout 240 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2>&1 | tail -3 | tee $O/memcheck.txt
grep -q "sanitize_small ok" $O/memcheck.txt || { echo "ABORT: hang under the sanitizer"; exit 1; }
echo "== pytest -m gpu (encoder)"; timeout 400 python -m pytest tests/test_gpu_encoder.py -m gpu -q -x 2>&1 | tail -3 | tee $O/gpu_tests.txt
grep -q " passed" $O/gpu_tests.txt || { echo "ABORT: encoder tests"; exit 1; }
for v in resident serial resident2 serial2; do
  unset B2H264_NO_RESIDENT_DEBLOCK; case $v in serial*) export B2H264_NO_RESIDENT_DEBLOCK=1;; esac
  timeout 300 python bench.py --steps 8 --warmup 3 --no-hard --no-api --no-cpu-baseline --no-decode > $O/bench_$v.json 2> $O/bench_$v.err
  python -c "
import json; d=json.load(open('$O/bench_$v.json')); print('$v: value %.0f e2e_l2 %.0f parity %s'%(d['value'], d['e2e_layer2']['value'], d.get('parity_checked')), d['breakdown_ms_per_step'])" || { tail -3 $O/bench_$v.err; echo "ABORT: bench $v"; exit 1; }
done
