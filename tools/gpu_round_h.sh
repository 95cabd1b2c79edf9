#!/bin/bash
mkdir -p gpurun_out/rh; O=gpurun_out/rh
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee $O/gpu_tests.txt
echo "== MC+SAD roofline"; timeout 300 python tools/mc_sad_roofline.py | tee $O/mc_sad.txt
echo "== stage statistics 256 (stages inlined)"; timeout 600 python tools/enc_stats.py 256 > $O/enc_stats256.txt 2>&1; grep -E "^frame 5|batches" $O/enc_stats256.txt | tail -2 | cut -c1-700
echo "== stage statistics 256 (stages as calls)"; B2H264_LIB=$PWD/tools/_build/lib_stage_calls.so timeout 600 python tools/enc_stats.py 256 > $O/enc_stats256_calls.txt 2>&1; grep -E "^frame 5|batches" $O/enc_stats256_calls.txt | tail -2 | cut -c1-700
for v in main calls; do
  L=""; [ $v = calls ] && L=$PWD/tools/_build/lib_stage_calls.so
  B2H264_LIB=$L timeout 1200 python bench.py --steps 10 --warmup 3 --no-hard --no-decode --no-api --no-parity --no-cpu-baseline > $O/bench_$v.json 2> $O/bench_$v.err
  python -c "
import json; d=json.load(open('$O/bench_$v.json')); print('$v: value %.0f e2e_l2 %.0f'%(d['value'], d['e2e_layer2']['value']), d['breakdown_ms_per_step'])" || tail -5 $O/bench_$v.err
done
