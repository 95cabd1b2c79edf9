#!/bin/bash
mkdir -p gpurun_out/rk; O=gpurun_out/rk
echo "== pytest -m gpu (encoder + kernels)"; timeout 1500 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -3 | tee $O/gpu_tests.txt
for v in main roundj main2 roundj2; do
  L=""; case $v in roundj*) L=$PWD/tools/_build/lib_round_j.so;; esac
  B2H264_LIB=$L timeout 1200 python bench.py --steps 10 --warmup 3 --no-hard --no-api --no-parity --no-cpu-baseline --no-decode > $O/bench_$v.json 2> $O/bench_$v.err
  python -c "
import json; d=json.load(open('$O/bench_$v.json')); print('$v: value %.0f e2e_l2 %.0f'%(d['value'], d['e2e_layer2']['value']), d['breakdown_ms_per_step'])" || tail -5 $O/bench_$v.err
done
echo "== stage statistics 256"; timeout 600 python tools/enc_stats.py 256 > $O/enc_stats256.txt 2>&1; grep -E "^frame 5|batches" $O/enc_stats256.txt | tail -2 | cut -c1-700
