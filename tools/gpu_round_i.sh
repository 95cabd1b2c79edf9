#!/bin/bash
mkdir -p gpurun_out/ri; O=gpurun_out/ri
echo "== pytest -m gpu (encoder)"; timeout 1500 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -4 | tee $O/gpu_tests.txt
echo "== stage statistics 256 (scheduler warp, search inlined)"; timeout 600 python tools/enc_stats.py 256 > $O/enc_stats256.txt 2>&1; grep -E "^frame 5|batches" $O/enc_stats256.txt | tail -2 | cut -c1-700
echo "== stage statistics 256 (search as call)"; B2H264_LIB=$PWD/tools/_build/lib_me_call.so timeout 600 python tools/enc_stats.py 256 > $O/enc_stats256_mecall.txt 2>&1; grep -E "^frame 5|batches" $O/enc_stats256_mecall.txt | tail -2 | cut -c1-700
for v in main mecall legacy; do
  L=""; [ $v = mecall ] && L=$PWD/tools/_build/lib_me_call.so
  unset B2H264_LEGACY_CLAIM; [ $v = legacy ] && export B2H264_LEGACY_CLAIM=1
  B2H264_LIB=$L timeout 1200 python bench.py --steps 10 --warmup 3 --no-hard --no-decode --no-api --no-parity --no-cpu-baseline > $O/bench_$v.json 2> $O/bench_$v.err
  python -c "
import json; d=json.load(open('$O/bench_$v.json')); print('$v: value %.0f e2e_l2 %.0f'%(d['value'], d['e2e_layer2']['value']), d['breakdown_ms_per_step'])" || tail -5 $O/bench_$v.err
done
