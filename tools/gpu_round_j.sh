#!/bin/bash
mkdir -p gpurun_out/rj; O=gpurun_out/rj
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee $O/gpu_tests.txt
echo "== stage statistics 256"; timeout 600 python tools/enc_stats.py 256 > $O/enc_stats256.txt 2>&1; grep -E "^frame 5|batches" $O/enc_stats256.txt | tail -2 | cut -c1-700
echo "== stage statistics 256 (before the single-site searches; control lines + back-off only)"; B2H264_LIB=$PWD/tools/_build/lib_before_unify.so timeout 600 python tools/enc_stats.py 256 > $O/enc_stats256_before.txt 2>&1; grep -E "^frame 5|batches" $O/enc_stats256_before.txt | tail -2 | cut -c1-700
for v in main before; do
  L=""; [ $v = before ] && L=$PWD/tools/_build/lib_before_unify.so
  X="--no-api --no-parity --no-cpu-baseline --no-decode"; [ $v = main ] && X=""
  B2H264_LIB=$L timeout 2400 python bench.py --steps 10 --warmup 3 --no-hard $X > $O/bench_$v.json 2> $O/bench_$v.err
  python -c "
import json; d=json.load(open('$O/bench_$v.json')); print('$v: value %.0f e2e %.0f e2e_l2 %.0f'%(d['value'], d['e2e']['value'], d['e2e_layer2']['value']), d['breakdown_ms_per_step'], d.get('parity_checked'), (d.get('decode') or {}).get('value'))" || tail -5 $O/bench_$v.err
done
