#!/bin/bash
mkdir -p gpurun_out/rw; O=gpurun_out/rw
echo "== small encode + decode with B2H264_BATCH_SYNC"; B2H264_BATCH_SYNC=1 timeout 120 python tools/sanitize_small.py 2>&1 | tail -1 | tee $O/small.txt
grep -q "sanitize_small ok" $O/small.txt || { echo "ABORT: hang or failure with batch sync"; exit 1; }
for v in base sync base2 sync2; do
  unset B2H264_BATCH_SYNC; case $v in sync*) export B2H264_BATCH_SYNC=1;; esac
  timeout 300 python bench.py --steps 8 --warmup 3 --no-hard --no-api --no-cpu-baseline --no-decode > $O/bench_$v.json 2> $O/bench_$v.err
  python -c "
import json; d=json.load(open('$O/bench_$v.json')); print('$v: value %.0f e2e_l2 %.0f parity %s'%(d['value'], d['e2e_layer2']['value'], d.get('parity_checked')), d['breakdown_ms_per_step'])" || { tail -3 $O/bench_$v.err; echo "ABORT: bench $v"; exit 1; }
done
