#!/bin/bash
mkdir -p gpurun_out/rx; O=gpurun_out/rx
echo "== small encode + decode (default: sync in A and B)"; timeout 120 python tools/sanitize_small.py 2>&1 | tail -1 | tee $O/small.txt
grep -q "sanitize_small ok" $O/small.txt || { echo "ABORT"; exit 1; }
for v in 0 1 3 2 3b; do
  export B2H264_BATCH_SYNC=${v%b}
  timeout 300 python bench.py --steps 8 --warmup 3 --no-hard --no-api --no-cpu-baseline --no-decode > $O/bench_$v.json 2> $O/bench_$v.err
  python -c "
import json; d=json.load(open('$O/bench_$v.json')); print('sync=$v: value %.0f e2e_l2 %.0f parity %s'%(d['value'], d['e2e_layer2']['value'], d.get('parity_checked')), d['breakdown_ms_per_step'])" || { tail -3 $O/bench_$v.err; echo "ABORT: bench $v"; exit 1; }
done
