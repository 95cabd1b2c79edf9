#!/bin/bash
mkdir -p gpurun_out/rab; O=gpurun_out/rab
echo "== pytest -m gpu"; timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee $O/gpu_tests.txt
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.txt
echo "== bench (no api / decode / cpu baseline)"; timeout 400 python bench.py --steps 8 --warmup 3 --no-api --no-decode --no-cpu-baseline > $O/bench.json 2> $O/bench.err; python -c "
import json; d=json.load(open('$O/bench.json')); print('value %.0f e2e_l2 %.0f parity %s skip %.3f hard %.0f hard_skip %.3f'%(d['value'], d['e2e_layer2']['value'], d.get('parity_checked'), d['config']['skip_ratio'], d['config']['workload_hard']['value'], d['config']['workload_hard']['skip_ratio']))" || tail -5 $O/bench.err
