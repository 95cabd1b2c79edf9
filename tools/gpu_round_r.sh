#!/bin/bash
mkdir -p gpurun_out/rr; O=gpurun_out/rr
echo "== 2-GPU bench"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-hard > $O/bench2.json 2> $O/bench2.err; tail -1 $O/bench2.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('2 GPUs: value %.0f e2e %.0f e2e_l2 %.0f'%(d['value'], d['e2e']['value'], d['e2e_layer2']['value']), d['breakdown_ms_per_step'], 'decode', (d.get('decode') or {}).get('value'))" || tail -5 $O/bench2.err
echo "== 1-GPU bench (same box)"; timeout 1500 python bench.py --steps 10 --warmup 3 --no-hard > $O/bench1.json 2> $O/bench1.err; tail -1 $O/bench1.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('1 GPU: value %.0f e2e %.0f e2e_l2 %.0f'%(d['value'], d['e2e']['value'], d['e2e_layer2']['value']), d['breakdown_ms_per_step'], 'decode', (d.get('decode') or {}).get('value'))" || tail -5 $O/bench1.err
