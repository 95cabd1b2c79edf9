#!/bin/bash
mkdir -p gpurun_out/rq; O=gpurun_out/rq
echo "== decoder tests"; timeout 900 python -m pytest tests/test_gpu_decoder.py -m gpu -q -x 2>&1 | tail -2
echo "== decoder timing"; timeout 600 python tools/dec_timing.py 64 2>&1 | tee $O/dec_timing.txt | tail -16 | cut -c1-260
