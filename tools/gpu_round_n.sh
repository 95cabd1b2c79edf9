#!/bin/bash
mkdir -p gpurun_out/rn; O=gpurun_out/rn
echo "== stage statistics 256"; timeout 600 python tools/enc_stats.py 256 > $O/enc_stats256.txt 2>&1; grep -E "^frame 5|batches|task wall" $O/enc_stats256.txt | tail -3 | cut -c1-700
echo "== api sweep"; timeout 1500 python tools/api_sweep.py 2>&1 | tee $O/api_sweep.txt | cut -c1-300
