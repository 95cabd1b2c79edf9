#!/bin/bash
mkdir -p gpurun_out/rf; O=gpurun_out/rf
echo "== MC+SAD roofline"; timeout 300 python tools/mc_sad_roofline.py | tee $O/mc_sad.txt
echo "== stage statistics 256"; timeout 600 python tools/enc_stats.py 256 > $O/enc_stats256.txt 2>&1; grep -E "^frame 5|fill histogram|batches" $O/enc_stats256.txt | tail -8 | cut -c1-900
echo "== stage statistics 768"; timeout 900 python tools/enc_stats.py 768 > $O/enc_stats768.txt 2>&1; grep -E "^frame 5|fill histogram|batches" $O/enc_stats768.txt | tail -8 | cut -c1-900
echo "== api sweep"; timeout 1500 python tools/api_sweep.py 2>&1 | tee $O/api_sweep.txt | cut -c1-400
