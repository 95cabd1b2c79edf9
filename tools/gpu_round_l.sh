#!/bin/bash
mkdir -p gpurun_out/rl; O=gpurun_out/rl
echo "== ncu --set full: k_encode_mbs at 256 streams (P picture)"
timeout 1500 ncu --set full --import-source on --clock-control none -k regex:"k_encode_mbs" -s 1 -c 1 -o $O/enc256 -f python tools/enc_once.py 256 3 > $O/ncu_enc.log 2>&1; tail -2 $O/ncu_enc.log
python tools/ncu_summary.py $O/enc256.ncu-rep > $O/enc256.txt 2>/dev/null; head -40 $O/enc256.txt
