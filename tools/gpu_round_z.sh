#!/bin/bash
# final-style GPU round: tests, roofline units, ncu launch list + full captures, the full bench line, reference arm, sanitizers
mkdir -p gpurun_out/rz; O=gpurun_out/rz
echo "== pytest -m gpu"; timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee $O/gpu_tests.txt
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.txt
echo "== MC+SAD roofline"; timeout 300 python tools/mc_sad_roofline.py | tee $O/mc_sad.txt
echo "== stage statistics"; timeout 200 python tools/enc_stats.py 256 > $O/enc_stats.txt 2>&1; grep "^frame [45]" $O/enc_stats.txt | cut -c1-220
echo "== ncu launch list of a short bench run"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-api --no-hard --no-parity --no-cpu-baseline --no-decode > $O/bench_under_ncu.log 2>&1; wc -l $O/launches.csv
echo "== ncu --set full: k_encode_mbs at 256 streams (P picture), k_deblock_rows, k_mc_sad_tma"
timeout 500 ncu --set full --import-source on --clock-control none -k regex:"k_encode_mbs|k_deblock_rows" -s 2 -c 2 -o $O/enc256 -f python tools/enc_once.py 256 3 > $O/ncu_enc.log 2>&1; tail -2 $O/ncu_enc.log
timeout 200 ncu --set full --clock-control none -k regex:k_mc_sad_tma -s 2 -c 1 -o $O/mcsad -f python tools/mc_sad_once.py > $O/ncu_mc.log 2>&1; tail -1 $O/ncu_mc.log
for r in enc256 mcsad; do python tools/ncu_summary.py $O/$r.ncu-rep > $O/$r.txt 2>/dev/null; done; head -30 $O/enc256.txt
echo "== bench (default flags)"; timeout 700 python bench.py > $O/bench.json 2> $O/bench.err; cut -c1-6000 $O/bench.json; tail -3 $O/bench.err
echo "== reference arm"; timeout 300 python bench.py --impl reference > $O/bench_ref.json 2>/dev/null; cut -c1-1500 $O/bench_ref.json
echo "== sanitizers"; bash tools/sanitize.sh; cp gpurun_out/sanitize_*.log $O/
rm -f $O/enc256.ncu-rep.tmp
