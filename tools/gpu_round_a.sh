#!/bin/bash
# one GPU call: TMA probe, search-window modes, full GPU test-suite, sanitizers, stage statistics, a short bench
mkdir -p gpurun_out; O=gpurun_out/ra; mkdir -p $O
echo "== tma probe"; timeout 120 tools/_build/tma_probe > $O/tma_probe.txt 2>&1; cat $O/tma_probe.txt
for m in 2 1 3 0; do
  echo "== window mode $m: small encode + decode"
  B2H264_ENC_WIN=$m timeout 300 python tools/sanitize_small.py > $O/small_win$m.txt 2>&1; tail -2 $O/small_win$m.txt
done
# which mode do we trust for the rest?  TMA (1) if it works, else TMA-global (3), else warp loads (2)
MODE=2
grep -q "sanitize_small ok" $O/small_win3.txt && MODE=3
grep -q "sanitize_small ok" $O/small_win1.txt && MODE=1
echo "== using window mode $MODE" | tee $O/mode.txt
export B2H264_ENC_WIN=$MODE
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $O/gpu_tests.txt; cat $O/gpu_tests.txt
echo "== stage statistics"
for m in $MODE 2 0; do
  B2H264_ENC_WIN=$m timeout 600 python tools/enc_stats.py 256 > $O/enc_stats_win$m.txt 2>&1; echo "-- mode $m"; tail -4 $O/enc_stats_win$m.txt
done
echo "== bench (short)"; timeout 1500 python bench.py --steps 8 --warmup 3 > $O/bench.json 2> $O/bench.err; cut -c1-3000 $O/bench.json; tail -5 $O/bench.err
echo "== sanitizers"; bash tools/sanitize.sh
