#!/bin/bash
# GPU round C: TMA alignment hypothesis, window modes again (64-wide aligned window), row deblocking, MC+SAD, full tests, bench
mkdir -p gpurun_out/rc; O=gpurun_out/rc
echo "== tma probe: aligned vs unaligned start"
for cfg in "2 64 48 0 32" "2 64 48 0 37" "3 64 48 0 32" "3 64 48 0 48" "3 64 48 1 32" "3 48 48 0 32" "3 64 48 0 36"; do
  timeout 60 tools/_build/tma_probe $cfg 2>&1 | head -2
done | tee $O/tma_probe.txt
for m in 1 3 2; do
  echo "== window mode $m: small encode + decode"
  B2H264_ENC_WIN=$m timeout 300 python tools/sanitize_small.py > $O/small_win$m.txt 2>&1; tail -1 $O/small_win$m.txt
done
MODE=2
grep -q "sanitize_small ok" $O/small_win3.txt && MODE=3
grep -q "sanitize_small ok" $O/small_win1.txt && MODE=1
echo "== using window mode $MODE" | tee $O/mode.txt
export B2H264_ENC_WIN=$MODE
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $O/gpu_tests.txt
echo "== MC+SAD roofline"; timeout 300 python tools/mc_sad_roofline.py | tee $O/mc_sad.txt
echo "== stage statistics"; timeout 600 python tools/enc_stats.py 256 > $O/enc_stats.txt 2>&1; grep "^frame [45]" $O/enc_stats.txt | cut -c1-220
echo "== bench"; timeout 2400 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err; python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/rc/bench.json"))
    print("value %.0f e2e(api) %.0f e2e_l2 %.0f ms/step %.1f" % (d["value"], d["e2e"]["value"], d["e2e_layer2"]["value"], d["ms_per_step"]), d["breakdown_ms_per_step"], "parity", d["parity_checked"])
    print("hard", d["config"]["workload_hard"]); print("decode", d.get("decode")); print("mc_sad", {k: round(v["frac"],3) for k,v in d["roofline_mc_sad"].items() if isinstance(v,dict)})
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/rc/bench.err").read()[-1500:])
PY
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_ref.json 2>/dev/null; cut -c1-400 $O/bench_ref.json
echo "== sanitizers"; bash tools/sanitize.sh; cp gpurun_out/sanitize_*.log $O/
