#!/bin/bash
mkdir -p gpurun_out/rd; O=gpurun_out/rd
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee $O/gpu_tests.txt
echo "== MC+SAD roofline"; timeout 300 python tools/mc_sad_roofline.py | tee $O/mc_sad.txt
echo "== stage statistics (rolled loops build)"; timeout 600 python tools/enc_stats.py 256 > $O/enc_stats.txt 2>&1; grep "^frame [45]" $O/enc_stats.txt | cut -c1-220; grep "batches" $O/enc_stats.txt | tail -1 | cut -c1-400
echo "== bench"; timeout 2400 python bench.py --steps 10 --warmup 3 --no-hard --no-decode > $O/bench.json 2> $O/bench.err; python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/rd/bench.json"))
    print("value %.0f e2e(api) %.0f e2e_l2 %.0f ms/step %.1f" % (d["value"], d["e2e"]["value"], d["e2e_layer2"]["value"], d["ms_per_step"]), d["breakdown_ms_per_step"], "parity", d["parity_checked"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/rd/bench.err").read()[-1500:])
PY
