#!/bin/bash
mkdir -p gpurun_out/rp; O=gpurun_out/rp
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee $O/gpu_tests.txt
echo "== bench (no hard / api)"; timeout 2400 python bench.py --steps 10 --warmup 3 --no-hard --no-api > $O/bench.json 2> $O/bench.err; python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/rp/bench.json"))
    print("value %.0f e2e_l2 %.0f ms/step %.1f" % (d["value"], d["e2e_layer2"]["value"], d["ms_per_step"]), d["breakdown_ms_per_step"], "parity", d["parity_checked"], "decode", d["decode"]["value"], d["decode"].get("roofline"))
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/rp/bench.err").read()[-1500:])
PY
echo "== racecheck"; bash tools/sanitize.sh > $O/sanitize.txt 2>&1; grep -E "SUMMARY|exit=" gpurun_out/sanitize_*.log $O/sanitize.txt 2>/dev/null | head
