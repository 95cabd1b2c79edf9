#!/bin/bash
mkdir -p gpurun_out/rg; O=gpurun_out/rg
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee $O/gpu_tests.txt
echo "== stage statistics 256 (claim retry)"; timeout 600 python tools/enc_stats.py 256 > $O/enc_stats256.txt 2>&1; grep -E "^frame 5|batches" $O/enc_stats256.txt | tail -3 | cut -c1-700
echo "== stage statistics 256 (legacy claim)"; B2H264_LEGACY_CLAIM=1 timeout 600 python tools/enc_stats.py 256 > $O/enc_stats256_legacy.txt 2>&1; grep -E "^frame 5|batches" $O/enc_stats256_legacy.txt | tail -3 | cut -c1-700
echo "== bench"; timeout 2400 python bench.py --steps 10 --warmup 3 --no-hard --no-decode > $O/bench.json 2> $O/bench.err; python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/rg/bench.json"))
    print("value %.0f e2e(api) %.0f e2e_l2 %.0f ms/step %.1f" % (d["value"], d["e2e"]["value"], d["e2e_layer2"]["value"], d["ms_per_step"]), d["breakdown_ms_per_step"], "parity", d["parity_checked"], "d2h", d["e2e"]["d2h_bytes_per_step"])
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/rg/bench.err").read()[-1500:])
PY
echo "== bench legacy claim"; B2H264_LEGACY_CLAIM=1 timeout 1200 python bench.py --steps 10 --warmup 3 --no-hard --no-decode --no-api --no-parity --no-cpu-baseline > $O/bench_legacy.json 2> $O/bench_legacy.err; python -c "
import json; d=json.load(open('gpurun_out/rg/bench_legacy.json')); print('legacy claim: value %.0f'%d['value'], d['breakdown_ms_per_step'])"
