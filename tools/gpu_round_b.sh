#!/bin/bash
# GPU round B: TMA probe matrix, CTA-shape variants, stream-count sweep, ncu captures of the two wavefront kernels
mkdir -p gpurun_out/rb; O=gpurun_out/rb
echo "== tma probe matrix"
for cfg in "2 48 48" "2 64 48" "2 128 48" "3 64 48" "3 128 48" "3 48 48" "3 48 32" "3 32 32" "3 16 16"; do
  timeout 60 tools/_build/tma_probe $cfg 0 2>&1 | head -2
done | tee $O/tma_probe.txt
export B2H264_ENC_WIN=2
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee $O/gpu_tests.txt
echo "== MC+SAD roofline"; timeout 300 python tools/mc_sad_roofline.py | tee $O/mc_sad.txt
echo "== CTA shape variants (256 streams)"
for v in "" openh264_b200/variants/lib_wpc12x2.so openh264_b200/variants/lib_wpc8x3.so; do
  B2H264_LIB=$v timeout 600 python tools/enc_stats.py 256 > $O/enc_stats_$(basename "${v:-default}").txt 2>&1
  echo "-- ${v:-default}"; grep "^frame [45]" $O/enc_stats_$(basename "${v:-default}").txt | cut -c1-200
done
echo "== stream-count sweep (layer 2 only)"
for S in 384 512 768; do
  timeout 900 python bench.py --streams $S --steps 6 --warmup 3 --no-api --no-hard --no-parity --no-cpu-baseline > $O/bench_S$S.json 2> $O/bench_S$S.err
  python - <<PY
import json
try:
    d=json.load(open("$O/bench_S$S.json")); print("S=$S value %.0f e2e_l2 %.0f ms/step %.1f enc %.1f dbk %.1f ent %.1f" % (d["value"], d["e2e_layer2"]["value"], d["ms_per_step"], d["breakdown_ms_per_step"]["encode_kernel"], d["breakdown_ms_per_step"]["deblock_expand"], d["breakdown_ms_per_step"]["host_entropy"]))
except Exception as e: print("S=$S failed", e)
PY
done
echo "== ncu (64 streams, second picture = P)"
timeout 1500 ncu --set full --import-source on --clock-control none -k regex:"k_encode_mbs|k_deblock_mbs" -s 2 -c 2 -o $O/enc64 -f python tools/enc_once.py 64 3 > $O/ncu.log 2>&1; tail -3 $O/ncu.log
ncu -i $O/enc64.ncu-rep --page raw --csv > $O/enc64_raw.csv 2>/dev/null; ls -la $O | head -20
