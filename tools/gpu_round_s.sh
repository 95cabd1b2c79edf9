#!/bin/bash
mkdir -p gpurun_out/rs; O=gpurun_out/rs
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee $O/gpu_tests.txt
echo "== many decoder objects (API)"; for T in 1 16 64; do B2H264_BROKER_SLOTS=64 timeout 600 oracle/_ref/wels_mt_dec_driver openh264_b200/libopenh264_b200_wels.so $T 3 - tests/golden/conformance/Zhling_1280x720.264 2>&1 | tail -1; done | tee $O/api_dec.txt
echo "== reference, same driver"; for T in 1 16; do timeout 600 oracle/_ref/wels_mt_dec_driver oracle/_ref/libopenh264_ref.so $T 3 - tests/golden/conformance/Zhling_1280x720.264 2>&1 | tail -1; done | tee -a $O/api_dec.txt
