#!/bin/bash
mkdir -p gpurun_out/rac; O=gpurun_out/rac
echo "== pytest -m gpu"; timeout 420 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee $O/gpu_tests.txt
