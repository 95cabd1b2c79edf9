#!/bin/bash
mkdir -p gpurun_out/rt; O=gpurun_out/rt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee $O/gpu_tests.txt
