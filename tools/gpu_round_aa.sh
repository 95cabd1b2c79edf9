#!/bin/bash
mkdir -p gpurun_out/raa
timeout 600 python tools/api_sweep.py 2>&1 | tee gpurun_out/raa/api_sweep.txt | cut -c1-300
