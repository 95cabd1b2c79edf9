#!/bin/bash
# debug: run the API driver with both libraries on a synthetic clip, keep outputs in gpurun_out/
cd /root/repo
python - <<'PY'
import sys; sys.path.insert(0,'tests')
import h264lib
h264lib.synth_clip(176,144,6,seed=7).tofile('/tmp/in.yuv')
PY
mkdir -p gpurun_out/wels
oracle/_ref/wels_driver oracle/_ref/libopenh264_ref.so /tmp/in.yuv 176 144 6 26 3 gpurun_out/wels/ref.264 gpurun_out/wels/ref.layout
oracle/_ref/wels_driver openh264_b200/libopenh264_b200_wels.so /tmp/in.yuv 176 144 6 26 3 gpurun_out/wels/b2.264 gpurun_out/wels/b2.layout
diff gpurun_out/wels/ref.layout gpurun_out/wels/b2.layout && echo LAYOUT SAME
cmp gpurun_out/wels/ref.264 gpurun_out/wels/b2.264 && echo BITSTREAM SAME
