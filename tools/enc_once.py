"""Encode a few 1080p frames on S streams (profiling driver for ncu)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import h264lib
from openh264_b200.binding import BatchEncoder
W, H = 1920, 1080
S = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4
clip = h264lib.synth_clip(W, H, N); fsz = W * H * 3 // 2
enc = BatchEncoder(W, H, qp=26, fps=30.0, n_streams=S)
for f in range(N):
    enc.encode([clip[f * fsz:(f + 1) * fsz]] * S)
print("us", enc.timing_us())
