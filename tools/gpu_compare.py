"""Debug driver (GPU box): encode a synthetic clip with the CUDA encoder and with the reference; compare per frame.
usage: python tools/gpu_compare.py w h n qp seed [repeats] [idr_at]"""
import os, sys, hashlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import h264lib
from make_encoder_golden import ref_encode
from openh264_b200.binding import BatchEncoder
w, h, n, qp, seed = [int(x) for x in sys.argv[1:6]]
rep = int(sys.argv[6]) if len(sys.argv) > 6 else 1
idr_at = int(sys.argv[7]) if len(sys.argv) > 7 else -1
clip = h264lib.synth_clip(w, h, n, seed=seed); fsz = w * h * 3 // 2
ref_bs, ref_fb, _ = ref_encode(clip, w, h, n, qp, 30.0)
ref_bs = bytes(ref_bs)
off = np.concatenate([[0], np.cumsum(ref_fb)])
for r in range(rep):
    enc = BatchEncoder(w, h, qp=qp, fps=30.0, n_streams=1)
    bad = []
    for f in range(n):
        bs, _ = enc.encode([clip[f * fsz:(f + 1) * fsz]])
        if idr_at < 0 and bytes(bs[0]) != ref_bs[off[f]:off[f + 1]]:
            a = np.frombuffer(bytes(bs[0]), np.uint8); b = np.frombuffer(ref_bs[off[f]:off[f + 1]], np.uint8)
            m = min(len(a), len(b)); d = np.nonzero(a[:m] != b[:m])[0]
            bad.append((f, len(a), len(b), int(d[0]) if len(d) else m))
    print("run", r, "mismatching frames (frame, our bytes, ref bytes, first diff):", bad)
    enc.close()
