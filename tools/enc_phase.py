"""Profiling build only (make EXTRA_NVFLAGS=-DB2H264_PHASE_STATS): cycles per phase of the MB pipeline."""
import os, sys, ctypes as C, numpy as np
os.environ["B2H264_ENC_STATS"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import h264lib
from openh264_b200.binding import BatchEncoder, lib
W, H = 1920, 1080
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
clip = h264lib.synth_clip(W, H, 5); fsz = W * H * 3 // 2
enc = BatchEncoder(W, H, qp=26, fps=30.0, n_streams=S)
L = lib(0)
ph = np.zeros(32, np.uint64); st = np.zeros(16, np.uint64)
names = {31: "sched/idle->start", 0: "load nb+cur+borders", 1: "inter cache", 2: "pskip test", 3: "P16x16 ME", 4: "I16 MD", 5: "intra enc",
         6: "sub-partition ME", 7: "refine+chroma MC", 8: "residual+recon", 12: "  skip: mvp+checks", 13: "  skip: luma MC", 14: "  skip: SAD+chroma MC+SAD", 15: "  skip: dct tests", 16: "  skip: SATD", 10: "bookkeeping", 11: "store+publish"}
for f in range(5):
    enc.encode([clip[f * fsz:(f + 1) * fsz]] * S)
    L.b2h264_debug_phase_stats(ph.ctypes.data, 1); L.b2h264_debug_enc_stats(st.ctypes.data, 1)
    n = int(sum(st[1::2]))
    print("frame", f, "kernel us", round(enc.timing_us()[0]), "MBs", n)
    tot = float(ph.sum())
    for k in sorted(names):
        if ph[k]: print("   %-22s %8.0f cyc/MB  %5.1f%%" % (names[k], ph[k] / n, 100 * ph[k] / tot))
