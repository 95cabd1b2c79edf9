"""prints bench.py's roofline_mc_sad block alone (BASELINE metric 2: MC+SAD unit against the measured HBM peak)"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from openh264_b200.binding import lib
torch.cuda.set_device(0)
peak, _ = bench.measured_peaks()
r = bench.mc_sad_roofline(lib(0), 0, peak)
for k in ("integer_mv", "quarter_pel_mv", "mixed_9_candidates"):
    print(k, json.dumps(r[k]))
