"""Quick device-side timing of the layer-1 hot kernels (not the bench contract; see bench.py)."""
import sys, os, json, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import openh264_b200 as m
from openh264_b200.binding import check
import h264lib

L = m.lib(0)
w, h, pad = 1920, 1088, 32
stride = 2048
nstreams = int(sys.argv[1]) if len(sys.argv) > 1 else 16
k = int(sys.argv[2]) if len(sys.argv) > 2 else 16
rows = h + 2 * pad
cur = torch.empty((nstreams, rows, stride), dtype=torch.uint8, device="cuda")
ref = torch.empty_like(cur)
f0 = torch.from_numpy(h264lib.synth_frame(w, h, 0)).cuda()
f1 = torch.from_numpy(h264lib.synth_frame(w, h, 1)).cuda()
for s in range(nstreams):
    ref[s].zero_(); cur[s].zero_()
    ref[s, pad:pad + h, pad:pad + w] = f0
    cur[s, pad:pad + h, pad:pad + w] = f1
    check(L.b2h264_k_expand_plane(ref[s].data_ptr() + pad * stride + pad, stride, w, h, pad, None))
mbw, mbh = w // 16, h // 16
mv = torch.randint(-32, 33, (nstreams, mbw * mbh, k, 2), dtype=torch.int16, device="cuda")
cost = torch.empty((nstreams, mbw * mbh, k), dtype=torch.int32, device="cuda")
o0 = pad * stride + pad
st = torch.cuda.current_stream().cuda_stream
def run():
    for s in range(nstreams):
        check(L.b2h264_k_mc_sad(cur[s].data_ptr() + o0, stride, ref[s].data_ptr() + o0, stride, mbw, mbh,
                                mv[s].data_ptr(), k, cost[s].data_ptr(), st))
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 10
e0.record()
for _ in range(iters): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
mbs = nstreams * mbw * mbh
alg = mbs * (512 + 8 * k)   # cur 256 + ref 256 per MB, mv 4k in + cost 4k out
print(json.dumps({"kernel": "k_mc_sad", "streams": nstreams, "k": k, "ms": ms, "mb_per_s": mbs / ms * 1e3,
                  "cand_per_s": mbs * k / ms * 1e3, "alg_GBps": alg / ms / 1e6}))
