"""Launch the MC+SAD unit on 64 stacked 1080p planes a few times (profiling driver for ncu)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from openh264_b200.binding import lib, check
L = lib(0)
S, stride, rows_per = 64, 2048, 1152
g = torch.Generator(device="cuda"); g.manual_seed(264)
cur = torch.randint(0, 256, (S * rows_per, stride), dtype=torch.uint8, device="cuda", generator=g)
ref = torch.randint(0, 256, (S * rows_per, stride), dtype=torch.uint8, device="cuda", generator=g)
mbw, mbh = 120, (S * rows_per) // 16 - 4
n = mbw * mbh
mv = torch.randint(-8, 9, (n, 1, 2), dtype=torch.int16, device="cuda", generator=g) * 4
cost = torch.empty((n, 1), dtype=torch.int32, device="cuda")
o0 = 32 * stride + 32
for _ in range(4):
    check(L.b2h264_k_mc_sad(cur.data_ptr() + o0, stride, ref.data_ptr() + o0, stride, mbw, mbh, mv.data_ptr(), 1, cost.data_ptr(), None))
torch.cuda.synchronize()
print("ok", int(cost.sum().item()))
