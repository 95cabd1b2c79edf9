"""Decoder phase timing (B2H264_DEC_TIMING=1): S copies of a 1080p stream, pageable vs page-locked output."""
import os, sys, time
os.environ["B2H264_DEC_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from openh264_b200.binding import BatchEncoder, BatchDecoder
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
clip = bench.make_clip(); seq = bench.ping_pong(bench.CLIP_FRAMES)
enc = BatchEncoder(bench.W, bench.H, qp=bench.QP, fps=bench.FPS, n_streams=1)
aus = []
for i in range(8):
    f = bench.stream_frame(seq, 0, i)
    bs, _ = enc.encode([clip[f * bench.FSZ:(f + 1) * bench.FSZ]])
    aus.append(bytes(bs[0]))
enc.close()
for pinned in (False, True):
    dec = BatchDecoder(bench.W, bench.H, n_streams=S, pinned_output=pinned)
    dec.decode([aus[0]] * S); dec.decode([aus[1]] * S)
    t0 = time.perf_counter()
    for au in aus[2:]:
        dec.decode([au] * S)
    dt = time.perf_counter() - t0
    print("pinned_output=%s: %.0f pictures/s (%.1f ms per call of %d)" % (pinned, 6 * S / dt, dt / 6 * 1e3, S), flush=True)
    dec.close()
