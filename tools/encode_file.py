"""Encode a raw I420 file with the CUDA encoder (the counterpart of the reference's `h264enc -org in.yuv ...` for
the supported configuration: constant QP, single layer / slice, CAVLC).
usage: python tools/encode_file.py in.yuv width height out.264 [--qp 26] [--fps 30] [--frames N]"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from openh264_b200.binding import BatchEncoder  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("yuv"); ap.add_argument("width", type=int); ap.add_argument("height", type=int); ap.add_argument("out")
    ap.add_argument("--qp", type=int, default=26); ap.add_argument("--fps", type=float, default=30.0)
    ap.add_argument("--frames", type=int, default=0)
    a = ap.parse_args()
    fsz = a.width * a.height * 3 // 2
    data = np.fromfile(a.yuv, dtype=np.uint8)
    n = data.size // fsz if not a.frames else min(a.frames, data.size // fsz)
    enc = BatchEncoder(a.width, a.height, qp=a.qp, fps=a.fps, n_streams=1)
    total = 0
    with open(a.out, "wb") as f:
        enc.submit([data[:fsz]])
        for i in range(1, n + 1):                      # pipelined: picture i is submitted before picture i-1 is collected
            if i < n:
                enc.submit([data[i * fsz:(i + 1) * fsz]])
            bs, _ = enc.collect()
            f.write(bytes(bs[0]))
            total += len(bs[0])
    enc.close()
    print("%d pictures -> %d bytes (%s)" % (n, total, a.out))


if __name__ == "__main__":
    main()
