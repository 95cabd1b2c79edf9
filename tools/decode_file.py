"""Decode an Annex-B file with the CUDA decoder construct path (the counterpart of the reference's
`h264dec in.264 out.yuv` for the supported stream class: Baseline, CAVLC, one slice per picture, one reference).
usage: python tools/decode_file.py in.264 width height out.yuv"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from openh264_b200.binding import BatchDecoder  # noqa: E402


def access_units(bs):
    """splits an Annex-B byte string into access units: each ends with its (single) slice NAL (types 1 / 5)"""
    starts = [i for i in range(len(bs) - 3) if bs[i] == 0 and bs[i + 1] == 0 and bs[i + 2] == 1]
    begin = None
    for k, s in enumerate(starts):
        first = s - 1 if s > 0 and bs[s - 1] == 0 else s
        if begin is None:
            begin = first
        if bs[s + 3] & 31 in (1, 5):
            nxt = starts[k + 1] if k + 1 < len(starts) else len(bs)
            end = nxt - 1 if nxt < len(bs) and bs[nxt - 1] == 0 else nxt
            yield bs[begin:end]
            begin = None


def main():
    path, w, h, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    bs = open(path, "rb").read()
    dec = BatchDecoder(w, h, n_streams=1)
    n = 0
    with open(out, "wb") as f:
        for au in access_units(bs):
            f.write(dec.decode([au])[0].tobytes())
            n += 1
    dec.close()
    print("%d pictures -> %s" % (n, out))


if __name__ == "__main__":
    main()
