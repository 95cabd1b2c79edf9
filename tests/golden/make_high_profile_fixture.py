"""Cuts the first access units of two of the reference's High-profile B-frame vectors (res/VID_1280x544_{cabac,cavlc}_temporal_direct.264:
x264 streams with the 8x8 transform, Intra_8x8, explicit weighted P prediction, B pyramids, temporal direct prediction, implicit
weights — the tool set of BASELINE.json configs[3]'s 1080p stream, at a size small enough to commit) into tests/golden/conformance_b/
and records the SHA-1 of what the UNMODIFIED reference decoder (oracle/_ref, ISVCDecoder::DecodeFrameNoDelay + flush) makes of each
prefix.  The whole vectors are checked against the published hashes where /root/reference is present (tests/test_decoder_emu.py)."""
import ctypes as C
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
import h264lib  # noqa: E402

N_AU = 14


def main():
    R = C.CDLL(h264lib.REFSHIM_SO)
    R.ref_decode.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]
    out = {}
    for name, n_au in (("VID_1280x544_cabac_temporal_direct.264", N_AU), ("VID_1280x544_cavlc_temporal_direct.264", N_AU),
                       ("VID_1920x1080_cabac_temporal_direct.264", 10)):      # the last one: BASELINE.json configs[3]'s stream itself
        bs = open("/root/reference/res/" + name, "rb").read()
        aus = h264lib.split_access_units(bs)
        prefix = b"".join(aus[:n_au])
        dst = name.replace(".264", "_first%d.264" % n_au)
        open(os.path.join(HERE, "conformance_b", dst), "wb").write(prefix)
        a = np.frombuffer(prefix, np.uint8)
        buf = np.zeros(64 << 20, np.uint8)
        W, H, s = C.c_int(), C.c_int(), C.c_double()
        n = R.ref_decode(a.ctypes.data, len(a), buf.ctypes.data, buf.size, C.byref(W), C.byref(H), C.byref(s))
        assert n == n_au, n
        out[dst] = {"sha1": hashlib.sha1(buf[:n * W.value * H.value * 3 // 2].tobytes()).hexdigest(), "pictures": n,
                    "width": W.value, "height": H.value, "bytes": len(prefix)}
    json.dump(out, open(os.path.join(HERE, "high_profile_prefix.json"), "w"), indent=1)
    print(out)


if __name__ == "__main__":
    main()
