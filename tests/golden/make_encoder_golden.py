"""Writes tests/golden/encoder.json: SHA-1 of the bitstream the UNMODIFIED reference encoder (oracle/_ref,
public API, constant QP, camera mode, single slice, complexity HIGH — oracle/ref_shim.cpp:ref_encode) produces
for the deterministic integer clips of tests/h264lib.py:synth_clip, plus per-frame sizes.  These are the
golden vectors the GPU encoder is held to on machines without the reference.
    python tests/golden/make_encoder_golden.py
"""
import ctypes as C, hashlib, json, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import h264lib  # noqa: E402

CASES = [  # (w, h, frames, qp, fps)
    (320, 192, 8, 26, 12.0), (176, 144, 6, 20, 15.0), (640, 360, 5, 30, 30.0), (1280, 720, 4, 32, 30.0),
    (1920, 1080, 4, 26, 30.0), (64, 64, 5, 10, 30.0), (320, 192, 4, 40, 30.0),
]


def ref_encode(yuv, w, h, n, qp, fps, complexity=2, threads=1, entropy=(0, 66), intra_period=0, loop_filter=(0, 0, 0)):
    """entropy = (iEntropyCodingModeFlag, uiProfileIdc; 0 = leave the profile to the encoder: High with CABAC)"""
    R = C.CDLL(h264lib.REFSHIM_SO)
    R.ref_set_entropy.argtypes = [C.c_int, C.c_int]
    R.ref_encode.restype = C.c_long
    R.ref_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p,
                             C.c_long, C.c_void_p, C.POINTER(C.c_double)]
    cap = 64 << 20
    out, fb, secs = np.zeros(cap, np.uint8), np.zeros(n, np.int32), C.c_double()
    R.ref_set_entropy(*entropy)
    R.ref_set_intra_period(intra_period)
    R.ref_set_loop_filter(*loop_filter)
    try:
        tot = R.ref_encode(yuv.ctypes.data, w, h, n, qp, complexity, threads, fps, out.ctypes.data, cap, fb.ctypes.data, C.byref(secs))
    finally:
        R.ref_set_entropy(0, 66)
        R.ref_set_intra_period(0)
        R.ref_set_loop_filter(0, 0, 0)
    assert tot > 0
    return out[:tot].tobytes(), fb.tolist(), secs.value


if __name__ == "__main__":
    assert h264lib.have_ref()
    gold = {}
    for (w, h, n, qp, fps) in CASES:
        yuv = h264lib.synth_clip(w, h, n)
        bs, fb, _ = ref_encode(yuv, w, h, n, qp, fps)
        gold[f"{w}x{h}_n{n}_qp{qp}_fps{fps:g}"] = {"sha1": hashlib.sha1(bs).hexdigest(), "frame_bytes": fb,
                                                  "yuv_sha1": hashlib.sha1(yuv.tobytes()).hexdigest()}
    json.dump(gold, open(os.path.join(HERE, "encoder.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(gold, indent=1))
