"""Writes tests/golden/encoder_edge.json: SHA-1 of the UNMODIFIED reference encoder's bitstream (oracle/_ref,
ref_shim.cpp:ref_encode, constant QP, complexity HIGH) for
  * BASELINE.json configs[1]: the reference's own res/CiscoVT2people_320x192_12fps.yuv (committed as
    tests/golden/CiscoVT2people_320x192_12fps.yuv: a reference-held test vector, 9 pictures), QP 26 and 34;
  * the edge-case list the reference's encoder tests sweep (test/api/encoder_test.cpp:103-180 resolution / QP
    sweep): QP 0 / 51, pictures that need cropping, 16x16, very wide / very tall pictures.
    python tests/golden/make_encoder_edge_golden.py
"""
import hashlib, json, os, shutil, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import h264lib  # noqa: E402
from make_encoder_golden import ref_encode  # noqa: E402

CLIP = "CiscoVT2people_320x192_12fps.yuv"
EDGE_CASES = [  # (w, h, n, qp, seed)
    (176, 144, 4, 0, 3), (176, 144, 4, 51, 3), (176, 144, 4, 12, 4), (176, 144, 4, 45, 4),
    (180, 148, 4, 26, 5), (164, 130, 4, 30, 6), (16, 16, 4, 26, 7), (32, 18, 4, 20, 8),
    (480, 32, 4, 28, 9), (64, 256, 4, 33, 10), (352, 288, 3, 38, 11),
]

LOW_CASES = [  # (w, h, n, qp, seed, noise)
    (640, 360, 4, 28, 4, 7), (180, 148, 4, 22, 5, 3), (176, 130, 5, 24, 3, 6), (352, 288, 5, 18, 4, 8), (1280, 720, 3, 30, 2, 6),
    (32, 18, 4, 20, 8, 3), (176, 144, 4, 51, 3, 3),
]

CABAC_CASES = [  # (w, h, n, qp, seed, noise, profile_idc: 0 = the encoder's choice (High), 77 = Main)
    (176, 144, 4, 26, 3, 3, 0), (176, 144, 4, 26, 3, 3, 77), (320, 192, 5, 32, 4, 3, 0), (64, 64, 3, 12, 5, 3, 77), (180, 148, 4, 26, 5, 3, 0),
    (16, 16, 4, 26, 7, 3, 0), (32, 18, 4, 20, 8, 3, 77), (176, 144, 4, 0, 3, 3, 0), (176, 144, 4, 51, 3, 3, 0), (480, 32, 4, 28, 9, 3, 77),
    (320, 192, 4, 24, 60, 40, 0), (640, 360, 3, 30, 2, 6, 0),
]

if __name__ == "__main__":
    assert h264lib.have_ref()
    src = os.path.join("/root/reference/res", CLIP)
    if not os.path.exists(os.path.join(HERE, CLIP)):
        shutil.copyfile(src, os.path.join(HERE, CLIP))
    gold = {"clip": {}, "edge": {}}
    yuv = np.fromfile(os.path.join(HERE, CLIP), dtype=np.uint8)
    for qp in (26, 34):
        bs, fb, _ = ref_encode(yuv, 320, 192, 9, qp, 12.0)
        gold["clip"]["qp%d" % qp] = {"sha1": hashlib.sha1(bs).hexdigest(), "frame_bytes": fb,
                                     "yuv_sha1": hashlib.sha1(yuv.tobytes()).hexdigest()}
    for (w, h, n, qp, seed) in EDGE_CASES:
        y = h264lib.synth_clip(w, h, n, seed=seed)
        bs, fb, _ = ref_encode(y, w, h, n, qp, 30.0)
        gold["edge"]["%dx%d_n%d_qp%d_seed%d" % (w, h, n, qp, seed)] = {"sha1": hashlib.sha1(bs).hexdigest(), "frame_bytes": fb}
    # iComplexityMode = LOW_COMPLEXITY (the reference's default; BASELINE.json configs[1] as SURVEY 8d defines it)
    gold["low"] = {}
    for qp in (26, 34):
        bs, fb, _ = ref_encode(yuv, 320, 192, 9, qp, 12.0, complexity=0)
        gold["low"]["clip_qp%d" % qp] = {"sha1": hashlib.sha1(bs).hexdigest(), "frame_bytes": fb}
    for (w, h, n, qp, seed, noise) in LOW_CASES:
        y = h264lib.synth_clip(w, h, n, seed=seed, noise=noise)
        bs, fb, _ = ref_encode(y, w, h, n, qp, 30.0, complexity=0)
        gold["low"]["%dx%d_n%d_qp%d_seed%d_noise%d" % (w, h, n, qp, seed, noise)] = {"sha1": hashlib.sha1(bs).hexdigest(), "frame_bytes": fb}
    # iEntropyCodingModeFlag = 1 (CABAC; High profile unless the layer asks for Main)
    gold["cabac"] = {}
    for prof in (0, 77):
        bs, fb, _ = ref_encode(yuv, 320, 192, 9, 26, 12.0, entropy=(1, prof))
        gold["cabac"]["clip_qp26_profile%d" % prof] = {"sha1": hashlib.sha1(bs).hexdigest(), "frame_bytes": fb}
    for (w, h, n, qp, seed, noise, prof) in CABAC_CASES:
        y = h264lib.synth_clip(w, h, n, seed=seed, noise=noise)
        bs, fb, _ = ref_encode(y, w, h, n, qp, 30.0, entropy=(1, prof))
        gold["cabac"]["%dx%d_n%d_qp%d_seed%d_noise%d_profile%d" % (w, h, n, qp, seed, noise, prof)] = {"sha1": hashlib.sha1(bs).hexdigest(), "frame_bytes": fb}
    bs, fb, _ = ref_encode(yuv, 320, 192, 9, 30, 12.0, complexity=0, entropy=(1, 0))
    gold["cabac"]["clip_qp30_profile0_low"] = {"sha1": hashlib.sha1(bs).hexdigest(), "frame_bytes": fb}
    # uiIntraPeriod: periodic IDR pictures (with the INCREASING_ID parameter-set strategy every IDR brings new SPS / PPS ids)
    gold["intra_period"] = {}
    for (w, h, n, qp, seed, period, ent) in [(176, 144, 8, 27, 3, 3, (0, 66)), (320, 192, 7, 31, 4, 2, (1, 0)), (64, 64, 4, 20, 5, 1, (0, 66))]:
        y = h264lib.synth_clip(w, h, n, seed=seed)
        bs, fb, _ = ref_encode(y, w, h, n, qp, 30.0, entropy=ent, intra_period=period)
        gold["intra_period"]["%dx%d_n%d_qp%d_seed%d_period%d_cabac%d" % (w, h, n, qp, seed, period, ent[0])] = {"sha1": hashlib.sha1(bs).hexdigest(), "frame_bytes": fb}
    # iLoopFilterDisableIdc / iLoopFilterAlphaC0Offset / iLoopFilterBetaOffset
    gold["loop_filter"] = {}
    for (w, h, n, qp, seed, lf, ent) in [(176, 144, 5, 30, 3, (1, 0, 0), (0, 66)), (176, 144, 5, 30, 3, (0, 3, -2), (0, 66)), (320, 192, 4, 36, 4, (2, -6, 6), (1, 0)),
                                           (64, 64, 4, 24, 5, (0, 6, 6), (0, 66)), (180, 148, 4, 40, 6, (0, -4, -5), (0, 66))]:
        y = h264lib.synth_clip(w, h, n, seed=seed, noise=6)
        bs, fb, _ = ref_encode(y, w, h, n, qp, 30.0, entropy=ent, loop_filter=lf)
        gold["loop_filter"]["%dx%d_n%d_qp%d_seed%d_idc%d_a%d_b%d_cabac%d" % (w, h, n, qp, seed, lf[0], lf[1], lf[2], ent[0])] = {"sha1": hashlib.sha1(bs).hexdigest(), "frame_bytes": fb}
    json.dump(gold, open(os.path.join(HERE, "encoder_edge.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(gold, indent=1))
