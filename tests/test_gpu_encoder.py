"""GPU parity tests, layer 2: the batched frame encoder through the C-ABI (include/b2h264_codec.h).
Bit-exact bitstreams against (i) golden SHA-1s generated from the unmodified reference
(tests/golden/encoder.json), (ii) the reference itself where oracle/_ref travelled with the repo;
plus size-independent properties at BASELINE.json's full 1080p size: the reference decoder
(oracle/_ref) decodes our stream to exactly our own reconstruction; identical streams in a batch give
identical bitstreams; pipelined (2 in flight) == synchronous."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import h264lib

pytestmark = pytest.mark.gpu
ROOT = h264lib.ROOT
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "encoder.json")))


def parse_key(key):
    w, h = map(int, key.split("_")[0].split("x"))
    return w, h, int(key.split("_n")[1].split("_")[0]), int(key.split("_qp")[1].split("_")[0]), float(key.split("_fps")[1])


def gpu_encode(yuv, w, h, n, qp, fps, n_streams=1, pipelined=False):
    from openh264_b200.binding import BatchEncoder
    enc = BatchEncoder(w, h, qp=qp, fps=fps, n_streams=n_streams)
    fsz = w * h * 3 // 2
    frames = [np.ascontiguousarray(yuv[i * fsz:(i + 1) * fsz]) for i in range(n)]
    outs = [[] for _ in range(n_streams)]
    if not pipelined:
        for f in frames:
            bs, _ = enc.encode([f] * n_streams)
            for s in range(n_streams):
                outs[s].append(bs[s])
    else:
        enc.submit([frames[0]] * n_streams)
        for i in range(1, n + 1):
            if i < n:
                enc.submit([frames[i]] * n_streams)
            bs, _ = enc.collect()
            for s in range(n_streams):
                outs[s].append(bs[s])
    rec = enc.recon(0)
    enc.close()
    return outs, rec


@pytest.mark.parametrize("key", sorted(GOLD))
def test_bitstream_matches_reference_golden(key):
    w, h, n, qp, fps = parse_key(key)
    yuv = h264lib.synth_clip(w, h, n)
    assert hashlib.sha1(yuv.tobytes()).hexdigest() == GOLD[key]["yuv_sha1"]
    outs, _ = gpu_encode(yuv, w, h, n, qp, fps)
    assert [len(b) for b in outs[0]] == GOLD[key]["frame_bytes"]
    assert hashlib.sha1(b"".join(outs[0])).hexdigest() == GOLD[key]["sha1"]


def test_batch_and_pipelining_are_transparent():
    w, h, n, qp, fps = 320, 192, 6, 26, 12.0
    yuv = h264lib.synth_clip(w, h, n)
    single, _ = gpu_encode(yuv, w, h, n, qp, fps)
    batch, _ = gpu_encode(yuv, w, h, n, qp, fps, n_streams=5, pipelined=True)
    for s in range(5):
        assert batch[s] == single[0]


@pytest.mark.parametrize("whn", [(320, 192, 6), (1920, 1080, 3)])
def test_reference_decoder_reproduces_our_reconstruction(whn):
    if not h264lib.have_ref():
        pytest.skip("oracle/_ref not present")
    w, h, n = whn
    yuv = h264lib.synth_clip(w, h, n, seed=77)
    outs, rec = gpu_encode(yuv, w, h, n, 28, 30.0)
    bs = np.frombuffer(b"".join(outs[0]), np.uint8)
    R = C.CDLL(h264lib.REFSHIM_SO)
    R.ref_decode.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]
    fsz = w * h * 3 // 2
    dec = np.zeros(n * fsz + 64, np.uint8)
    W, H, s = C.c_int(), C.c_int(), C.c_double()
    nf = R.ref_decode(bs.ctypes.data, len(bs), dec.ctypes.data, len(dec), C.byref(W), C.byref(H), C.byref(s))
    assert nf == n and W.value == w and H.value == h
    assert np.array_equal(dec[(n - 1) * fsz:n * fsz], rec)      # decoder output of the last frame == our deblocked recon


def test_against_reference_side_by_side():
    if not h264lib.have_ref():
        pytest.skip("oracle/_ref not present")
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_encoder_golden import ref_encode
    for (w, h, n, qp) in [(352, 288, 5, 24), (640, 480, 4, 36)]:
        yuv = h264lib.synth_clip(w, h, n, seed=5)
        ref_bs, ref_fb, _ = ref_encode(yuv, w, h, n, qp, 25.0)
        outs, _ = gpu_encode(yuv, w, h, n, qp, 25.0)
        assert b"".join(outs[0]) == ref_bs


def test_force_idr_and_bad_config():
    from openh264_b200.binding import BatchEncoder, B2H264Error
    w, h = 176, 144
    yuv = h264lib.synth_clip(w, h, 3)
    fsz = w * h * 3 // 2
    enc = BatchEncoder(w, h, qp=30, fps=15.0)
    _, t0 = enc.encode([yuv[:fsz]])
    _, t1 = enc.encode([yuv[fsz:2 * fsz]])
    enc.force_idr()
    bs2, t2 = enc.encode([yuv[2 * fsz:3 * fsz]])
    assert (t0[0], t1[0], t2[0]) == (1, 2, 1)
    assert bs2[0][4] & 31 == 7        # an IDR access unit starts with the SPS
    enc.close()
    with pytest.raises(B2H264Error):
        BatchEncoder(15, 15)


def test_heterogeneous_streams_with_one_forced_idr():
    """7 streams with different content in one encoder; stream 3 is forced to an IDR at picture 2 while the others
    code P pictures, so one launch carries IDR-picture macroblocks (list I) and P-picture macroblocks (lists A/B/C)
    together.  Every stream must equal what the reference's own API produces for that stream (API driver through
    the compiled reference, with the same ForceIntraFrame call)."""
    from test_wels_api import DRIVER, REFLIB, drive
    if not (os.path.exists(DRIVER) and os.path.exists(REFLIB)):
        pytest.skip("oracle/_ref not present")
    import tempfile
    from openh264_b200.binding import BatchEncoder
    w, h, n, qp, S = 176, 144, 5, 27, 7
    clips = [h264lib.synth_clip(w, h, n, seed=100 + s) for s in range(S)]
    fsz = w * h * 3 // 2
    enc = BatchEncoder(w, h, qp=qp, fps=30.0, n_streams=S)
    got = [b""] * S
    enc.submit([c[:fsz] for c in clips])
    for f in range(1, n + 1):
        if f < n:
            if f == 2:
                enc.force_idr(3)
            enc.submit([c[f * fsz:(f + 1) * fsz] for c in clips])
        bs, types = enc.collect()
        if f - 1 == 2:
            assert [int(t) for t in types] == [2, 2, 2, 1, 2, 2, 2]
        got = [got[s] + bytes(bs[s]) for s in range(S)]
    enc.close()
    with tempfile.TemporaryDirectory() as tmp:
        for s in range(S):
            r, ref_bs, _ = drive(REFLIB, clips[s], w, h, n, qp, 2 if s == 3 else -1, tmp, "ref%d" % s)
            assert r.returncode == 0, r.stderr
            assert got[s] == ref_bs, "stream %d differs from the reference" % s


# ---- BASELINE.json configs[1] and the edge-case sweep through the CUDA path ------------------------------------------
EDGE = json.load(open(os.path.join(ROOT, "tests", "golden", "encoder_edge.json")))


@pytest.mark.parametrize("qp", [26, 34])
def test_config1_reference_clip_through_cuda(qp):
    """BASELINE.json configs[1]: the reference's own res/CiscoVT2people_320x192_12fps.yuv (committed test vector),
    constant QP, single slice, through BatchEncoder on the GPU; golden = the unmodified reference's bitstream."""
    yuv = np.fromfile(os.path.join(ROOT, "tests", "golden", "CiscoVT2people_320x192_12fps.yuv"), dtype=np.uint8)
    g = EDGE["clip"]["qp%d" % qp]
    assert hashlib.sha1(yuv.tobytes()).hexdigest() == g["yuv_sha1"]
    outs, _ = gpu_encode(yuv, 320, 192, 9, qp, 12.0, n_streams=2, pipelined=True)
    assert [len(b) for b in outs[0]] == g["frame_bytes"]
    assert hashlib.sha1(b"".join(outs[0])).hexdigest() == g["sha1"]
    assert outs[1] == outs[0]


@pytest.mark.parametrize("key", sorted(EDGE["edge"]))
def test_edge_cases_through_cuda(key):
    """QP 0 / 51, pictures that need cropping, 16x16, very wide / tall pictures (the reference's encoder_test.cpp
    resolution / QP sweep) on the GPU; golden = the unmodified reference's bitstream."""
    w, h = map(int, key.split("_")[0].split("x"))
    n, qp, seed = int(key.split("_n")[1].split("_")[0]), int(key.split("_qp")[1].split("_")[0]), int(key.split("_seed")[1])
    yuv = h264lib.synth_clip(w, h, n, seed=seed)
    outs, _ = gpu_encode(yuv, w, h, n, qp, 30.0)
    assert [len(b) for b in outs[0]] == EDGE["edge"][key]["frame_bytes"]
    assert hashlib.sha1(b"".join(outs[0])).hexdigest() == EDGE["edge"][key]["sha1"]


def test_device_cavlc_bit_count_is_exact():
    """the macroblock kernel's CAVLC bit count (no bits emitted on the device) equals, macroblock by
    macroblock, what the host writer spends — IDR and P pictures, low and high QP, real clip and synthetic content."""
    from openh264_b200.binding import BatchEncoder
    clip = np.fromfile(os.path.join(ROOT, "tests", "golden", "CiscoVT2people_320x192_12fps.yuv"), dtype=np.uint8)
    cases = [(320, 192, 9, 12, clip), (320, 192, 9, 30, clip), (176, 144, 4, 0, h264lib.synth_clip(176, 144, 4, seed=3)),
             (640, 360, 3, 44, h264lib.synth_clip(640, 360, 3, seed=9, noise=8))]
    for w, h, n, qp, yuv in cases:
        enc = BatchEncoder(w, h, qp=qp, fps=30.0, n_streams=2)
        enc.set_mb_bits(True)
        fsz = w * h * 3 // 2
        coded = 0
        for f in range(n):
            bs, _ = enc.encode([yuv[f * fsz:(f + 1) * fsz]] * 2)
            for s in range(2):
                dev, host = enc.mb_bits(s)
                assert np.array_equal(dev, host), (w, h, qp, f, s, np.nonzero(dev != host)[0][:8])
                coded += int((host > 0).sum())
                # the macroblock bits account for the slice payload up to header, skip runs and trailing bits
                assert int(host.sum()) <= 8 * len(bs[s])
        assert coded > 0
        enc.close()


@pytest.mark.parametrize("key", sorted(EDGE["low"]))
def test_low_complexity_through_cuda(key):
    """iComplexityMode = LOW_COMPLEXITY (the reference's default and what SURVEY 8d defines BASELINE configs[1] on): SAD
    mode costs, VAA statistics kernel (8x8 SADs against the previous source picture) driving the partition search, pruned
    I4x4 search; golden = the unmodified reference at LOW_COMPLEXITY."""
    from openh264_b200.binding import BatchEncoder
    g = EDGE["low"][key]
    if key.startswith("clip"):
        w, h, n, qp, fps = 320, 192, 9, int(key.split("qp")[1]), 12.0
        yuv = np.fromfile(os.path.join(ROOT, "tests", "golden", "CiscoVT2people_320x192_12fps.yuv"), dtype=np.uint8)
    else:
        w, h = map(int, key.split("_")[0].split("x"))
        n, qp = int(key.split("_n")[1].split("_")[0]), int(key.split("_qp")[1].split("_")[0])
        yuv = h264lib.synth_clip(w, h, n, seed=int(key.split("_seed")[1].split("_")[0]), noise=int(key.split("_noise")[1]))
        fps = 30.0
    enc = BatchEncoder(w, h, qp=qp, fps=fps, n_streams=2, complexity_low=True)
    fsz = w * h * 3 // 2
    out = [[], []]
    for f in range(n):
        bs, _ = enc.encode([yuv[f * fsz:(f + 1) * fsz]] * 2)
        out[0].append(bs[0]); out[1].append(bs[1])
    enc.close()
    assert [len(b) for b in out[0]] == g["frame_bytes"]
    assert hashlib.sha1(b"".join(out[0])).hexdigest() == g["sha1"] and out[1] == out[0]


@pytest.mark.parametrize("key", sorted(EDGE["cabac"]))
def test_cabac_through_cuda(key):
    """iEntropyCodingModeFlag = 1: the same macroblock kernel, records entropy-coded by the host CABAC writer (csrc/h264_cabac.cpp);
    golden = the unmodified reference with CABAC (High profile by default, Main on request)."""
    from openh264_b200.binding import BatchEncoder
    g = EDGE["cabac"][key]
    prof = int(key.split("_profile")[1].split("_")[0])
    if key.startswith("clip"):
        w, h, n, qp, fps = 320, 192, 9, int(key.split("qp")[1].split("_")[0]), 12.0
        yuv = np.fromfile(os.path.join(ROOT, "tests", "golden", "CiscoVT2people_320x192_12fps.yuv"), dtype=np.uint8)
    else:
        w, h = map(int, key.split("_")[0].split("x"))
        n, qp = int(key.split("_n")[1].split("_")[0]), int(key.split("_qp")[1].split("_")[0])
        yuv = h264lib.synth_clip(w, h, n, seed=int(key.split("_seed")[1].split("_")[0]), noise=int(key.split("_noise")[1].split("_")[0]))
        fps = 30.0
    enc = BatchEncoder(w, h, qp=qp, fps=fps, n_streams=2, complexity_low=key.endswith("_low"), entropy_cabac=True, profile_idc=prof)
    fsz = w * h * 3 // 2
    out = [[], []]
    for f in range(n):
        bs, _ = enc.encode([yuv[f * fsz:(f + 1) * fsz]] * 2)
        out[0].append(bs[0]); out[1].append(bs[1])
    enc.close()
    assert [len(b) for b in out[0]] == g["frame_bytes"]
    assert hashlib.sha1(b"".join(out[0])).hexdigest() == g["sha1"] and out[1] == out[0]


@pytest.mark.parametrize("key", sorted(EDGE["intra_period"]))
def test_intra_period_through_cuda(key):
    """uiIntraPeriod in layer 2 (b2h264_enc_config::intra_period): periodic IDR pictures per stream; golden = the unmodified reference"""
    from openh264_b200.binding import BatchEncoder
    g = EDGE["intra_period"][key]
    w, h = map(int, key.split("_")[0].split("x"))
    f = {k: int(key.split("_" + k)[1].split("_")[0]) for k in ("n", "qp", "seed", "period", "cabac")}
    yuv = h264lib.synth_clip(w, h, f["n"], seed=f["seed"])
    enc = BatchEncoder(w, h, qp=f["qp"], fps=30.0, n_streams=2, entropy_cabac=bool(f["cabac"]), intra_period=f["period"])
    fsz = w * h * 3 // 2
    out = [[], []]
    for i in range(f["n"]):
        frames = [yuv[i * fsz:(i + 1) * fsz], yuv[i * fsz:(i + 1) * fsz]]
        bs, _ = enc.encode(frames)
        out[0].append(bs[0]); out[1].append(bs[1])
    enc.close()
    assert [len(b) for b in out[0]] == g["frame_bytes"]
    assert hashlib.sha1(b"".join(out[0])).hexdigest() == g["sha1"] and out[1] == out[0]


@pytest.mark.parametrize("key", sorted(EDGE["loop_filter"]))
def test_loop_filter_control_through_cuda(key):
    """iLoopFilterDisableIdc / alpha / beta offsets (b2h264_enc_config::loop_filter_*): slice header fields on the host, FilterOffsetA / B
    and the on / off switch in the deblocking kernel; golden = the unmodified reference"""
    from openh264_b200.binding import BatchEncoder
    g = EDGE["loop_filter"][key]
    w, h = map(int, key.split("_")[0].split("x"))
    f = {k: int(key.split("_" + k)[1].split("_")[0]) for k in ("n", "qp", "seed", "idc", "a", "b", "cabac")}
    yuv = h264lib.synth_clip(w, h, f["n"], seed=f["seed"], noise=6)
    enc = BatchEncoder(w, h, qp=f["qp"], fps=30.0, n_streams=2, entropy_cabac=bool(f["cabac"]), loop_filter=(f["idc"], f["a"], f["b"]))
    fsz = w * h * 3 // 2
    out = [[], []]
    for i in range(f["n"]):
        bs, _ = enc.encode([yuv[i * fsz:(i + 1) * fsz]] * 2)
        out[0].append(bs[0]); out[1].append(bs[1])
    enc.close()
    assert [len(b) for b in out[0]] == g["frame_bytes"]
    assert hashlib.sha1(b"".join(out[0])).hexdigest() == g["sha1"] and out[1] == out[0]
