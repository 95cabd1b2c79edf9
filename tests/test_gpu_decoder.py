"""Decoder construct path on the GPU (include/b2h264_codec.h b2h264_dec_*): host parser + one warp per macroblock
(k_decode_mbs) + the encoder path's deblocking / padding kernels must reproduce the reference decoder's pictures
bit for bit on streams produced by the REFERENCE encoder; batches of different streams decode independently."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import h264lib

pytestmark = pytest.mark.gpu
ROOT = h264lib.ROOT


def ref_streams(w, h, n, qp, seeds):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_encoder_golden import ref_encode
    out = []
    for seed in seeds:
        yuv = h264lib.synth_clip(w, h, n, seed=seed)
        bs, fb, _ = ref_encode(yuv, w, h, n, qp, 30.0)
        bs = bytes(bs)
        off = np.concatenate([[0], np.cumsum(fb)])
        out.append((bs, [bs[off[i]:off[i + 1]] for i in range(n)]))
    return out


def ref_decode(bs, w, h, n):
    R = C.CDLL(h264lib.REFSHIM_SO)
    R.ref_decode.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]
    out = np.zeros(n * w * h * 3 // 2 + 64, np.uint8)
    W, H, s = C.c_int(), C.c_int(), C.c_double()
    a = np.frombuffer(bs, np.uint8)
    assert R.ref_decode(a.ctypes.data, len(a), out.ctypes.data, out.size, C.byref(W), C.byref(H), C.byref(s)) == n
    return out


@pytest.mark.parametrize("w,h,n,qp,seeds", [(176, 144, 6, 26, (1, 2, 3)), (640, 360, 4, 32, (4, 5)), (180, 148, 4, 12, (6,))])
def test_gpu_decoder_matches_reference_decoder(w, h, n, qp, seeds):
    if not h264lib.have_ref():
        pytest.skip("oracle/_ref not present")
    from openh264_b200.binding import BatchDecoder
    streams = ref_streams(w, h, n, qp, seeds)
    want = [ref_decode(bs, w, h, n) for bs, _ in streams]
    dec = BatchDecoder(w, h, n_streams=len(seeds))
    fsz = w * h * 3 // 2
    for f in range(n):
        pics = dec.decode([aus[f] for _, aus in streams])
        for s in range(len(seeds)):
            assert np.array_equal(pics[s], want[s][f * fsz:(f + 1) * fsz]), "stream %d picture %d differs" % (s, f)
    dec.close()


def test_gpu_decoder_rejects_unsupported_stream():
    from openh264_b200.binding import BatchDecoder, B2H264Error
    dec = BatchDecoder(176, 144)
    with pytest.raises(B2H264Error):
        dec.decode([b"\x00\x00\x00\x01\x67\x6e\x00\x1f\xac\xd9\x40\x50\x05\xbb\x01\x10"])      # a High 10 profile SPS
    dec.close()


# ---- the reference's own conformance vectors through the GPU decoder ----------------------------------------------------
CONF_DIR = os.path.join(ROOT, "tests", "golden", "conformance")


def _conformance():
    import json
    tab = dict((p.split("/")[-1], s) for p, s in json.load(open(os.path.join(ROOT, "tests", "golden", "reference_decoder_hashes.json")))["pairs"])
    return sorted((f, tab[f]) for f in os.listdir(CONF_DIR) if f in tab)


def _fixture_size(name):
    from openh264_b200.binding import probe_access_unit
    aus = h264lib.split_access_units(open(os.path.join(CONF_DIR, name), "rb").read())
    return probe_access_unit(aus[0])[:2], aus


def test_gpu_decoder_conformance_batch():
    """The reference's decoder test vectors (test/api/decoder_test.cpp:90-142) of QCIF size — BASELINE.json configs[0]'s
    BA_MW_D.264 among them — decoded TOGETHER as one batch of streams by the GPU decoder: different slice structures,
    reference-frame counts (the picture-slot array grows while others are mid-stream), long-term references, several
    parameter sets, partition shapes; every stream must reproduce the PUBLISHED SHA-1 of its pictures."""
    import hashlib
    from openh264_b200.binding import BatchDecoder
    items, streams = [], []
    for f, sha in _conformance():
        size, aus = _fixture_size(f)
        if size == (176, 144):
            items.append((f, sha)); streams.append(aus)
    assert len(items) >= 17
    dec = BatchDecoder(176, 144, n_streams=len(items))
    hashes = [hashlib.sha1() for _ in items]
    for k in range(max(len(a) for a in streams)):
        pics = dec.decode2([a[k] if k < len(a) else None for a in streams])
        for i, p in enumerate(pics):
            assert (p is not None) == (k < len(streams[i])), (items[i][0], k)
            if p is not None:
                hashes[i].update(p.tobytes())
    dec.close()
    bad = [items[i][0] for i in range(len(items)) if hashes[i].hexdigest() != items[i][1]]
    assert not bad, bad


def test_gpu_decoder_conformance_other_sizes():
    import hashlib
    from openh264_b200.binding import BatchDecoder
    ran = 0
    for f, sha in _conformance():
        (w, h), aus = _fixture_size(f)
        if (w, h) == (176, 144):
            continue
        dec = BatchDecoder(w, h)
        hs = hashlib.sha1()
        for au in aus:
            hs.update(dec.decode([au])[0].tobytes())
        dec.close()
        assert hs.hexdigest() == sha, f
        ran += 1
    assert ran >= 2


def test_gpu_decoder_per_stream_status_and_slot_reuse():
    """b2h264_dec_decode3 (what the ISVCDecoder broker calls): one stream of the batch is outside the supported class, one sits
    the call out, one is truncated later on — each gets its own status and the healthy streams decode exactly as alone;
    page-locked output (b2h264_host_alloc); a slot handed to a new stream (b2h264_dec_reset_stream) starts over."""
    if not h264lib.have_ref():
        pytest.skip("oracle/_ref not present")
    from openh264_b200.binding import BatchDecoder
    w, h, n = 176, 144, 5
    streams = ref_streams(w, h, n, 28, (11, 12))
    want = [ref_decode(bs, w, h, n) for bs, _ in streams]
    fsz = w * h * 3 // 2
    high_sps = b"\x00\x00\x00\x01\x67\x6e\x00\x1f\xac\xd9\x40\x50\x05\xbb\x01\x10"
    dec = BatchDecoder(w, h, n_streams=4, pinned_output=True)
    for f in range(n):
        a0 = streams[0][1][f]
        a1 = streams[1][1][f] if f != 2 else streams[1][1][f][:len(streams[1][1][f]) // 2]      # picture 2 of stream 1 arrives truncated
        pics, st = dec.decode3([a0, a1, high_sps if f == 0 else None, None])
        assert st[0] == 1 and np.array_equal(pics[0], want[0][f * fsz:(f + 1) * fsz])
        if f < 2:
            assert st[1] == 1 and np.array_equal(pics[1], want[1][f * fsz:(f + 1) * fsz])
        elif f == 2:
            assert st[1] < 0 and pics[1] is None
        assert st[2] == (-102 if f == 0 else 0) and st[3] == 0
    # slot 1 is handed to a new stream: it starts over with that stream's parameter sets
    dec.reset_stream(1)
    for f in range(n):
        pics, st = dec.decode3([None, streams[0][1][f], None, None])
        assert st == [0, 1, 0, 0] and np.array_equal(pics[1], want[0][f * fsz:(f + 1) * fsz])
    dec.close()


def test_gpu_decoder_cabac_streams_of_the_gpu_encoder():
    """BatchEncoder writes the same pictures with CAVLC and with CABAC (High / Main parameter sets); the GPU decoder (host CABAC
    parser, csrc/h264_cabac_dec.h, in front of the same construct kernels) must turn all three streams into the same pictures —
    and into what the reference decoder makes of the CABAC stream where the reference is on the machine"""
    from openh264_b200.binding import BatchDecoder, BatchEncoder
    w, h, n, qp = 320, 192, 6, 25
    yuv = h264lib.synth_clip(w, h, n, seed=21, noise=8)
    fsz = w * h * 3 // 2
    decoded = []
    for cabac, prof in ((False, 0), (True, 0), (True, 77)):
        enc = BatchEncoder(w, h, qp=qp, fps=30.0, n_streams=1, entropy_cabac=cabac, profile_idc=prof)
        aus = [enc.encode([yuv[f * fsz:(f + 1) * fsz]])[0][0] for f in range(n)]
        enc.close()
        dec = BatchDecoder(w, h)
        pics = np.concatenate([dec.decode([au])[0] for au in aus])
        dec.close()
        if cabac and h264lib.have_ref():
            assert np.array_equal(ref_decode(b"".join(aus), w, h, n)[:pics.size], pics)
        decoded.append(pics)
    assert np.array_equal(decoded[0], decoded[1]) and np.array_equal(decoded[0], decoded[2])
    assert len(decoded[0]) == n * fsz
