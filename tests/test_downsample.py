"""VPP bilinear down-sampler (codec/processing/src/downsample/downsamplefuncs.cpp).
CPU: the oracle restatement against the compiled reference's own `_c` functions (C++ symbols of oracle/_ref).
GPU: b2h264_k_downsample against the oracle, incl. the simulcast ladder of BASELINE.json configs[4]
(1280x720 -> 640x360 / 320x180 / 160x90, test/api/BaseEncoderTest.cpp:43-44) and batches of planes."""
import ctypes as C
import os

import numpy as np
import pytest

import h264lib
from h264lib import ptr

REF_SYMS = {0: "_ZN6WelsVP27DyadicBilinearDownsampler_cEPhiS0_iii", 1: "_ZN6WelsVP34DyadicBilinearQuarterDownsampler_cEPhiS0_iii",
            2: "_ZN6WelsVP35DyadicBilinearOneThirdDownsampler_cEPhiS0_iii", 3: "_ZN6WelsVP32GeneralBilinearFastDownsampler_cEPhiiiS0_iii",
            4: "_ZN6WelsVP36GeneralBilinearAccurateDownsampler_cEPhiiiS0_iii"}
# (mode, src_w, src_h, dst_w, dst_h)
CASES = [(0, 1280, 720, 640, 360), (0, 640, 360, 320, 180), (0, 34, 18, 17, 9), (1, 1280, 720, 320, 180), (1, 64, 36, 16, 9),
         (2, 1920, 1080, 640, 360), (2, 96, 51, 32, 17), (3, 1280, 720, 160, 90), (3, 1280, 720, 854, 480), (3, 333, 217, 100, 77),
         (4, 640, 360, 80, 45), (4, 640, 360, 427, 240), (4, 167, 109, 50, 39)]


def orc_downsample(mode, src, dw, dh):
    O = h264lib.oracle().lib
    O.orc_downsample.argtypes = [C.c_int, h264lib.u8p, C.c_int, C.c_int, C.c_int, h264lib.u8p, C.c_int, C.c_int, C.c_int]
    O.orc_downsample.restype = None
    dst = np.zeros((dh, dw + 5), np.uint8)
    O.orc_downsample(mode, ptr(dst), dst.shape[1], dw, dh, ptr(src), src.shape[1], src.shape[1] - 3, src.shape[0] - 1)
    return dst[:, :dw].copy()


def make_src(rng, sw, sh):
    return rng.randint(0, 256, size=(sh + 1, sw + 3)).astype(np.uint8)        # stride > width, one spare row (2x2 reads at the edge)


@pytest.mark.parametrize("case", CASES)
def test_oracle_downsample_matches_reference(case):
    if not h264lib.have_ref():
        pytest.skip("oracle/_ref not present")
    mode, sw, sh, dw, dh = case
    R = C.CDLL(os.path.join(h264lib.REF_DIR, "libopenh264_ref.so"))
    fn = getattr(R, REF_SYMS[mode])
    fn.restype = None
    src = make_src(np.random.RandomState(7 + mode), sw, sh)
    dst = np.zeros((dh, dw + 5), np.uint8)
    if mode == 2:     # (dst, dst_stride, src, src_stride, src_width, DST height)
        fn(ptr(dst), C.c_int(dst.shape[1]), ptr(src), C.c_int(src.shape[1]), C.c_int(sw), C.c_int(dh))
    elif mode < 2:    # (dst, dst_stride, src, src_stride, src_width, src_height)
        fn(ptr(dst), C.c_int(dst.shape[1]), ptr(src), C.c_int(src.shape[1]), C.c_int(sw), C.c_int(sh))
    else:             # (dst, dst_stride, dst_w, dst_h, src, src_stride, src_w, src_h)
        fn(ptr(dst), C.c_int(dst.shape[1]), C.c_int(dw), C.c_int(dh), ptr(src), C.c_int(src.shape[1]), C.c_int(sw), C.c_int(sh))
    assert np.array_equal(dst[:, :dw], orc_downsample(mode, src, dw, dh))


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_downsample_matches_oracle(case):
    import openh264_b200 as m
    from openh264_b200.binding import check
    mode, sw, sh, dw, dh = case
    L = m.lib(0)
    rng = np.random.RandomState(70 + mode)
    n = 3                                                         # a batch of planes (the same plane of 3 streams)
    srcs = [make_src(rng, sw, sh) for _ in range(n)]
    stack = np.stack(srcs)
    dsrc = m.DeviceArray(stack)
    dstride = (dw + 7) & ~3
    ddst = m.DeviceArray(np.zeros((n, dh, dstride), np.uint8))
    check(L.b2h264_k_downsample(mode, ddst.ptr, dstride, dw, dh, dsrc.ptr, stack.shape[2], sw, sh, n, dh * dstride, stack.shape[1] * stack.shape[2], None))
    got = ddst.get()
    for i in range(n):
        assert np.array_equal(got[i, :, :dw], orc_downsample(mode, srcs[i], dw, dh)), "plane %d" % i
    assert not got[:, :, dw:].any()                               # nothing written beyond the plane's width


def test_downsample_mode_dispatch():
    """the function CDownsampling::Process picks (downsample.cpp:159-215)"""
    from openh264_b200.binding import load
    L = load()
    assert L.b2h264_downsample_mode(1280, 720, 640, 360, 0) == 0
    assert L.b2h264_downsample_mode(1280, 720, 320, 180, 0) == 1
    assert L.b2h264_downsample_mode(1920, 1080, 640, 360, 1) == 2
    assert L.b2h264_downsample_mode(1280, 720, 160, 90, 0) == 3
    assert L.b2h264_downsample_mode(640, 360, 80, 45, 1) == 4
