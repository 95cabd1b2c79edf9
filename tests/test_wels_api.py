"""Layer 3 (include/b2h264_wels_api.h): the reference's own entry points exported by libopenh264_b200_wels.so.
tests/wels/wels_driver.cpp is an application against the reference's public headers that dlopen()s the library
it is given; the same binary is run with the compiled reference and with our library."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest

import h264lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "wels_driver")
REFLIB = os.path.join(ROOT, "oracle", "_ref", "libopenh264_ref.so")
OURLIB = os.path.join(ROOT, "openh264_b200", "libopenh264_b200_wels.so")
need_built = pytest.mark.skipif(not (os.path.exists(DRIVER) and os.path.exists(OURLIB)),
                                reason="layer-3 shim / driver are built where the reference headers exist (build())")


def drive(lib, clip, w, h, n, qp, idr_at, tmp, tag):
    yuv = os.path.join(tmp, "in.yuv")
    with open(yuv, "wb") as f:
        f.write(clip.tobytes())
    out, lay = os.path.join(tmp, tag + ".264"), os.path.join(tmp, tag + ".layout")
    r = subprocess.run([DRIVER, lib, yuv, str(w), str(h), str(n), str(qp), str(idr_at), out, lay], capture_output=True, text=True,
                       timeout=300)
    return r, (open(out, "rb").read() if os.path.exists(out) else b""), (open(lay).read() if os.path.exists(lay) else "")


@need_built
def test_exports():
    syms = subprocess.run(["nm", "-D", "--defined-only", OURLIB], capture_output=True, text=True).stdout
    for s in ("WelsCreateSVCEncoder", "WelsDestroySVCEncoder", "WelsCreateDecoder", "WelsDestroyDecoder", "WelsGetDecoderCapability",
              "WelsGetCodecVersion", "WelsGetCodecVersionEx"):           # openh264.def
        assert (" T " + s) in syms, s


@need_built
def test_driver_with_reference_matches_golden(tmp_path):
    """pins the driver itself: through the reference it reproduces the bitstream ref_encode() gives (encoder.json source)"""
    w, h, n, qp = 176, 144, 5, 26
    clip = h264lib.synth_clip(w, h, n)
    r, bs, lay = drive(REFLIB, clip, w, h, n, qp, -1, str(tmp_path), "ref")
    assert r.returncode == 0, r.stderr
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_encoder_golden import ref_encode
    ref_bs, _, _ = ref_encode(clip, w, h, n, qp, 30.0)
    assert len(ref_bs) > 0 and hashlib.sha1(bs).hexdigest() == hashlib.sha1(bytes(ref_bs)).hexdigest()
    assert "frame 0 type 1 layers 2" in lay and "frame 1 type 3 layers 1" in lay


@need_built
def test_no_device_fails_loudly(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    clip = h264lib.synth_clip(176, 144, 1)
    r, bs, _ = drive(OURLIB, clip, 176, 144, 1, 26, -1, str(tmp_path), "b2")
    assert r.returncode != 0 and "no CUDA device" in r.stderr and bs == b""


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,n,qp,idr_at", [(176, 144, 6, 26, 3), (320, 192, 5, 32, -1), (640, 368, 4, 22, 2)])
def test_drop_in_same_driver_two_libraries(tmp_path, w, h, n, qp, idr_at):
    assert os.path.exists(DRIVER) and os.path.exists(OURLIB), "prebuilt layer-3 artefacts missing on the GPU box"
    clip = h264lib.synth_clip(w, h, n, seed=7)
    r0, bs0, lay0 = drive(REFLIB, clip, w, h, n, qp, idr_at, str(tmp_path), "ref")
    r1, bs1, lay1 = drive(OURLIB, clip, w, h, n, qp, idr_at, str(tmp_path), "b2")
    assert r0.returncode == 0, r0.stderr
    assert r1.returncode == 0, r1.stderr
    assert bs0 == bs1, "bitstream through ISVCEncoder differs from the reference"
    assert lay0 == lay1, "SFrameBSInfo layout / defaults differ:\n" + lay0 + "\n---\n" + lay1
