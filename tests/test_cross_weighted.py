"""SURVEY 8(a) rows a3 (cross / line search of the screen-content motion estimation) and a15 (weighted / bi-directional decoder
prediction) as layer-1 units:
  * the oracle restatements (orc_me_cross_search, orc_weight_pred, orc_biweight_pred, orc_bi_pred) are pinned call by call
    against the UNMODIFIED reference — WelsMotionCrossSearch through oracle/ref_shim.cpp, the file-local WeightPrediction /
    BiWeightPrediction / BiPrediction through oracle/ref_shim_dec.cpp (compiled with the reference's own rec_mb.cpp) — and
    against golden SHA-1s that travel without the reference build;
  * the CUDA kernels (b2h264_k_me_cross_search, b2h264_k_weighted_pred) must equal the oracle bit for bit."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

import h264lib
from h264lib import MeResult, ptr
from kernel_cases import BLK_DIMS

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "cross_weighted.json")
REFSHIM_DEC = os.path.join(h264lib.ROOT, "oracle", "_ref", "librefshim_dec.so")
u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)


class CrossJob(C.Structure):
    _fields_ = [("blk", C.c_int32), ("cur_off", C.c_int32), ("ref_off", C.c_int32), ("mvp_x", C.c_int16), ("mvp_y", C.c_int16),
                ("mv_min_x", C.c_int16), ("mv_min_y", C.c_int16), ("mv_max_x", C.c_int16), ("mv_max_y", C.c_int16),
                ("qp", C.c_int32), ("sad_cost_threshold", C.c_uint32)]


class WeightJob(C.Structure):
    _fields_ = [("log2_denom", C.c_int32), ("w1", C.c_int32), ("o1", C.c_int32), ("w2", C.c_int32), ("o2", C.c_int32)]


def frames(w, h, pad, seed):
    rng = np.random.RandomState(seed)
    stride = w + 2 * pad
    base = rng.randint(0, 256, size=(h + 2 * pad, stride)).astype(np.uint8)
    # smooth-ish content so that a line search has a meaningful minimum, plus a shifted copy as the reference
    k = np.ones(5) / 5
    sm = np.apply_along_axis(lambda r: np.convolve(r, k, mode="same"), 1, base.astype(np.float32))
    sm = np.apply_along_axis(lambda c: np.convolve(c, k, mode="same"), 0, sm).astype(np.uint8)
    ref = np.roll(np.roll(sm, 3, axis=0), -5, axis=1)
    return np.ascontiguousarray(sm), np.ascontiguousarray(ref), stride


def cross_jobs(rng, w, h, pad, stride, n):
    out = []
    for _ in range(n):
        blk = int(rng.randint(0, 4))
        bw, bh = BLK_DIMS[blk]
        x, y = int(rng.randint(0, w - bw + 1)), int(rng.randint(0, h - bh + 1))
        j = CrossJob()
        j.blk = blk
        j.cur_off = j.ref_off = (pad + y) * stride + pad + x
        j.mvp_x, j.mvp_y = int(rng.randint(-60, 61)), int(rng.randint(-60, 61))
        lim = pad - 4
        j.mv_min_x, j.mv_max_x = -int(rng.randint(1, lim)), int(rng.randint(1, lim))
        j.mv_min_y, j.mv_max_y = -int(rng.randint(1, lim)), int(rng.randint(1, lim))
        j.qp = int(rng.choice([12, 26, 40]))
        r = MeResult()
        r.mv_x, r.mv_y = int(rng.randint(-3, 4)), int(rng.randint(-3, 4))      # what the diamond search left behind
        r.ref_off = j.ref_off + r.mv_y * stride + r.mv_x
        r.sad_cost = int(rng.choice([50, 2000, 20000, 0x7fffffff]))
        j.sad_cost_threshold = int(rng.choice([0, 1000, 30000]))
        out.append((j, r))
    return out


def run_cross(fn, cur, ref, stride, jobs):
    res = []
    for j, r0 in jobs:
        r = MeResult()
        C.memmove(C.byref(r), C.byref(r0), C.sizeof(MeResult))
        fn(ptr(cur), stride, ptr(ref), stride, C.byref(j), C.byref(r))
        res.append((r.mv_x, r.mv_y, r.sad_cost, r.ref_off))
    return res


def cross_fn(lib, name):
    fn = getattr(lib, name)
    fn.restype = None
    fn.argtypes = [u8p, C.c_int, u8p, C.c_int, C.POINTER(CrossJob), C.POINTER(MeResult)]
    return fn


def cross_case():
    w, h, pad = 128, 96, 40
    cur, ref, stride = frames(w, h, pad, 77)
    jobs = cross_jobs(np.random.RandomState(78), w, h, pad, stride, 400)
    return cur, ref, stride, jobs


def weight_cases(rng, n):
    cases = []
    for _ in range(n):
        w, h = [(16, 16), (16, 8), (8, 16), (8, 8), (4, 4), (8, 4)][int(rng.randint(0, 6))]
        sy, sc = 32, 16
        planes = [rng.randint(0, 256, size=(16 * s,)).astype(np.uint8) for s in (sy, sc, sc, sy, sc, sc)]
        l2 = [int(rng.randint(0, 8)), int(rng.randint(0, 8))]
        w1 = [int(rng.randint(-128, 128)) for _ in range(3)]
        o1 = [int(rng.randint(-128, 128)) for _ in range(3)]
        w2 = [int(rng.randint(-128, 128)) for _ in range(3)]
        o2 = [int(rng.randint(-128, 128)) for _ in range(3)]
        cases.append((w, h, sy, sc, planes, l2, w1, o1, w2, o2))
    return cases


def oracle_weighted(orc, mode, case, implicit=False):
    """mode 0 weight, 1 bi-weight, 2 average; returns the three planes after the in-place operation"""
    w, h, sy, sc, planes, l2, w1, o1, w2, o2 = case
    out = [p.copy() for p in planes[:3]]
    for pl in range(3):
        pw, ph, st = (w, h, sy) if pl == 0 else (w // 2, h // 2, sc)
        ld = l2[0] if pl == 0 else l2[1]
        if mode == 0:
            orc.orc_weight_pred(ptr(out[pl]), st, pw, ph, ld, w1[pl], o1[pl])
        elif mode == 1:
            if implicit:
                orc.orc_biweight_pred(ptr(out[pl]), ptr(planes[3 + pl]), st, pw, ph, 5, w1[0] & 63, 0, 64 - (w1[0] & 63), 0)
            else:
                orc.orc_biweight_pred(ptr(out[pl]), ptr(planes[3 + pl]), st, pw, ph, ld, w1[pl], o1[pl], w2[pl], o2[pl])
        else:
            orc.orc_bi_pred(ptr(out[pl]), ptr(planes[3 + pl]), st, pw, ph)
    return out


def ref_weighted(R, mode, case, implicit=False):
    w, h, sy, sc, planes, l2, w1, o1, w2, o2 = case
    out = [p.copy() for p in planes[:3]]
    A = lambda v: (C.c_int32 * 3)(*v)
    if mode == 0:
        R.ref_weight_pred(ptr(out[0]), ptr(out[1]), ptr(out[2]), sy, sc, w, h, l2[0], l2[1], A(w1), A(o1))
    elif mode == 1:
        if implicit:
            R.ref_biweight_pred(ptr(out[0]), ptr(out[1]), ptr(out[2]), ptr(planes[3]), ptr(planes[4]), ptr(planes[5]), sy, sc, w, h, 0, 5, 5,
                                A([w1[0] & 63] * 3), A([0] * 3), A([0] * 3), A([0] * 3))
        else:
            R.ref_biweight_pred(ptr(out[0]), ptr(out[1]), ptr(out[2]), ptr(planes[3]), ptr(planes[4]), ptr(planes[5]), sy, sc, w, h, 1, l2[0], l2[1],
                                A(w1), A(o1), A(w2), A(o2))
    else:
        R.ref_bi_pred(ptr(out[0]), ptr(out[1]), ptr(out[2]), ptr(planes[3]), ptr(planes[4]), ptr(planes[5]), sy, sc, w, h)
    return out


def orc_lib():
    lib = h264lib.oracle().lib
    lib.orc_weight_pred.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.orc_biweight_pred.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.orc_bi_pred.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int]
    for f in (lib.orc_weight_pred, lib.orc_biweight_pred, lib.orc_bi_pred):
        f.restype = None
    return lib


def digest(obj):
    return hashlib.sha1(json.dumps(obj, sort_keys=True).encode()).hexdigest()


def oracle_digests():
    cur, ref, stride, jobs = cross_case()
    d = {"cross_search": digest(run_cross(cross_fn(h264lib.oracle().lib, "orc_me_cross_search"), cur, ref, stride, jobs))}
    orc = orc_lib()
    cases = weight_cases(np.random.RandomState(5), 60)
    for name, mode, imp in (("weight", 0, False), ("biweight_explicit", 1, False), ("biweight_implicit", 1, True), ("bi_pred", 2, False)):
        d[name] = digest([[p.tolist() for p in oracle_weighted(orc, mode, c, imp)] for c in cases])
    return d


def test_oracle_cross_search_matches_reference():
    if not h264lib.have_ref():
        pytest.skip("oracle/_ref not built on this machine")
    cur, ref, stride, jobs = cross_case()
    a = run_cross(cross_fn(h264lib.oracle().lib, "orc_me_cross_search"), cur, ref, stride, jobs)
    b = run_cross(cross_fn(h264lib.ref().lib, "ref_me_cross_search"), cur, ref, stride, jobs)
    assert a == b
    assert len({x[:2] for x in a}) > 20                     # the cases do move the vector around


@pytest.mark.parametrize("name,mode,implicit", [("weight", 0, False), ("biweight_explicit", 1, False), ("biweight_implicit", 1, True), ("bi_pred", 2, False)])
def test_oracle_weighted_prediction_matches_reference(name, mode, implicit):
    if not os.path.exists(REFSHIM_DEC):
        pytest.skip("oracle/_ref not built on this machine")
    R = C.CDLL(REFSHIM_DEC)
    R.ref_weight_pred.argtypes = [u8p, u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p, i32p]
    R.ref_biweight_pred.argtypes = [u8p] * 6 + [C.c_int] * 7 + [i32p] * 4
    R.ref_bi_pred.argtypes = [u8p] * 6 + [C.c_int] * 4
    orc = orc_lib()
    for case in weight_cases(np.random.RandomState(5), 60):
        a = oracle_weighted(orc, mode, case, implicit)
        b = ref_weighted(R, mode, case, implicit)
        for pl in range(3):
            assert np.array_equal(a[pl], b[pl]), (name, pl, case[:2])


def test_oracle_matches_golden():
    assert oracle_digests() == json.load(open(GOLDEN))


@pytest.mark.gpu
def test_gpu_cross_search_matches_oracle():
    import torch
    from openh264_b200.binding import lib, check
    L = lib(0)
    w, h, pad = 320, 192, 40
    cur, ref, stride = frames(w, h, pad, 91)
    jobs = cross_jobs(np.random.RandomState(92), w, h, pad, stride, 3000)
    exp = run_cross(cross_fn(h264lib.oracle().lib, "orc_me_cross_search"), cur, ref, stride, jobs)
    n = len(jobs)
    jarr = (CrossJob * n)(*[j for j, _ in jobs])
    rarr = (MeResult * n)(*[r for _, r in jobs])
    dev = lambda a: torch.from_numpy(np.frombuffer(bytes(a), np.uint8).copy()).cuda()
    dcur, dref = torch.from_numpy(cur).cuda(), torch.from_numpy(ref).cuda()
    dj, dr = dev(jarr), dev(rarr)
    check(L.b2h264_k_me_cross_search(dcur.data_ptr(), stride, dref.data_ptr(), stride, dj.data_ptr(), n, dr.data_ptr(), None))
    torch.cuda.synchronize()
    res = (MeResult * n).from_buffer_copy(dr.cpu().numpy().tobytes())
    assert [(r.mv_x, r.mv_y, r.sad_cost, r.ref_off) for r in res] == exp


@pytest.mark.gpu
@pytest.mark.parametrize("mode,implicit", [(0, False), (1, False), (1, True), (2, False)])
def test_gpu_weighted_prediction_matches_oracle(mode, implicit):
    import torch
    from openh264_b200.binding import lib, check
    L = lib(0)
    orc = orc_lib()
    rng = np.random.RandomState(17 + mode)
    for (w, h) in ((16, 16), (8, 8), (16, 8), (4, 4), (2, 2)):
        n, stride = 500, 24
        dst = rng.randint(0, 256, size=(n * 16 * stride,)).astype(np.uint8)
        tmp = rng.randint(0, 256, size=(n * 16 * stride,)).astype(np.uint8)
        offs = (np.arange(n) * 16 * stride + rng.randint(0, stride - w + 1, size=n)).astype(np.int32)
        jobs = (WeightJob * n)()
        exp = dst.copy()
        for i in range(n):
            jb = jobs[i]
            if implicit:
                jb.log2_denom, jb.w1, jb.o1, jb.o2 = 5, int(rng.randint(0, 64)), 0, 0
                jb.w2 = 64 - jb.w1
            else:
                jb.log2_denom = int(rng.randint(0, 8))
                jb.w1, jb.o1, jb.w2, jb.o2 = [int(rng.randint(-128, 128)) for _ in range(4)]
            o = int(offs[i])
            if mode == 0:
                orc.orc_weight_pred(ptr(exp, off=o), stride, w, h, jb.log2_denom, jb.w1, jb.o1)
            elif mode == 1:
                orc.orc_biweight_pred(ptr(exp, off=o), ptr(tmp, off=o), stride, w, h, jb.log2_denom, jb.w1, jb.o1, jb.w2, jb.o2)
            else:
                orc.orc_bi_pred(ptr(exp, off=o), ptr(tmp, off=o), stride, w, h)
        ddst, dtmp, doff = torch.from_numpy(dst.copy()).cuda(), torch.from_numpy(tmp).cuda(), torch.from_numpy(offs).cuda()
        dj = torch.from_numpy(np.frombuffer(bytes(jobs), np.uint8).copy()).cuda()
        check(L.b2h264_k_weighted_pred(mode, ddst.data_ptr(), dtmp.data_ptr(), stride, doff.data_ptr(), doff.data_ptr(), dj.data_ptr(), w, h, n, None))
        torch.cuda.synchronize()
        assert np.array_equal(ddst.cpu().numpy(), exp), (mode, implicit, w, h)
