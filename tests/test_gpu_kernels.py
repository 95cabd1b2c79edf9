"""GPU parity tests, layer 1: every batched kernel of include/b2h264.h, called through the C-ABI,
against the CPU oracle (oracle/h264_oracle.c) on the same seeded inputs.  Bit-exact (integer path).
Method follows the reference's test/encoder/EncUT_*.cpp / test/decoder/DecUT_*.cpp.
"""
import ctypes as C

import numpy as np
import pytest

import h264lib
from h264lib import BLK_DIMS, ptr
from kernel_cases import make_me_jobs

pytestmark = pytest.mark.gpu

H, STRIDE = 96, 128


@pytest.fixture(scope="module")
def env():
    import openh264_b200 as m
    from openh264_b200 import binding as libmod
    L = m.lib(0)
    return m, libmod, L, h264lib.oracle()


def planes(rng, n=2):
    return [rng.randint(0, 256, size=(H, STRIDE)).astype(np.uint8) for _ in range(n)]


def offsets(rng, n, margin=24):
    y = rng.randint(margin, H - margin - 17, size=n)
    x = rng.randint(margin, STRIDE - margin - 17, size=n)
    return (y * STRIDE + x).astype(np.int32)


@pytest.mark.parametrize("blk", range(7))
def test_sad_satd_sadfour(env, blk):
    m, lm, L, orc = env
    rng = np.random.RandomState(100 + blk)
    a, b = planes(rng)
    n = 600
    oa, ob = offsets(rng, n), offsets(rng, n)
    da, db, doa, dob = m.DeviceArray(a), m.DeviceArray(b), m.DeviceArray(oa), m.DeviceArray(ob)
    sad, satd, sad4 = (m.DeviceArray(shape=(n,), dtype=np.int32), m.DeviceArray(shape=(n,), dtype=np.int32),
                       m.DeviceArray(shape=(n, 4), dtype=np.int32))
    lm.check(L.b2h264_k_sad(da.ptr, STRIDE, doa.ptr, db.ptr, STRIDE, dob.ptr, blk, n, sad.ptr, satd.ptr, sad4.ptr, None))
    g_sad, g_satd, g_sad4 = sad.get(), satd.get(), sad4.get()
    for j in range(n):
        pa, pb = ptr(a, off=int(oa[j])), ptr(b, off=int(ob[j]))
        assert g_sad[j] == orc.sad(blk, pa, STRIDE, pb, STRIDE)
        assert g_satd[j] == orc.satd(blk, pa, STRIDE, pb, STRIDE)
        four = np.zeros(4, np.int32)
        orc.sad_four(blk, pa, STRIDE, pb, STRIDE, ptr(four))
        assert np.array_equal(g_sad4[j], four)


def test_sad_empty_batch(env):
    m, lm, L, orc = env
    assert L.b2h264_k_sad(None, 0, None, None, 0, None, 0, 0, None, None, None, None) == 0


@pytest.mark.parametrize("wh", [(16, 16), (16, 8), (8, 16), (8, 8), (8, 4), (4, 8), (4, 4)])
def test_mc_luma_all_positions(env, wh):
    m, lm, L, orc = env
    w, h = wh
    rng = np.random.RandomState(7 + w * 17 + h)
    src = planes(rng, 1)[0]
    src[:, 40:60] = ((np.indices((H, 20)).sum(0) & 1) * 255).astype(np.uint8)   # clip-stressing stripe
    n = 16 * 24
    off = offsets(rng, n)
    mv = np.array([[i % 4 + 4 * rng.randint(-3, 4), (i // 4) % 4 + 4 * rng.randint(-3, 4)] for i in range(n)], np.int16)
    dsrc, doff, dmv = m.DeviceArray(src), m.DeviceArray(off), m.DeviceArray(mv)
    dst = m.DeviceArray(np.zeros((n, 16, 16), np.uint8))
    lm.check(L.b2h264_k_mc_luma(dsrc.ptr, STRIDE, doff.ptr, dmv.ptr, w, h, n, dst.ptr, None))
    got = dst.get()
    for j in range(n):
        exp = np.zeros((16, 16), np.uint8)
        orc.mc_luma(ptr(src, off=int(off[j])), STRIDE, ptr(exp), 16, int(mv[j, 0]), int(mv[j, 1]), w, h)
        assert np.array_equal(got[j], exp), (j, mv[j])


@pytest.mark.parametrize("wh", [(8, 8), (8, 4), (4, 8), (4, 4), (4, 2), (2, 4), (2, 2)])
def test_mc_chroma(env, wh):
    m, lm, L, orc = env
    w, h = wh
    rng = np.random.RandomState(11 + w * 17 + h)
    src = planes(rng, 1)[0]
    n = 64 * 3
    off = offsets(rng, n)
    mv = np.array([[i % 8 + 8 * rng.randint(-2, 3), (i // 8) % 8 + 8 * rng.randint(-2, 3)] for i in range(n)], np.int16)
    dsrc, doff, dmv = m.DeviceArray(src), m.DeviceArray(off), m.DeviceArray(mv)
    dst = m.DeviceArray(np.zeros((n, 8, 8), np.uint8))
    lm.check(L.b2h264_k_mc_chroma(dsrc.ptr, STRIDE, doff.ptr, dmv.ptr, w, h, n, dst.ptr, None))
    got = dst.get()
    for j in range(n):
        exp = np.zeros((8, 8), np.uint8)
        orc.mc_chroma(ptr(src, off=int(off[j])), STRIDE, ptr(exp), 8, int(mv[j, 0]), int(mv[j, 1]), w, h)
        assert np.array_equal(got[j], exp)


def test_halfpel_planes_and_avg(env):
    m, lm, L, orc = env
    rng = np.random.RandomState(5)
    a, b = planes(rng)
    n = 200
    oa, ob = offsets(rng, n), offsets(rng, n)
    da, db, doa, dob = m.DeviceArray(a), m.DeviceArray(b), m.DeviceArray(oa), m.DeviceArray(ob)
    # pfLumaHalfpelHor (W+1 x H), Ver (W x H+1), Cen (W+1 x H+1) == McLuma at (2,0)/(0,2)/(2,2) with those sizes
    for which, (w, h), (fx, fy) in [(0, (17, 16), (2, 0)), (1, (16, 17), (0, 2)), (2, (17, 17), (2, 2)), (2, (9, 9), (2, 2))]:
        dst = m.DeviceArray(np.zeros((n, 17, 17), np.uint8))
        lm.check(L.b2h264_k_halfpel(which, da.ptr, STRIDE, doa.ptr, w, h, n, dst.ptr, None))
        got = dst.get()
        for j in range(n):
            exp = np.zeros((17, 17), np.uint8)
            orc.mc_luma(ptr(a, off=int(oa[j])), STRIDE, ptr(exp), 17, fx, fy, w, h)
            assert np.array_equal(got[j], exp)
    dst = m.DeviceArray(np.zeros((n, 16, 16), np.uint8))
    lm.check(L.b2h264_k_pixel_avg(da.ptr, STRIDE, doa.ptr, db.ptr, STRIDE, dob.ptr, 16, 16, n, dst.ptr, None))
    got = dst.get()
    for j in range(n):
        exp = np.zeros((16, 16), np.uint8)
        orc.pixel_avg(ptr(exp), 16, ptr(a, off=int(oa[j])), STRIDE, ptr(b, off=int(ob[j])), STRIDE, 16, 16)
        assert np.array_equal(got[j], exp)


@pytest.mark.parametrize("qp", [0, 11, 12, 26, 37, 51])
def test_dct_quant_dequant_idct_chain(env, qp):
    m, lm, L, orc = env
    rng = np.random.RandomState(200 + qp)
    a, b = planes(rng)
    n = 500
    oa, ob = offsets(rng, n), offsets(rng, n)
    da, db, doa, dob = m.DeviceArray(a), m.DeviceArray(b), m.DeviceArray(oa), m.DeviceArray(ob)
    dct = m.DeviceArray(np.zeros((n, 64), np.int16))
    lm.check(L.b2h264_k_dct_four4x4(da.ptr, STRIDE, doa.ptr, db.ptr, STRIDE, dob.ptr, n, dct.ptr, None))
    g_dct = dct.get()
    e_dct = np.zeros((n, 64), np.int16)
    for j in range(n):
        orc.dct_four4x4(ptr(e_dct[j]), ptr(a, off=int(oa[j])), STRIDE, ptr(b, off=int(ob[j])), STRIDE)
    assert np.array_equal(g_dct, e_dct)
    for intra in (0, 1):
        q = m.DeviceArray(e_dct)
        mx = m.DeviceArray(np.zeros((n, 4), np.int16))
        lm.check(L.b2h264_k_quant_four4x4(q.ptr, qp, intra, n, mx.ptr, None))
        ff = np.ctypeslib.as_array(orc.quant_ff(qp + 6 * intra), shape=(8,)).copy()
        mf = np.ctypeslib.as_array(orc.quant_mf(qp), shape=(8,)).copy()
        e_q, e_mx = e_dct.copy(), np.zeros((n, 4), np.int16)
        for j in range(n):
            orc.quant_four4x4_max(ptr(e_q[j]), ptr(ff), ptr(mf), ptr(e_mx[j]))
        assert np.array_equal(q.get(), e_q) and np.array_equal(mx.get(), e_mx)
        # scan / score / nzc of every 4x4
        lv_in = m.DeviceArray(e_q.reshape(-1, 16))
        dcac, ac = m.DeviceArray(np.zeros((4 * n, 16), np.int16)), m.DeviceArray(np.zeros((4 * n, 16), np.int16))
        cn = m.DeviceArray(np.zeros((4 * n, 2), np.int32))
        lm.check(L.b2h264_k_scan4x4(lv_in.ptr, 4 * n, dcac.ptr, ac.ptr, cn.ptr, None))
        g_dcac, g_ac, g_cn = dcac.get(), ac.get(), cn.get()
        blocks = e_q.reshape(-1, 16)
        for t in range(0, 4 * n, 7):
            l1, l2 = np.zeros(16, np.int16), np.zeros(16, np.int16)
            orc.scan4x4_dcac(ptr(l1), ptr(blocks[t]))
            orc.scan4x4_ac(ptr(l2), ptr(blocks[t]))
            assert np.array_equal(g_dcac[t], l1) and np.array_equal(g_ac[t], l2)
            assert g_cn[t, 0] == orc.single_ctr4x4(ptr(l1)) and g_cn[t, 1] == orc.nonzero_count(ptr(l1))
        # dequant + idct + reconstruction
        dq = m.DeviceArray(e_q)
        lm.check(L.b2h264_k_dequant_four4x4(dq.ptr, qp, n, None))
        dqc = np.ctypeslib.as_array(orc.dequant_coeff(qp), shape=(8,)).copy()
        e_dq = e_q.copy()
        for j in range(n):
            orc.dequant_four4x4(ptr(e_dq[j]), ptr(dqc))
        assert np.array_equal(dq.get(), e_dq)
        rec = m.DeviceArray(np.zeros((n, 8, 8), np.uint8))
        lm.check(L.b2h264_k_idct_four4x4_rec(db.ptr, STRIDE, dob.ptr, dq.ptr, n, rec.ptr, None))
        g_rec = rec.get()
        for j in range(n):
            exp = np.zeros((8, 8), np.uint8)
            orc.idct_four4x4_rec(ptr(exp), 8, ptr(b, off=int(ob[j])), STRIDE, ptr(e_dq[j]))
            assert np.array_equal(g_rec[j], exp)


def test_dc_paths(env):
    m, lm, L, orc = env
    rng = np.random.RandomState(31)
    n = 300
    mb = rng.randint(-4096, 4096, size=(n, 256)).astype(np.int16)
    dc = m.DeviceArray(np.zeros((n, 16), np.int16))
    dmb = m.DeviceArray(mb)
    lm.check(L.b2h264_k_hadamard_t4_dc(dmb.ptr, n, dc.ptr, None))
    e = np.zeros((n, 16), np.int16)
    for j in range(n):
        orc.hadamard_t4_dc(ptr(e[j]), ptr(mb[j]))
    assert np.array_equal(dc.get(), e)
    for qp in (4, 26, 45):
        ff = int(np.ctypeslib.as_array(orc.quant_ff(qp), shape=(8,))[0]) << 1
        mf = int(np.ctypeslib.as_array(orc.quant_mf(qp), shape=(8,))[0]) >> 1
        # luma DC quant
        q = m.DeviceArray(e)
        lm.check(L.b2h264_k_quant4x4_dc(q.ptr, ff, mf, n, None))
        eq = e.copy()
        for j in range(n):
            orc.quant4x4_dc(ptr(eq[j]), ff, mf)
        assert np.array_equal(q.get(), eq)
        # chroma DC 2x2
        ch = rng.randint(-2048, 2048, size=(n, 64)).astype(np.int16)
        ch[::5, [0, 16, 32, 48]] = rng.randint(-3, 4, size=(len(ch[::5]), 4))
        dch = m.DeviceArray(ch)
        d4, nz, sk = (m.DeviceArray(np.zeros((n, 4), np.int16)), m.DeviceArray(np.zeros(n, np.int32)),
                      m.DeviceArray(np.zeros(n, np.int32)))
        lm.check(L.b2h264_k_hadamard_quant2x2(dch.ptr, ff, mf, n, d4.ptr, nz.ptr, sk.ptr, None))
        g_ch, g_d4, g_nz, g_sk = dch.get(), d4.get(), nz.get(), sk.get()
        for j in range(n):
            c2 = ch[j].copy()
            assert g_sk[j] == orc.hadamard_quant2x2_skip(ptr(c2), ff, mf)
            dd, bb = np.zeros(4, np.int16), np.zeros(4, np.int16)
            assert g_nz[j] == orc.hadamard_quant2x2(ptr(c2), ff, mf, ptr(dd), ptr(bb))
            assert np.array_equal(g_d4[j], dd) and np.array_equal(g_ch[j], c2)
    lv = rng.randint(-300, 300, size=(n, 16)).astype(np.int16)
    for qp in (0, 7, 11):
        r = m.DeviceArray(lv)
        lm.check(L.b2h264_k_dequant_luma_dc(r.ptr, qp, n, None))
        ex = lv.copy()
        for j in range(n):
            orc.ihadamard4x4_dc(ptr(ex[j]))
            orc.dequant_luma_dc4x4(ptr(ex[j]), qp)
        assert np.array_equal(r.get(), ex)
    wide = rng.randint(-32768, 32768, size=(n, 16)).astype(np.int16)
    for mf in (16, 104, 1024):
        for src in (lv, wide):
            r = m.DeviceArray(src)
            lm.check(L.b2h264_k_dequant_ihadamard4x4(r.ptr, mf, n, None))
            ex = src.copy()
            for j in range(n):
                orc.dequant_ihadamard4x4(ptr(ex[j]), mf)
            assert np.array_equal(r.get(), ex)
            r2 = m.DeviceArray(src[:, :4].copy())
            lm.check(L.b2h264_k_dequant_ihadamard2x2(r2.ptr, mf, n, None))
            ex2 = src[:, :4].copy()
            for j in range(n):
                orc.dequant_ihadamard2x2_dc(ptr(ex2[j]), mf)
            assert np.array_equal(r2.get(), ex2)
    pred = planes(rng, 1)[0]
    off = offsets(rng, n)
    dpred, doff, ddc = m.DeviceArray(pred), m.DeviceArray(off), m.DeviceArray(lv)
    rec = m.DeviceArray(np.zeros((n, 16, 16), np.uint8))
    lm.check(L.b2h264_k_idct_rec_i16x16_dc(dpred.ptr, STRIDE, doff.ptr, ddc.ptr, n, rec.ptr, None))
    g = rec.get()
    for j in range(n):
        exp = np.zeros((16, 16), np.uint8)
        orc.idct_rec_i16x16_dc(ptr(exp), 16, ptr(pred, off=int(off[j])), STRIDE, ptr(lv[j]))
        assert np.array_equal(g[j], exp)


@pytest.mark.parametrize("size", [4, 8])
def test_idct_res_add_pred(env, size):
    m, lm, L, orc = env
    rng = np.random.RandomState(41 + size)
    pic = planes(rng, 1)[0]
    # non-overlapping blocks on a grid
    pos = [(y, x) for y in range(8, H - 16, 8) for x in range(8, STRIDE - 16, 8)]
    n = len(pos)
    off = np.array([y * STRIDE + x for (y, x) in pos], np.int32)
    ne = size * size
    rs = rng.randint(-2000, 2001, size=(n, ne)).astype(np.int16)
    rs[::4] = rng.randint(-32768, 32768, size=rs[::4].shape)       # int16 wrap-around cases
    rs[1::9] = 0                                                     # all-zero residual
    dpic, doff, drs = m.DeviceArray(pic), m.DeviceArray(off), m.DeviceArray(rs)
    lm.check(L.b2h264_k_idct_res_add_pred(dpic.ptr, STRIDE, doff.ptr, drs.ptr, size, n, None))
    exp = pic.copy()
    f = orc.idct_res_add_pred if size == 4 else orc.idct_res_add_pred8x8
    for j in range(n):
        f(ptr(exp, off=int(off[j])), STRIDE, ptr(rs[j]))
    assert np.array_equal(dpic.get(), exp)


def test_deblock_filters(env):
    from openh264_b200.binding import EdgeJob
    m, lm, L, orc = env
    rng = np.random.RandomState(77)
    base = rng.randint(40, 200, size=(H // 16, STRIDE // 16)).repeat(16, 0).repeat(16, 1)
    pic = np.clip(base + rng.randint(-9, 10, size=(H, STRIDE)), 0, 255).astype(np.uint8)
    pic2 = np.clip(base + rng.randint(-5, 6, size=(H, STRIDE)), 0, 255).astype(np.uint8)
    for vertical_edge in (True, False):
        jobs_np = []
        # disjoint edges: one per 16x16 cell
        for cy in range(1, H // 16 - 1):
            for cx in range(1, STRIDE // 16 - 1):
                e = EdgeJob()
                e.off = (cy * 16) * STRIDE + cx * 16 + (8 if vertical_edge else 8 * STRIDE)
                e.sx, e.sy = (1, STRIDE) if vertical_edge else (STRIDE, 1)
                e.alpha, e.beta = int(rng.randint(4, 90)), int(rng.randint(2, 19))
                for k in range(4):
                    e.tc[k] = int(rng.randint(-1, 7))
                e.strong = int(rng.rand() < 0.4)
                jobs_np.append(e)
        n = len(jobs_np)
        arr = (EdgeJob * n)(*jobs_np)
        djobs = m.DeviceArray(np.frombuffer(bytes(arr), dtype=np.uint8))
        dl = m.DeviceArray(pic)
        lm.check(L.b2h264_k_deblock_luma(dl.ptr, djobs.ptr, n, None))
        exp = pic.copy()
        for e in jobs_np:
            tc = np.array(list(e.tc), np.int8)
            if e.strong:
                orc.deblock_luma_eq4(ptr(exp, off=e.off), e.sx, e.sy, e.alpha, e.beta)
            else:
                orc.deblock_luma_lt4(ptr(exp, off=e.off), e.sx, e.sy, e.alpha, e.beta, ptr(tc))
        assert np.array_equal(dl.get(), exp)
        dcb, dcr = m.DeviceArray(pic), m.DeviceArray(pic2)
        lm.check(L.b2h264_k_deblock_chroma(dcb.ptr, dcr.ptr, djobs.ptr, n, None))
        ecb, ecr = pic.copy(), pic2.copy()
        for e in jobs_np:
            tc = np.array(list(e.tc), np.int8)
            if e.strong:
                orc.deblock_chroma_eq4(ptr(ecb, off=e.off), ptr(ecr, off=e.off), e.sx, e.sy, e.alpha, e.beta)
            else:
                orc.deblock_chroma_lt4(ptr(ecb, off=e.off), ptr(ecr, off=e.off), e.sx, e.sy, e.alpha, e.beta, ptr(tc))
        assert np.array_equal(dcb.get(), ecb) and np.array_equal(dcr.get(), ecr)


@pytest.mark.parametrize("whp", [(32, 16, 32), (176, 144, 32), (1920, 1088, 32), (16, 16, 16), (960, 544, 16)])
def test_expand_plane(env, whp):
    m, lm, L, orc = env
    w, h, pad = whp
    rng = np.random.RandomState(w + h)
    stride = w + 2 * pad
    pic = rng.randint(0, 256, size=(h + 2 * pad, stride)).astype(np.uint8)
    d = m.DeviceArray(pic)
    lm.check(L.b2h264_k_expand_plane(d.at(pad * stride + pad), stride, w, h, pad, None))
    exp = pic.copy()
    orc.expand_plane(ptr(exp, off=pad * stride + pad), stride, w, h, pad)
    assert np.array_equal(d.get(), exp)


def _padded_pair(w, h, pad=32):
    stride = w + 2 * pad
    cur = np.zeros((h + 2 * pad, stride), np.uint8)
    ref = np.zeros((h + 2 * pad, stride), np.uint8)
    cur[pad:pad + h, pad:pad + w] = h264lib.synth_frame(w, h, t=1)
    ref[pad:pad + h, pad:pad + w] = h264lib.synth_frame(w, h, t=0)
    h264lib.oracle().expand_plane(ptr(ref, off=pad * stride + pad), stride, w, h, pad)
    return cur, ref, stride, pad


@pytest.mark.parametrize("calc_satd", [0, 1])
def test_me_search(env, calc_satd):
    from openh264_b200.binding import MeJob as GJob, MeResult as GRes
    m, lm, L, orc = env
    w, h = 320, 192
    cur, ref, stride, pad = _padded_pair(w, h)
    rng = np.random.RandomState(900 + calc_satd)
    jobs = make_me_jobs(rng, w, h, 3000, calc_satd=calc_satd)
    n = len(jobs)
    garr = (GJob * n)()
    exp = []
    for i, (j, x, y) in enumerate(jobs):
        j.cur_off = (pad + y) * stride + pad + x
        j.ref_off = j.cur_off
        C.memmove(C.byref(garr[i]), C.byref(j), C.sizeof(GJob))
        r = h264lib.MeResult()
        orc.me_search(ptr(cur), stride, ptr(ref), stride, C.byref(j), C.byref(r))
        exp.append((r.mv_x, r.mv_y, r.sad_cost, r.satd_cost, r.ref_off))
    dcur, dref = m.DeviceArray(cur), m.DeviceArray(ref)
    djobs = m.DeviceArray(np.frombuffer(bytes(garr), dtype=np.uint8))
    dout = m.DeviceArray(shape=(n * C.sizeof(GRes),), dtype=np.uint8)
    lm.check(L.b2h264_k_me_search(dcur.ptr, stride, dref.ptr, stride, djobs.ptr, n, dout.ptr, None))
    res = (GRes * n).from_buffer_copy(dout.get().tobytes())
    got = [(r.mv_x, r.mv_y, r.sad_cost, r.satd_cost, r.ref_off) for r in res]
    assert got == exp


@pytest.mark.parametrize("wh,shift", [((320, 192), 0), ((1920, 1088), 0), ((320, 192), 4), ((336, 208), 0)])
def test_mc_sad_unit(env, wh, shift):
    """MC+SAD roofline unit.  Small frame: every cost against the oracle.  1080p (BASELINE.json's full
    size): a deterministic sample against the oracle plus two size-independent properties —
    (i) at integer MVs MC is a copy, so cost == plain SAD; (ii) cost(mv) for cur := ref is 0 at mv = 0.
    shift = 0: 16-byte aligned planes -> the tiled cp.async kernel; shift = 4: the per-warp fallback kernel;
    336x208: a picture that is not a whole number of 8x4-macroblock tiles."""
    m, lm, L, orc = env
    w, h = wh
    cur, ref, stride, pad = _padded_pair(w, h)
    mbw, mbh = w // 16, h // 16
    k = 9
    rng = np.random.RandomState(w)
    mv = rng.randint(-32, 33, size=(mbw * mbh, k, 2)).astype(np.int16)
    mv[:, 0] = 0
    mv[:, 1] = (mv[:, 1] // 4) * 4               # integer candidates
    o0 = pad * stride + pad + shift
    dcur, dref, dmv = m.DeviceArray(cur), m.DeviceArray(ref), m.DeviceArray(mv)
    cost = m.DeviceArray(shape=(mbw * mbh, k), dtype=np.int32)
    lm.check(L.b2h264_k_mc_sad(dcur.at(o0), stride, dref.at(o0), stride, mbw, mbh, dmv.ptr, k, cost.ptr, None))
    got = cost.get()
    step = 1 if w <= 320 else 37
    tmp = np.zeros((16, 16), np.uint8)
    for mb in range(0, mbw * mbh, step):
        x, y = (mb % mbw) * 16, (mb // mbw) * 16
        co = o0 + y * stride + x
        for c in range(k):
            mx, my = int(mv[mb, c, 0]), int(mv[mb, c, 1])
            ro = co + (my >> 2) * stride + (mx >> 2)
            orc.mc_luma(ptr(ref, off=ro), stride, ptr(tmp), 16, mx, my, 16, 16)
            assert got[mb, c] == orc.sad(0, ptr(cur, off=co), stride, ptr(tmp), 16), (mb, c)
    # property (i): integer candidates equal plain SAD for every MB
    for mb in range(0, mbw * mbh, 11):
        x, y = (mb % mbw) * 16, (mb // mbw) * 16
        co = o0 + y * stride + x
        mx, my = int(mv[mb, 1, 0]), int(mv[mb, 1, 1])
        assert got[mb, 1] == orc.sad(0, ptr(cur, off=co), stride, ptr(ref, off=co + (my >> 2) * stride + (mx >> 2)), stride)
    # property (ii): identical frames -> zero cost at zero MV everywhere
    lm.check(L.b2h264_k_mc_sad(dref.at(o0), stride, dref.at(o0), stride, mbw, mbh, dmv.ptr, k, cost.ptr, None))
    assert not cost.get()[:, 0].any()
