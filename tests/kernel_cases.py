"""Seeded per-kernel cases, in the spirit of the reference's test/encoder/EncUT_*.cpp and
test/decoder/DecUT_*.cpp (random buffers -> bit-exact comparison).  Each case takes a library object
exposing the shared oracle/reference signatures (tests/h264lib.py) and returns the output arrays, so
the same case can be run on the oracle, on the compiled reference and hashed into tests/golden/.
"""
import ctypes as C

import numpy as np

from h264lib import BLK_DIMS, MeJob, MeResult, ptr, synth_frame

STRIDE = 64


def _planes(rng, n=2, h=48):
    return [rng.randint(0, 256, size=(h, STRIDE)).astype(np.uint8) for _ in range(n)]


def _coef(rng, n, lo=-2048, hi=2048):
    return rng.randint(lo, hi, size=n).astype(np.int16)


def case_sad(lib, rng):
    out = []
    for blk in range(7):
        for _ in range(8):
            a, b = _planes(rng)
            oa, ob = 16 * STRIDE + 16 + rng.randint(0, 8), 16 * STRIDE + 16 + rng.randint(0, 8)
            out.append(lib.sad(blk, ptr(a, off=oa), STRIDE, ptr(b, off=ob), STRIDE))
            out.append(lib.satd(blk, ptr(a, off=oa), STRIDE, ptr(b, off=ob), STRIDE))
            four = np.zeros(4, np.int32)
            lib.sad_four(blk, ptr(a, off=oa), STRIDE, ptr(b, off=ob), STRIDE, ptr(four))
            out.extend(four.tolist())
    # flat / extreme blocks (EncUT_Sample.cpp hand-written expectations use constant differences)
    a = np.full((48, STRIDE), 255, np.uint8)
    b = np.zeros((48, STRIDE), np.uint8)
    for blk in range(7):
        out.append(lib.sad(blk, ptr(a, off=1040), STRIDE, ptr(b, off=1040), STRIDE))
        out.append(lib.satd(blk, ptr(a, off=1040), STRIDE, ptr(b, off=1040), STRIDE))
    return [np.array(out, np.int64)]


def case_mc_luma(lib, rng):
    outs = []
    for (w, h) in [(16, 16), (16, 8), (8, 16), (8, 8), (8, 4), (4, 8), (4, 4)]:
        src = _planes(rng, 1)[0]
        for fx in range(4):
            for fy in range(4):
                dst = np.zeros((16, 16), np.uint8)
                lib.mc_luma(ptr(src, off=12 * STRIDE + 12), STRIDE, ptr(dst), 16, fx, fy, w, h)
                outs.append(dst)
    # saturating content: alternating 0/255 drives the 6-tap filter through both clip ends
    src = ((np.indices((48, STRIDE)).sum(0) & 1) * 255).astype(np.uint8)
    for fx in range(4):
        for fy in range(4):
            dst = np.zeros((16, 16), np.uint8)
            lib.mc_luma(ptr(src, off=12 * STRIDE + 12), STRIDE, ptr(dst), 16, fx, fy, 16, 16)
            outs.append(dst)
    return outs


def case_mc_chroma(lib, rng):
    outs = []
    for (w, h) in [(8, 8), (8, 4), (4, 8), (4, 4), (4, 2), (2, 4), (2, 2)]:
        src = _planes(rng, 1)[0]
        for dx in range(8):
            for dy in range(8):
                dst = np.zeros((8, 8), np.uint8)
                lib.mc_chroma(ptr(src, off=8 * STRIDE + 8), STRIDE, ptr(dst), 8, dx, dy, w, h)
                outs.append(dst)
    return outs


def case_pixel_avg(lib, rng):
    outs = []
    for (w, h) in [(16, 16), (8, 8), (4, 4), (16, 8)]:
        a, b = _planes(rng)
        dst = np.zeros((16, 16), np.uint8)
        lib.pixel_avg(ptr(dst), 16, ptr(a, off=70), STRIDE, ptr(b, off=133), STRIDE, w, h)
        outs.append(dst)
    return outs


def case_dct_quant(lib, rng):
    outs = []
    for _ in range(16):
        a, b = _planes(rng)
        d = np.zeros(64, np.int16)
        lib.dct_four4x4(ptr(d), ptr(a, off=130), STRIDE, ptr(b, off=261), STRIDE)
        outs.append(d.copy())
        d1 = np.zeros(16, np.int16)
        lib.dct4x4(ptr(d1), ptr(a, off=130), STRIDE, ptr(b, off=261), STRIDE)
        outs.append(d1)
        for qp in (0, 11, 12, 26, 40, 51):
            for intra in (0, 6):
                ff = np.ctypeslib.as_array(lib.quant_ff(qp + intra), shape=(8,)).copy()
                mf = np.ctypeslib.as_array(lib.quant_mf(qp), shape=(8,)).copy()
                q = d.copy()
                lib.quant_four4x4(ptr(q), ptr(ff), ptr(mf))
                outs.append(q)
                q2 = d.copy()
                mx = np.zeros(4, np.int16)
                lib.quant_four4x4_max(ptr(q2), ptr(ff), ptr(mf), ptr(mx))
                outs.extend([q2, mx])
                q3 = d[:16].copy()
                lib.quant4x4(ptr(q3), ptr(ff), ptr(mf))
                outs.append(q3)
                q4 = d[16:32].copy()
                lib.quant4x4_dc(ptr(q4), int(ff[0]) << 1, int(mf[0]) >> 1)
                outs.append(q4)
    # full-range int16 input exercises the truncating stores
    for _ in range(8):
        d = _coef(rng, 64, -32768, 32768)
        ff = np.ctypeslib.as_array(lib.quant_ff(20), shape=(8,)).copy()
        mf = np.ctypeslib.as_array(lib.quant_mf(20), shape=(8,)).copy()
        q = d.copy()
        mx = np.zeros(4, np.int16)
        lib.quant_four4x4_max(ptr(q), ptr(ff), ptr(mf), ptr(mx))
        outs.extend([q, mx])
    return outs


def case_hadamard(lib, rng):
    outs = []
    for _ in range(16):
        mb = _coef(rng, 256, -4096, 4096)
        dc = np.zeros(16, np.int16)
        lib.hadamard_t4_dc(ptr(dc), ptr(mb))
        outs.append(dc)
        ch = _coef(rng, 64, -2048, 2048)
        for qp in (10, 26, 39):
            ff = int(np.ctypeslib.as_array(lib.quant_ff(qp), shape=(8,))[0]) << 1
            mf = int(np.ctypeslib.as_array(lib.quant_mf(qp), shape=(8,))[0]) >> 1
            outs.append(np.array([lib.hadamard_quant2x2_skip(ptr(ch), ff, mf)]))
            c2 = ch.copy()
            dct = np.zeros(4, np.int16)
            blk = np.zeros(4, np.int16)
            nz = lib.hadamard_quant2x2(ptr(c2), ff, mf, ptr(dct), ptr(blk))
            outs.extend([c2, dct, blk, np.array([nz])])
    small = np.zeros(64, np.int16)
    small[[0, 16, 32, 48]] = [3, -2, 1, 0]
    outs.append(np.array([lib.hadamard_quant2x2_skip(ptr(small), 22, 400)]))
    return outs


def case_scan(lib, rng):
    outs = []
    for i in range(32):
        d = _coef(rng, 16, -3, 4) if i % 2 else (_coef(rng, 16, -1, 2) * (rng.rand(16) < 0.2)).astype(np.int16)
        l1 = np.zeros(16, np.int16)
        l2 = np.zeros(16, np.int16)
        lib.scan4x4_dcac(ptr(l1), ptr(d))
        lib.scan4x4_ac(ptr(l2), ptr(d))
        outs.extend([l1, l2, np.array([lib.single_ctr4x4(ptr(l1)), lib.nonzero_count(ptr(l1))])])
    z = np.zeros(16, np.int16)
    outs.append(np.array([lib.single_ctr4x4(ptr(z)), lib.nonzero_count(ptr(z))]))
    return outs


def case_dequant_idct(lib, rng):
    outs = []
    for i in range(24):
        wide = i >= 16
        lv = _coef(rng, 64, -32768, 32768) if wide else _coef(rng, 64, -40, 41)
        for qp in (0, 5, 11, 12, 26, 51):
            mf = np.ctypeslib.as_array(lib.dequant_coeff(qp), shape=(8,)).copy()
            r = lv.copy()
            lib.dequant_four4x4(ptr(r), ptr(mf))
            outs.append(r.copy())
            r1 = lv[:16].copy()
            lib.dequant4x4(ptr(r1), ptr(mf))
            outs.append(r1)
            pred, _ = _planes(rng)
            rec = np.zeros((8, 8), np.uint8)
            lib.idct_four4x4_rec(ptr(rec), 8, ptr(pred, off=200), STRIDE, ptr(r))
            outs.append(rec)
            rec1 = np.zeros((4, 4), np.uint8)
            lib.idct4x4_rec(ptr(rec1), 4, ptr(pred, off=200), STRIDE, ptr(r))
            outs.append(rec1)
            h4 = lv[:16].copy()
            lib.dequant_ihadamard4x4(ptr(h4), int(mf[0]) >> 2 if qp >= 12 else int(mf[0]))
            outs.append(h4)
            h2 = lv[:4].copy()
            lib.dequant_ihadamard2x2_dc(ptr(h2), int(mf[0]))
            outs.append(h2)
            if qp < 12:
                h = lv[16:32].copy()
                lib.ihadamard4x4_dc(ptr(h))
                outs.append(h.copy())
                lib.dequant_luma_dc4x4(ptr(h), qp)
                outs.append(h)
            rec16 = np.zeros((16, 16), np.uint8)
            lib.idct_rec_i16x16_dc(ptr(rec16), 16, ptr(pred, off=64), STRIDE, ptr(r[:16]))
            outs.append(rec16)
    return outs


def case_idct_res_add_pred(lib, rng):
    outs = []
    for i in range(48):
        rs = _coef(rng, 64, -32768, 32768) if i % 4 == 3 else _coef(rng, 64, -2000, 2001)
        p, _ = _planes(rng)
        a = p.copy()
        lib.idct_res_add_pred(ptr(a, off=3 * STRIDE + 20), STRIDE, ptr(rs))
        b = p.copy()
        lib.idct_res_add_pred8x8(ptr(b, off=3 * STRIDE + 20), STRIDE, ptr(rs))
        outs.extend([a, b])
    return outs


def case_deblock(lib, rng):
    outs = []
    for i in range(64):
        # smooth-ish content so the alpha/beta gates open on many lines
        base = rng.randint(40, 200)
        pic = np.clip(base + rng.randint(-12, 13, size=(48, STRIDE)), 0, 255).astype(np.uint8)
        pic2 = np.clip(base + rng.randint(-6, 7, size=(48, STRIDE)), 0, 255).astype(np.uint8)
        alpha, beta = int(rng.randint(4, 80)), int(rng.randint(2, 18))
        tc = rng.randint(-1, 6, size=4).astype(np.int8)
        off = 16 * STRIDE + 16
        for (sx, sy) in ((STRIDE, 1), (1, STRIDE)):
            a = pic.copy()
            lib.deblock_luma_lt4(ptr(a, off=off), sx, sy, alpha, beta, ptr(tc))
            b = pic.copy()
            lib.deblock_luma_eq4(ptr(b, off=off), sx, sy, alpha, beta)
            c, d = pic.copy(), pic2.copy()
            lib.deblock_chroma_lt4(ptr(c, off=off), ptr(d, off=off), sx, sy, alpha, beta, ptr(tc))
            e, f = pic.copy(), pic2.copy()
            lib.deblock_chroma_eq4(ptr(e, off=off), ptr(f, off=off), sx, sy, alpha, beta)
            outs.extend([a, b, c, d, e, f])
    return outs


def case_expand(lib, rng):
    outs = []
    for (w, h, pad) in [(32, 16, 32), (48, 32, 32), (16, 16, 16), (24, 8, 16)]:
        stride = w + 2 * pad
        pic = rng.randint(0, 256, size=(h + 2 * pad, stride)).astype(np.uint8)
        lib.expand_plane(ptr(pic, off=pad * stride + pad), stride, w, h, pad)
        outs.append(pic)
    return outs


def make_me_jobs(rng, w, h, n, stride_pad=32, qp_choices=(20, 26, 34), calc_satd=1):
    """Random but valid ME jobs on a (w x h) frame pair with `stride_pad` pixels of padding."""
    jobs = []
    mbw, mbh = w // 16, h // 16
    for _ in range(n):
        blk = int(rng.randint(0, 4)) if rng.rand() < 0.85 else int(rng.randint(4, 7))
        bw, bh = BLK_DIMS[blk]
        mbx, mby = int(rng.randint(0, mbw)), int(rng.randint(0, mbh))
        ox = int(rng.randint(0, 16 // bw)) * bw
        oy = int(rng.randint(0, 16 // bh)) * bh
        j = MeJob()
        j.blk = blk
        # window per SetMvWithinIntegerMvRange (svc_motion_estimate.h:345-353), range 16 keeps reads inside padding
        rngmv = 16
        j.mv_min_x = max(-((mbx + 1) << 4) + 3, -rngmv)
        j.mv_min_y = max(-((mby + 1) << 4) + 3, -rngmv)
        j.mv_max_x = min(((mbw - mbx) << 4) - 3, rngmv)
        j.mv_max_y = min(((mbh - mby) << 4) - 3, rngmv)
        j.mvp_x, j.mvp_y = int(rng.randint(-40, 41)), int(rng.randint(-40, 41))
        j.n_mvc = int(rng.randint(0, 6))
        for k in range(5):
            j.mvc[k][0], j.mvc[k][1] = int(rng.randint(-80, 81)), int(rng.randint(-80, 81))
        j.sad_pred = int(rng.choice([0, 0, 300, 2000, 100000]))
        j.qp = int(rng.choice(qp_choices))
        j.calc_satd = calc_satd
        jobs.append((j, mbx * 16 + ox, mby * 16 + oy))
    return jobs


def case_me_search(lib, rng):
    w, h, pad = 128, 96, 32
    stride = w + 2 * pad
    cur = np.zeros((h + 2 * pad, stride), np.uint8)
    ref = np.zeros((h + 2 * pad, stride), np.uint8)
    cur[pad:pad + h, pad:pad + w] = synth_frame(w, h, t=1)
    ref[pad:pad + h, pad:pad + w] = synth_frame(w, h, t=0)
    # replicate-pad the reference like ExpandReferencingPicture would
    ref[:pad] = ref[pad]
    ref[pad + h:] = ref[pad + h - 1]
    ref[:, :pad] = ref[:, pad:pad + 1]
    ref[:, pad + w:] = ref[:, pad + w - 1:pad + w]
    out = []
    for (j, x, y) in make_me_jobs(rng, w, h, 200):
        j.cur_off = (pad + y) * stride + pad + x
        j.ref_off = j.cur_off
        r = MeResult()
        lib.me_search(ptr(cur), stride, ptr(ref), stride, C.byref(j), C.byref(r))
        out.extend([r.mv_x, r.mv_y, r.sad_cost, r.satd_cost, r.ref_off])
    return [np.array(out, np.int64)]


def case_tables(lib, rng):
    outs = []
    for q in range(58):
        outs.append(np.ctypeslib.as_array(lib.quant_ff(q), shape=(8,)).copy())
    for q in range(52):
        outs.append(np.ctypeslib.as_array(lib.quant_mf(q), shape=(8,)).copy())
        outs.append(np.ctypeslib.as_array(lib.dequant_coeff(q), shape=(8,)).copy())
        outs.append(np.array([lib.qp_lambda(q), lib.chroma_qp(q)]))
    return outs


CASES = {
    "tables": case_tables,
    "sad_satd": case_sad,
    "mc_luma": case_mc_luma,
    "mc_chroma": case_mc_chroma,
    "pixel_avg": case_pixel_avg,
    "dct_quant": case_dct_quant,
    "hadamard": case_hadamard,
    "scan": case_scan,
    "dequant_idct": case_dequant_idct,
    "idct_res_add_pred": case_idct_res_add_pred,
    "deblock": case_deblock,
    "expand": case_expand,
    "me_search": case_me_search,
}


def run_case(name, lib, seed=1234):
    rng = np.random.RandomState(seed + sum(map(ord, name)))
    return CASES[name](lib, rng)
