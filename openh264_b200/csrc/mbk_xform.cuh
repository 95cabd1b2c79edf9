// mbk_xform.cuh — per-thread 4x4 integer transform / quantisation / reconstruction primitives.
// One thread owns one 4x4 block (a warp covers the 16 luma + 8 chroma blocks of a macroblock).
// Replaces (semantics of) codec/encoder/core/src/encode_mb_aux.cpp:155-462,
// codec/encoder/core/src/decode_mb_aux.cpp:40-235 and codec/decoder/core/src/decode_mb_aux.cpp:42-190.
// The reference's int16 storage (and its wrap-around) is reproduced with explicit (int16_t) casts.
#pragma once
#include "mbk_common.cuh"

namespace mbk {

MBK_HD int16_t s16(int v) { return (int16_t)v; }

// residual + forward core transform (WelsDctT4_c, encode_mb_aux.cpp:313)
MBK_HD void dct4x4(int16_t d[16], const uint8_t* p1, int s1, const uint8_t* p2, int s2) {
  int16_t m[16];
#pragma unroll
  for (int y = 0; y < 4; y++) {
    const uint32_t wa = ld4u(p1 + y * s1), wb = ld4u(p2 + y * s2);
    const int r0 = (int)(wa & 0xff) - (int)(wb & 0xff), r1 = (int)((wa >> 8) & 0xff) - (int)((wb >> 8) & 0xff);
    const int r2 = (int)((wa >> 16) & 0xff) - (int)((wb >> 16) & 0xff), r3 = (int)(wa >> 24) - (int)(wb >> 24);
    const int a = r0 + r3, b = r1 + r2, c = r1 - r2, e = r0 - r3;
    m[4 * y] = s16(a + b); m[4 * y + 2] = s16(a - b); m[4 * y + 1] = s16(2 * e + c); m[4 * y + 3] = s16(e - 2 * c);
  }
#pragma unroll
  for (int x = 0; x < 4; x++) {
    const int16_t a = s16(m[x] + m[12 + x]), b = s16(m[4 + x] + m[8 + x]), c = s16(m[4 + x] - m[8 + x]),
                  e = s16(m[x] - m[12 + x]);
    d[x] = s16(a + b); d[8 + x] = s16(a - b); d[4 + x] = s16(2 * e + c); d[12 + x] = s16(e - 2 * c);
  }
}

// sign * (((ff + |x|) * mf) >> 16)   (encode_mb_aux.cpp:161-163)
MBK_HD int16_t quant1(int16_t x, int ff, int mf) {
  const int sign = x < 0 ? -1 : 0;
  const int mag = ((ff + ((sign ^ (int)x) - sign)) * mf) >> 16;
  return s16((sign ^ mag) - sign);
}
// quantise one 4x4 block in place; returns the block's largest magnitude as WelsQuantFour4x4Max_c does
MBK_HD int16_t quant4x4_max(int16_t d[16], const int16_t* ff, const int16_t* mf) {
  int16_t mx = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const int16_t x = d[i];
    const int sign = x < 0 ? -1 : 0;
    const int16_t mag = s16(((ff[i & 7] + ((sign ^ (int)x) - sign)) * mf[i & 7]) >> 16);
    if (mx < mag) mx = mag;
    d[i] = s16((sign ^ (int)mag) - sign);
  }
  return mx;
}
MBK_HD void quant4x4(int16_t d[16], const int16_t* ff, const int16_t* mf) {
#pragma unroll
  for (int i = 0; i < 16; i++) d[i] = quant1(d[i], ff[i & 7], mf[i & 7]);
}
MBK_HD void quant4x4_dc(int16_t d[16], int ff, int mf) {
#pragma unroll
  for (int i = 0; i < 16; i++) d[i] = quant1(d[i], ff, mf);
}

// chroma DC 2x2 Hadamard (+quant). in[4] = DC of blocks 0..3 (raster). (encode_mb_aux.cpp:226-277)
MBK_HD void hadamard2x2(const int16_t in[4], int16_t out[4]) {
  const int16_t s0 = s16(in[0] + in[2]), s1 = s16(in[0] - in[2]), s2 = s16(in[1] + in[3]), s3 = s16(in[1] - in[3]);
  out[0] = s16(s0 + s2); out[1] = s16(s0 - s2); out[2] = s16(s1 + s3); out[3] = s16(s1 - s3);
}
MBK_HD int hadamard_quant2x2_skip(const int16_t in[4], int16_t ff, int16_t mf) {
  const int16_t thr = s16(65535 / mf - ff);
  int16_t d[4];
  hadamard2x2(in, d);
  return iabs(d[0]) > thr || iabs(d[1]) > thr || iabs(d[2]) > thr || iabs(d[3]) > thr;
}
MBK_HD int hadamard_quant2x2(const int16_t in[4], int16_t ff, int16_t mf, int16_t out[4]) {
  int16_t d[4];
  hadamard2x2(in, d);
  int nz = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) { out[i] = quant1(d[i], ff, mf); nz += out[i] != 0; }
  return nz;
}

// I16x16 luma DC 4x4 Hadamard, (x+1)>>1, saturate (WelsHadamardT4Dc_c, encode_mb_aux.cpp:280).
// dc_in[k] = DC of 4x4 block k in the reference's coefficient storage order (4 groups of 4 = 8x8 z-order).
MBK_HD void hadamard_t4_dc(int16_t out[16], const int16_t dc_in[16]) {
  int p[16];
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    // reference reads pDct[idx], +16, +64, +80 with idx = ((i&8)<<4)+((i&4)<<3): blocks (idx>>4)+{0,1,4,5}
    const int k = ((i & 8) << 4 | (i & 4) << 3) >> 4;
    const int a = dc_in[k] + dc_in[k + 5], e = dc_in[k] - dc_in[k + 5];
    const int b = dc_in[k + 1] + dc_in[k + 4], c = dc_in[k + 1] - dc_in[k + 4];
    p[i] = a + b; p[i + 2] = a - b; p[i + 1] = e + c; p[i + 3] = e - c;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int a = p[i] + p[i + 12], e = p[i] - p[i + 12], b = p[i + 4] + p[i + 8], c = p[i + 4] - p[i + 8];
    out[i] = s16(clip3((a + b + 1) >> 1, -32768, 32767));
    out[i + 8] = s16(clip3((a - b + 1) >> 1, -32768, 32767));
    out[i + 4] = s16(clip3((e + c + 1) >> 1, -32768, 32767));
    out[i + 12] = s16(clip3((e - c + 1) >> 1, -32768, 32767));
  }
}

// frame zig-zag (WelsScan4x4DcAc_c / WelsScan4x4Ac_c, encode_mb_aux.cpp:371-401)
MBK_HD int zigzag_pos(int i) { return (int)((0xFEB7ADC963258410ull >> (4 * i)) & 0xf); }
MBK_HD void scan4x4_dcac(int16_t lv[16], const int16_t d[16]) {
#pragma unroll
  for (int i = 0; i < 16; i++) lv[i] = d[zigzag_pos(i)];
}
MBK_HD void scan4x4_ac(int16_t lv[16], const int16_t d[16]) {
#pragma unroll
  for (int i = 1; i < 16; i++) lv[i - 1] = d[zigzag_pos(i)];
  lv[15] = 0;
}
// JVT-O079 single-coefficient cost (WelsCalculateSingleCtr4x4_c, encode_mb_aux.cpp:418)
MBK_HD int single_ctr4x4(const int16_t lv[16]) {
  // every non-zero level costs T[number of zero levels directly below it], T = {3,2,2,1,1,1,0,...} (2 bits each)
  int total = 0, run = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    if (lv[i] != 0) { total += (0x56B >> (2 * run)) & 3; run = 0; }
    else run++;
  }
  return total;
}
MBK_HD int nonzero_count(const int16_t lv[16]) {
  int n = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) n += lv[i] != 0;
  return n;
}

// ---- dequant / inverse transforms (encoder/core/src/decode_mb_aux.cpp) ----
MBK_HD void dequant4x4(int16_t r[16], const uint16_t* mf) {
#pragma unroll
  for (int i = 0; i < 16; i++) r[i] = s16((int)r[i] * (int)mf[i & 7]);
}
MBK_HD void ihadamard4x4(int16_t r[16]) {   // butterflies only, int16 storage
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    const int16_t a = s16(r[i] + r[i + 2]), b = s16(r[i] - r[i + 2]), c = s16(r[i + 1] - r[i + 3]), e = s16(r[i + 1] + r[i + 3]);
    r[i] = s16(a + e); r[i + 1] = s16(b + c); r[i + 2] = s16(b - c); r[i + 3] = s16(a - e);
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int16_t a = s16(r[i] + r[8 + i]), b = s16(r[i] - r[8 + i]), c = s16(r[4 + i] - r[12 + i]), e = s16(r[4 + i] + r[12 + i]);
    r[i] = s16(a + e); r[4 + i] = s16(b + c); r[8 + i] = s16(b - c); r[12 + i] = s16(a - e);
  }
}
// WelsDequantIHadamard4x4_c (:98): inverse Hadamard then * mf (qp >= 12 path)
MBK_HD void dequant_ihadamard4x4(int16_t r[16], uint16_t mf) {
#pragma unroll
  for (int i = 0; i < 16; i += 4) {
    const int16_t a = s16(r[i] + r[i + 2]), b = s16(r[i] - r[i + 2]), c = s16(r[i + 1] - r[i + 3]), e = s16(r[i + 1] + r[i + 3]);
    r[i] = s16(a + e); r[i + 1] = s16(b + c); r[i + 2] = s16(b - c); r[i + 3] = s16(a - e);
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int16_t a = s16(r[i] + r[8 + i]), b = s16(r[i] - r[8 + i]), c = s16(r[4 + i] - r[12 + i]), e = s16(r[4 + i] + r[12 + i]);
    r[i] = s16((a + e) * (int)mf); r[4 + i] = s16((b + c) * (int)mf);
    r[8 + i] = s16((b - c) * (int)mf); r[12 + i] = s16((a - e) * (int)mf);
  }
}
// WelsDequantLumaDc4x4 (:80): qp < 12 luma DC scaling after WelsIHadamard4x4Dc
MBK_HD void dequant_luma_dc4x4(int16_t r[16], int qp) {
  const int v = tbl_dequant(qp % 6)[0];
  const int qf0 = qp / 6, sh = 2 - qf0, rnd = 1 << (1 - qf0);
#pragma unroll
  for (int i = 0; i < 16; i++) r[i] = s16(((int)r[i] * v + rnd) >> sh);
}
// WelsDequantIHadamard2x2Dc (:127)
MBK_HD void dequant_ihadamard2x2_dc(int16_t d[4], uint16_t mf) {
  const int16_t su = s16(d[0] + d[2]), du = s16(d[0] - d[2]), sd = s16(d[1] + d[3]), dd = s16(d[1] - d[3]);
  d[0] = s16(((su + sd) * (int)mf) >> 1); d[1] = s16(((su - sd) * (int)mf) >> 1);
  d[2] = s16(((du + dd) * (int)mf) >> 1); d[3] = s16(((du - dd) * (int)mf) >> 1);
}
// WelsIDctT4Rec_c (:164): inverse core transform + prediction + clip; row pass stored as int16
MBK_HD void idct4x4_rec(uint8_t* rec, int rs, const uint8_t* pred, int ps, const int16_t c[16]) {
  int16_t t[16];
#pragma unroll
  for (int y = 0; y < 4; y++) {
    const int su = c[4 * y] + c[4 * y + 2], du = c[4 * y] - c[4 * y + 2];
    const int sd = c[4 * y + 1] + (c[4 * y + 3] >> 1), dd = (c[4 * y + 1] >> 1) - c[4 * y + 3];
    t[4 * y] = s16(su + sd); t[4 * y + 1] = s16(du + dd); t[4 * y + 2] = s16(du - dd); t[4 * y + 3] = s16(su - sd);
  }
#pragma unroll
  for (int x = 0; x < 4; x++) {
    const int sl = t[x] + t[8 + x], dl = t[x] - t[8 + x], dr = (t[4 + x] >> 1) - t[12 + x], sr = t[4 + x] + (t[12 + x] >> 1);
    rec[x] = (uint8_t)clip255(pred[x] + ((sl + sr + 32) >> 6));
    rec[rs + x] = (uint8_t)clip255(pred[ps + x] + ((dl + dr + 32) >> 6));
    rec[2 * rs + x] = (uint8_t)clip255(pred[2 * ps + x] + ((dl - dr + 32) >> 6));
    rec[3 * rs + x] = (uint8_t)clip255(pred[3 * ps + x] + ((sl - sr + 32) >> 6));
  }
}

// ---- decoder: IdctResAddPred_c / IdctResAddPred8x8_c (decoder/core/src/decode_mb_aux.cpp:42,79) ----
MBK_HD void idct_res_add_pred(uint8_t* pred, int stride, const int16_t rs[16]) {
  int16_t t[16];
#pragma unroll
  for (int y = 0; y < 4; y++) {
    const int e0 = rs[4 * y] + rs[4 * y + 2], e1 = rs[4 * y] - rs[4 * y + 2];
    const int e2 = (rs[4 * y + 1] >> 1) - rs[4 * y + 3], e3 = rs[4 * y + 1] + (rs[4 * y + 3] >> 1);
    t[4 * y] = s16(e0 + e3); t[4 * y + 1] = s16(e1 + e2); t[4 * y + 2] = s16(e1 - e2); t[4 * y + 3] = s16(e0 - e3);
  }
#pragma unroll
  for (int x = 0; x < 4; x++) {
    const int a = t[x] + t[8 + x], b = t[4 + x] + (t[12 + x] >> 1), c = t[x] - t[8 + x], d = (t[4 + x] >> 1) - t[12 + x];
    pred[x] = (uint8_t)clip255(((32 + a + b) >> 6) + pred[x]);
    pred[3 * stride + x] = (uint8_t)clip255(((32 + a - b) >> 6) + pred[3 * stride + x]);
    pred[stride + x] = (uint8_t)clip255(((32 + c + d) >> 6) + pred[stride + x]);
    pred[2 * stride + x] = (uint8_t)clip255(((32 + c - d) >> 6) + pred[2 * stride + x]);
  }
}
MBK_HD void idct8_1d(const int16_t p[8], int16_t o[8]) {   // all int16, as the reference
  int16_t a0 = s16(p[0] + p[4]), a1 = s16(p[0] - p[4]), a2 = s16(p[6] - (p[2] >> 1)), a3 = s16(p[2] + (p[6] >> 1));
  const int16_t b0 = s16(a0 + a3), b2 = s16(a1 - a2), b4 = s16(a1 + a2), b6 = s16(a0 - a3);
  a0 = s16(-p[3] + p[5] - p[7] - (p[7] >> 1));
  a1 = s16(p[1] + p[7] - p[3] - (p[3] >> 1));
  a2 = s16(-p[1] + p[7] + p[5] + (p[5] >> 1));
  a3 = s16(p[3] + p[5] + p[1] + (p[1] >> 1));
  const int16_t b1 = s16(a0 + (a3 >> 2)), b3 = s16(a1 + (a2 >> 2)), b5 = s16(a2 - (a1 >> 2)), b7 = s16(a3 - (a0 >> 2));
  o[0] = s16(b0 + b7); o[1] = s16(b2 - b5); o[2] = s16(b4 + b3); o[3] = s16(b6 + b1);
  o[4] = s16(b6 - b1); o[5] = s16(b4 - b3); o[6] = s16(b2 + b5); o[7] = s16(b0 - b7);
}

}  // namespace mbk
