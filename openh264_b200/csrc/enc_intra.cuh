// enc_intra.cuh — intra prediction, intra mode decision and intra residual coding for one macroblock
// owned by one warp.  Restates (not ports) the decision order of
//   WelsMdI16x16 / WelsMdI4x4 / WelsMdIntraChroma      codec/encoder/core/src/svc_base_layer_md.cpp:365,418,867
//   WelsEncRecI16x16Y / WelsEncRecI4x4Y / WelsEncRecUV  codec/encoder/core/src/svc_encode_mb.cpp:54,139,244
//   the predictors of codec/encoder/core/src/get_intra_predictor.cpp (H.264 8.3.1-8.3.4)
// Every tie-break (mode list order, strict '<', the early break of the I4x4 loop) is observable in
// the bitstream and is kept.
#pragma once
#include "enc_types.h"
#include "mbk_common.cuh"
#include "mbk_sad.cuh"
#include "mbk_xform.cuh"

namespace mbk {

// ---- reconstruction tile ----------------------------------------------------------------------
// The macroblock is reconstructed in a small tile that also carries the neighbouring samples intra
// prediction needs: luma rows -1..15 x cols -1..23 (pitch 32), chroma rows -1..7 x cols -1..7 (pitch 16).
#define TY_PITCH 32
#define TC_PITCH 16
struct RecTile {
  uint8_t y[17 * TY_PITCH];
  uint8_t u[9 * TC_PITCH];
  uint8_t v[9 * TC_PITCH];
};
MBK_HD uint8_t* tile_y(RecTile& t, int x, int y) { return t.y + (y + 1) * TY_PITCH + (x + 1); }
MBK_HD uint8_t* tile_c(uint8_t* plane, int x, int y) { return plane + (y + 1) * TC_PITCH + (x + 1); }

// neighbour availability bits (wels_common_defs.h: LEFT 1, TOP 2, TOPLEFT 4, TOPRIGHT 8 as used by uiNeighborIntra)
enum { NB_LEFT = 1, NB_TOP = 2, NB_TOPLEFT = 4, NB_TOPRIGHT = 8 };

// mode ids follow the reference (wels_common_defs.h:330-371)
enum { I16_V = 0, I16_H, I16_DC, I16_P, I16_DC_L, I16_DC_T, I16_DC_128 };
enum { I4_V = 0, I4_H, I4_DC, I4_DDL, I4_DDR, I4_VR, I4_HD, I4_VL, I4_HU, I4_DC_L, I4_DC_T, I4_DC_128 };
enum { C_DC = 0, C_H, C_V, C_P, C_DC_L, C_DC_T, C_DC_128 };

MBK_HD int map_i16(int m) { return m <= 3 ? m : 2; }        // g_kiMapModeI16x16
MBK_HD int map_i4(int m) { return m <= 8 ? m : 2; }         // g_kiMapModeI4x4 (no I4_PRED_MODE_EXTEND)
MBK_HD int map_chroma(int m) { return m <= 3 ? m : 0; }     // g_kiMapModeIntraChroma
MBK_HD int ue_bits(int v) { return 2 * (31 - clz32((uint32_t)v + 1)) + 1; }

// ---- 16x16 / 8x8 predictors: one sample ---------------------------------------------------------
struct PlaneCoef { int a, b, c; };
MBK_HD PlaneCoef plane_coef(const uint8_t* org, int pitch, int n /*16 or 8*/) {
  const uint8_t* top = org - pitch;
  const uint8_t* left = org - 1;
  const int h = n >> 1;
  int ts = 0, ls = 0;
  for (int i = 0; i < h; i++) {
    ts += (i + 1) * ((int)top[h + i] - (int)top[h - 2 - i]);
    ls += (i + 1) * ((int)left[(h + i) * pitch] - (int)left[(h - 2 - i) * pitch]);
  }
  PlaneCoef pc;
  pc.a = ((int)left[(n - 1) * pitch] + (int)top[n - 1]) << 4;
  if (n == 16) { pc.b = (5 * ts + 32) >> 6; pc.c = (5 * ls + 32) >> 6; }
  else { pc.b = (17 * ts + 16) >> 5; pc.c = (17 * ls + 16) >> 5; }
  return pc;
}

// fills a 16x16 luma prediction (dst stride 16) for mode m from the tile neighbours
MBK_FN void pred_i16(uint8_t* dst, const uint8_t* org, int pitch, int m) {
  int dc = 128;
  PlaneCoef pc = {0, 0, 0};
  if (m == I16_DC || m == I16_DC_L || m == I16_DC_T) {
    int st = 0, sl = 0;
    for (int i = 0; i < 16; i++) { st += org[-pitch + i]; sl += org[i * pitch - 1]; }
    dc = m == I16_DC ? (st + sl + 16) >> 5 : m == I16_DC_L ? (sl + 8) >> 4 : (st + 8) >> 4;
  } else if (m == I16_P) {
    pc = plane_coef(org, pitch, 16);
  }
  for (int i = lane_id(); i < 256; i += MBK_WS) {
    const int y = i >> 4, x = i & 15;
    int v;
    if (m == I16_V) v = org[-pitch + x];
    else if (m == I16_H) v = org[y * pitch - 1];
    else if (m == I16_P) v = clip255((pc.a + pc.b * (x - 7) + pc.c * (y - 7) + 16) >> 5);
    else v = dc;
    dst[i] = (uint8_t)v;
  }
  warp_sync();
}

// fills an 8x8 chroma prediction (dst stride 8) for mode m
MBK_FN void pred_chroma(uint8_t* dst, const uint8_t* org, int pitch, int m) {
  int dcq[4] = {128, 128, 128, 128};   // per 4x4 quadrant: TL, TR, BL, BR
  PlaneCoef pc = {0, 0, 0};
  if (m == C_DC || m == C_DC_L || m == C_DC_T) {
    int t0 = 0, t1 = 0, l0 = 0, l1 = 0;
    if (m != C_DC_L) for (int i = 0; i < 4; i++) { t0 += org[-pitch + i]; t1 += org[-pitch + 4 + i]; }
    if (m != C_DC_T) for (int i = 0; i < 4; i++) { l0 += org[i * pitch - 1]; l1 += org[(4 + i) * pitch - 1]; }
    if (m == C_DC) {
      dcq[0] = (t0 + l0 + 4) >> 3; dcq[1] = (t1 + 2) >> 2; dcq[2] = (l1 + 2) >> 2; dcq[3] = (t1 + l1 + 4) >> 3;
    } else if (m == C_DC_L) {
      dcq[0] = dcq[1] = (l0 + 2) >> 2; dcq[2] = dcq[3] = (l1 + 2) >> 2;
    } else {
      dcq[0] = dcq[2] = (t0 + 2) >> 2; dcq[1] = dcq[3] = (t1 + 2) >> 2;
    }
  } else if (m == C_P) {
    pc = plane_coef(org, pitch, 8);
  }
  for (int i = lane_id(); i < 64; i += MBK_WS) {
    const int y = i >> 3, x = i & 7;
    int v;
    if (m == C_V) v = org[-pitch + x];
    else if (m == C_H) v = org[y * pitch - 1];
    else if (m == C_P) v = clip255((pc.a + pc.b * (x - 3) + pc.c * (y - 3) + 16) >> 5);
    else v = dcq[(y >> 2) * 2 + (x >> 2)];
    dst[i] = (uint8_t)v;
  }
  warp_sync();
}

// ---- 4x4 predictors: all 16 samples by one thread ----------------------------------------------
MBK_FN void pred_i4(uint8_t p[16], const uint8_t* org, int pitch, int m) {
  // e[4] = top-left, e[5..12] = top (incl. top-right), e[3..0] = left 0..3
  int t[8], l[4], lt = 0;
  const bool need_top = !(m == I4_H || m == I4_HU || m == I4_DC_L || m == I4_DC_128);
  const bool need_left = !(m == I4_V || m == I4_DDL || m == I4_VL || m == I4_DC_T || m == I4_DC_128);
  const bool need_tr = (m == I4_DDL || m == I4_VL);
  const bool need_lt = (m == I4_DDR || m == I4_VR || m == I4_HD);
#pragma unroll
  for (int i = 0; i < 4; i++) { t[i] = need_top ? org[-pitch + i] : 0; l[i] = need_left ? org[i * pitch - 1] : 0; }
#pragma unroll
  for (int i = 4; i < 8; i++) t[i] = need_tr ? org[-pitch + i] : 0;
  if (need_lt) lt = org[-pitch - 1];
#define T(i) ((i) < 0 ? lt : t[i])
#define L(i) ((i) < 0 ? lt : l[i])
#pragma unroll
  for (int y = 0; y < 4; y++) {
#pragma unroll
    for (int x = 0; x < 4; x++) {
      int v;
      switch (m) {
        case I4_V: v = t[x]; break;
        case I4_H: v = l[y]; break;
        case I4_DC: v = (t[0] + t[1] + t[2] + t[3] + l[0] + l[1] + l[2] + l[3] + 4) >> 3; break;
        case I4_DC_L: v = (l[0] + l[1] + l[2] + l[3] + 2) >> 2; break;
        case I4_DC_T: v = (t[0] + t[1] + t[2] + t[3] + 2) >> 2; break;
        case I4_DDL:
          v = (x == 3 && y == 3) ? (t[6] + 3 * t[7] + 2) >> 2 : (t[x + y] + 2 * t[x + y + 1] + t[x + y + 2] + 2) >> 2;
          break;
        case I4_DDR:
          if (x > y) v = (T(x - y - 2) + 2 * T(x - y - 1) + T(x - y) + 2) >> 2;
          else if (x < y) v = (L(y - x - 2) + 2 * L(y - x - 1) + L(y - x) + 2) >> 2;
          else v = (t[0] + 2 * lt + l[0] + 2) >> 2;
          break;
        case I4_VR: {
          const int z = 2 * x - y, k = x - (y >> 1);
          if (z >= 0 && !(z & 1)) v = (T(k - 1) + T(k) + 1) >> 1;
          else if (z >= 0) v = (T(k - 2) + 2 * T(k - 1) + T(k) + 2) >> 2;
          else if (z == -1) v = (l[0] + 2 * lt + t[0] + 2) >> 2;
          else v = (L(y - 1) + 2 * L(y - 2) + L(y - 3) + 2) >> 2;
          break;
        }
        case I4_HD: {
          const int z = 2 * y - x, k = y - (x >> 1);
          if (z >= 0 && !(z & 1)) v = (L(k - 1) + L(k) + 1) >> 1;
          else if (z >= 0) v = (L(k - 2) + 2 * L(k - 1) + L(k) + 2) >> 2;
          else if (z == -1) v = (l[0] + 2 * lt + t[0] + 2) >> 2;
          else v = (T(x - 1) + 2 * T(x - 2) + T(x - 3) + 2) >> 2;
          break;
        }
        case I4_VL: {
          const int k = x + (y >> 1);
          v = (y & 1) ? (t[k] + 2 * t[k + 1] + t[k + 2] + 2) >> 2 : (t[k] + t[k + 1] + 1) >> 1;
          break;
        }
        case I4_HU: {
          const int z = x + 2 * y, k = y + (x >> 1);
          if (z > 5) v = l[3];
          else if (z == 5) v = (l[2] + 3 * l[3] + 2) >> 2;
          else if (z & 1) v = (l[k] + 2 * l[k + 1] + l[k + 2] + 2) >> 2;
          else v = (l[k] + l[k + 1] + 1) >> 1;
          break;
        }
        default: v = 128; break;
      }
      p[4 * y + x] = (uint8_t)v;
    }
  }
#undef T
#undef L
}

// SATD of a 4x4 block: cur (pixels, stride cs) against a prediction held in registers
MBK_FN int satd4x4_pred(const uint8_t p[16], const uint8_t* cur, int cs) {
  int t[4][4];
#pragma unroll
  for (int y = 0; y < 4; y++) {
    // the reference calls pfSampleSatd(pred, 4, enc, stride): difference = pred - enc
    const int d0 = (int)p[4 * y] - cur[y * cs], d1 = (int)p[4 * y + 1] - cur[y * cs + 1];
    const int d2 = (int)p[4 * y + 2] - cur[y * cs + 2], d3 = (int)p[4 * y + 3] - cur[y * cs + 3];
    const int e0 = d0 + d2, e1 = d1 + d3, e2 = d0 - d2, e3 = d1 - d3;
    t[y][0] = e0 + e1; t[y][1] = e2 + e3; t[y][2] = e2 - e3; t[y][3] = e0 - e1;
  }
  int sum = 0;
#pragma unroll
  for (int x = 0; x < 4; x++) {
    const int e0 = t[0][x] + t[2][x], e1 = t[1][x] + t[3][x], e2 = t[0][x] - t[2][x], e3 = t[1][x] - t[3][x];
    sum += iabs(e0 + e1) + iabs(e2 + e3) + iabs(e2 - e3) + iabs(e0 - e1);
  }
  return (sum + 1) >> 1;
}

// ---- mode lists (svc_base_layer_md.cpp:48-221) --------------------------------------------------
// I16x16 / chroma candidates by (left | top<<1 | topleft<<2), in the reference's evaluation order
MBK_HD int i16_modes(int nb3, int out[4]) {
  const bool l = nb3 & 1, t = nb3 & 2;
  if (l && t) { out[0] = I16_V; out[1] = I16_H; out[2] = I16_DC; out[3] = I16_P; return (nb3 & 4) ? 4 : 3; }
  if (l) { out[0] = I16_DC_L; out[1] = I16_H; return 2; }
  if (t) { out[0] = I16_DC_T; out[1] = I16_V; return 2; }
  out[0] = I16_DC_128; return 1;
}
MBK_HD int chroma_modes(int nb3, int out[4]) {
  const bool l = nb3 & 1, t = nb3 & 2;
  if (l && t) { out[0] = C_V; out[1] = C_H; out[2] = C_DC; out[3] = C_P; return (nb3 & 4) ? 4 : 3; }
  if (l) { out[0] = C_DC_L; out[1] = C_H; return 2; }
  if (t) { out[0] = C_DC_T; out[1] = C_V; return 2; }
  out[0] = C_DC_128; return 1;
}
// I4x4 candidates by the block's own availability nibble (left | top<<1 | topleft<<2 | topright<<3)
MBK_HD int i4_modes(int av, int out[9]) {
  const bool l = av & 1, t = av & 2, tl = av & 4, tr = av & 8;
  if (!l && !t) { out[0] = I4_DC_128; return 1; }
  if (l && !t) { out[0] = I4_DC_L; out[1] = I4_H; out[2] = I4_HU; return 3; }
  if (!l && t) {
    out[0] = I4_DC_T; out[1] = I4_V;
    if (tr) { out[2] = I4_DDL; out[3] = I4_VL; return 4; }
    return 2;
  }
  int n = 0;
  out[n++] = I4_DC; out[n++] = I4_H; out[n++] = I4_V; out[n++] = I4_HU;
  if (tr) { out[n++] = I4_DDL; out[n++] = I4_VL; }
  if (tl) { out[n++] = I4_DDR; out[n++] = I4_VR; out[n++] = I4_HD; }
  return n;
}
// availability nibble of 4x4 block `blk` (coding order) given the MB's neighbour availability;
// generates the reference's table g_kiNeighborIntraToI4x4 (svc_base_layer_md.cpp:223-240).
MBK_HD int i4_avail(int nb, int blk) {
  const int bx = (blk & 1) | ((blk >> 1) & 2), by = ((blk >> 1) & 1) | ((blk >> 2) & 2);
  const bool L = nb & NB_LEFT, T = nb & NB_TOP, TL = nb & NB_TOPLEFT, TR = nb & NB_TOPRIGHT;
  const bool l = bx > 0 || L;
  const bool t = by > 0 || T;
  const bool tl = (bx > 0 && by > 0) ? true : (bx > 0 ? T : (by > 0 ? L : TL));
  // top-right: row 0 looks into the top / top-right MB; inside the MB it exists iff the block to the
  // upper right has already been coded (raster rows 1..3: 1010 / 1110 / 1010)
  const bool tr = by == 0 ? (bx < 3 ? T : TR) : ((0x5750 >> (by * 4 + bx)) & 1) != 0;
  return (l ? 1 : 0) | (t ? 2 : 0) | (tl ? 4 : 0) | (tr ? 8 : 0);
}

}  // namespace mbk
