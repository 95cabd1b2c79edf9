// dec_t8x8.cuh — decoder, High profile: the 8x8 transform path of one macroblock (transform_size_8x8_flag): 8x8 zig-zag, dequantisation
// with the flat default scaling list (Rec. H.264 8.5.13: LevelScale8x8 = 16 x normAdjust8x8), the 8x8 inverse transform (8.5.13 — the same
// butterflies as IdctResAddPred8x8_c, codec/decoder/core/src/decode_mb_aux.cpp:79, idct8_1d in mbk_xform.cuh), and Intra_8x8 prediction
// with its reference sample filter (8.3.2.2; the reference: codec/decoder/core/src/get_intra_predictor.cpp WelsI8x8Luma*Pred_c,
// rec_mb.cpp RecI8x8Luma).  A lane handles one 8x8 block at a time (four blocks per macroblock); Intra_8x8 blocks depend on each
// other and run one after the other on lane 0, like the Intra_4x4 blocks of dec_mb.cuh.
#pragma once
#include "enc_mb.cuh"

namespace mbk {

MBK_HD int zigzag8_pos(int i) {
  const uint8_t zz[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                          35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
  return zz[i];
}
// 16 x normAdjust8x8(qp % 6, i, j) (Table in 8.5.9; g_kuiDequantCoeff8x8, codec/common/src/common_tables.cpp:237)
MBK_HD int level_scale8x8(int q6, int pos) {
  const uint8_t v[6][6] = {{20, 18, 32, 19, 25, 24}, {22, 19, 35, 21, 28, 26}, {26, 23, 42, 24, 33, 31}, {28, 25, 45, 26, 35, 33}, {32, 28, 51, 30, 40, 38}, {36, 32, 58, 34, 46, 43}};
  const int i = pos >> 3, j = pos & 7;
  int k;
  if ((i & 3) == 0 && (j & 3) == 0) k = 0;
  else if ((i & 1) == 1 && (j & 1) == 1) k = 1;
  else if ((i & 3) == 2 && (j & 3) == 2) k = 2;
  else if (((i & 3) == 0 && (j & 1) == 1) || ((i & 1) == 1 && (j & 3) == 0)) k = 3;
  else if (((i & 3) == 0 && (j & 3) == 2) || ((i & 3) == 2 && (j & 3) == 0)) k = 4;
  else k = 5;
  return 16 * v[q6][k];
}
// levels in 8x8 zig-zag order -> dequantised coefficients in raster order (WelsResidualBlockCavlc8x8, parse_mb_syn_cavlc.cpp:1059)
MBK_HD void unscan_dequant8x8(int16_t d[64], const int16_t lv[64], int qp) {
  const int q6 = qp % 6, sh = qp / 6;
  for (int i = 0; i < 64; i++) d[i] = 0;
  for (int i = 0; i < 64; i++) {
    const int l = lv[i];
    if (!l) continue;
    const int pos = zigzag8_pos(i), ls = level_scale8x8(q6, pos);
    d[pos] = (int16_t)(sh >= 6 ? (l * ls) * (1 << (sh - 6)) : (l * ls + (1 << (5 - sh))) >> (6 - sh));
  }
}
// rec = clip(pred + ((idct(c) + 32) >> 6)); rows first, then columns, 16-bit intermediates (IdctResAddPred8x8_c)
MBK_HD void idct8x8_rec(uint8_t* rec, int rs, const uint8_t* pred, int ps, const int16_t c[64]) {
  int16_t t[64], in[8], out[8];
  for (int r = 0; r < 8; r++) {
    for (int x = 0; x < 8; x++) in[x] = c[8 * r + x];
    idct8_1d(in, out);
    for (int x = 0; x < 8; x++) t[8 * r + x] = out[x];
  }
  for (int x = 0; x < 8; x++) {
    for (int y = 0; y < 8; y++) in[y] = t[8 * y + x];
    idct8_1d(in, out);
    for (int y = 0; y < 8; y++) rec[y * rs + x] = (uint8_t)clip255(pred[y * ps + x] + ((32 + out[y]) >> 6));
  }
}

// Intra_8x8 prediction of the block whose sample (0,0) is at org (pitch `pitch`, neighbours in place around it).
// avail: bit 0 left, bit 1 top, bit 2 top-left, bit 3 top-right.  mode 0..8 (Table 8-3); DC adapts to what is available.
MBK_HD void pred_i8x8(uint8_t pr[64], const uint8_t* org, int pitch, int mode, int avail) {
  const bool L = avail & 1, T = (avail & 2) != 0, TL = (avail & 4) != 0, TR = (avail & 8) != 0;
  int top[16], left[8], tl = 128;                                 // filtered reference samples p'[x,-1], p'[-1,y], p'[-1,-1]
  {
    int rt[16], rl[8];
    const int rtl = TL ? org[-pitch - 1] : 0;
    if (T) {
      for (int x = 0; x < 8; x++) rt[x] = org[-pitch + x];
      for (int x = 8; x < 16; x++) rt[x] = TR ? org[-pitch + x] : rt[7];
      top[0] = TL ? (rtl + 2 * rt[0] + rt[1] + 2) >> 2 : (3 * rt[0] + rt[1] + 2) >> 2;
      for (int x = 1; x < 15; x++) top[x] = (rt[x - 1] + 2 * rt[x] + rt[x + 1] + 2) >> 2;
      top[15] = (rt[14] + 3 * rt[15] + 2) >> 2;
    } else {
      for (int x = 0; x < 16; x++) { rt[x] = 0; top[x] = 0; }
    }
    if (L) {
      for (int y = 0; y < 8; y++) rl[y] = org[y * pitch - 1];
      left[0] = TL ? (rtl + 2 * rl[0] + rl[1] + 2) >> 2 : (3 * rl[0] + rl[1] + 2) >> 2;
      for (int y = 1; y < 7; y++) left[y] = (rl[y - 1] + 2 * rl[y] + rl[y + 1] + 2) >> 2;
      left[7] = (rl[6] + 3 * rl[7] + 2) >> 2;
    } else {
      for (int y = 0; y < 8; y++) { rl[y] = 0; left[y] = 0; }
    }
    if (TL) {
      if (T && L) tl = (rt[0] + 2 * rtl + rl[0] + 2) >> 2;
      else if (T) tl = (3 * rtl + rt[0] + 2) >> 2;
      else if (L) tl = (3 * rtl + rl[0] + 2) >> 2;
      else tl = rtl;
    }
  }
  // edge[] = p'[-1,7] .. p'[-1,0], p'[-1,-1], p'[0,-1] .. p'[15,-1] (index 8 = corner): the diagonal modes read it linearly
  int e[25];
  for (int y = 0; y < 8; y++) e[7 - y] = left[y];
  e[8] = tl;
  for (int x = 0; x < 16; x++) e[9 + x] = top[x];
  for (int y = 0; y < 8; y++)
    for (int x = 0; x < 8; x++) {
      int v;
      switch (mode) {
        case 0: v = top[x]; break;
        case 1: v = left[y]; break;
        case 2: {
          int sum = 0;
          if (T && L) { for (int i = 0; i < 8; i++) sum += top[i] + left[i]; v = (sum + 8) >> 4; }
          else if (L) { for (int i = 0; i < 8; i++) sum += left[i]; v = (sum + 4) >> 3; }
          else if (T) { for (int i = 0; i < 8; i++) sum += top[i]; v = (sum + 4) >> 3; }
          else v = 128;
          break;
        }
        case 3:                                                   // diagonal down left
          v = (x == 7 && y == 7) ? (top[14] + 3 * top[15] + 2) >> 2 : (top[x + y] + 2 * top[x + y + 1] + top[x + y + 2] + 2) >> 2;
          break;
        case 4: {                                                 // diagonal down right: along e[], centre 8 + x - y
          const int k = 8 + x - y;
          v = (e[k - 1] + 2 * e[k] + e[k + 1] + 2) >> 2;
          break;
        }
        case 5: {                                                 // vertical right
          const int z = 2 * x - y;
          if (z >= 0 && !(z & 1)) { const int k = 8 + x - (y >> 1); v = (e[k] + e[k + 1] + 1) >> 1; }
          else if (z >= 0) { const int k = 8 + x - (y >> 1); v = (e[k - 1] + 2 * e[k] + e[k + 1] + 2) >> 2; }
          else if (z == -1) v = (left[0] + 2 * tl + top[0] + 2) >> 2;
          else { const int k = 8 - (y - 2 * x); v = (e[k] + 2 * e[k + 1] + e[k + 2] + 2) >> 2; }   // p'[-1, y-2x-1], [-1, y-2x-2], [-1, y-2x-3]
          break;
        }
        case 6: {                                                 // horizontal down
          const int z = 2 * y - x;
          if (z >= 0 && !(z & 1)) { const int k = 8 - (y - (x >> 1)); v = (e[k] + e[k - 1] + 1) >> 1; }   // p'[-1, y-(x>>1)-1], p'[-1, y-(x>>1)]
          else if (z >= 0) { const int k = 8 - (y - (x >> 1)); v = (e[k + 1] + 2 * e[k] + e[k - 1] + 2) >> 2; }
          else if (z == -1) v = (left[0] + 2 * tl + top[0] + 2) >> 2;
          else { const int k = 8 + (x - 2 * y); v = (e[k] + 2 * e[k - 1] + e[k - 2] + 2) >> 2; }   // p'[x-2y-1,-1], [x-2y-2,-1], [x-2y-3,-1]
          break;
        }
        case 7:                                                   // vertical left
          v = !(y & 1) ? (top[x + (y >> 1)] + top[x + (y >> 1) + 1] + 1) >> 1
                       : (top[x + (y >> 1)] + 2 * top[x + (y >> 1) + 1] + top[x + (y >> 1) + 2] + 2) >> 2;
          break;
        default: {                                                // horizontal up
          const int z = x + 2 * y;
          if (z > 13) v = left[7];
          else if (z == 13) v = (left[6] + 3 * left[7] + 2) >> 2;
          else if (!(z & 1)) v = (left[y + (x >> 1)] + left[y + (x >> 1) + 1] + 1) >> 1;
          else v = (left[y + (x >> 1)] + 2 * left[y + (x >> 1) + 1] + left[y + (x >> 1) + 2] + 2) >> 2;
          break;
        }
      }
      pr[8 * y + x] = (uint8_t)v;
    }
}

}  // namespace mbk
