// mbk_sad.cuh — warp-cooperative SAD / 4-neighbour SAD / SATD.
// Replaces (semantics of) WelsSampleSad*_c, WelsSampleSadFour*_c (codec/common/src/sad_common.cpp:44-165)
// and WelsSampleSatd*_c (codec/encoder/core/src/sample.cpp:48-148).
#pragma once
#include "mbk_common.cuh"

namespace mbk {

// SAD of a (1<<lw) x (1<<lh) block: the block is cut into 4-pixel groups, lane g takes groups
// g, g+32, ...; each group is one __vsadu4 on packed bytes; warp total by REDUX.
MBK_HD int warp_sad(const uint8_t* a, int sa, const uint8_t* b, int sb, int lw, int lh) {
  const int lg = lw - 2;                    // log2(groups per row)
  const int ngroups = 1 << (lg + lh);
  int s = 0;
  for (int g = lane_id(); g < ngroups; g += MBK_WS) {
    const int row = g >> lg, col = (g & ((1 << lg) - 1)) << 2;
    s += vsadu4(ld4u(a + row * sa + col), ld4u(b + row * sb + col));
  }
  return warp_sum(s);
}

// SADs against b shifted up, down, left, right by one pixel (pfSample4Sad order), cur read once.
MBK_STAGE void warp_sad_four(const uint8_t* a, int sa, const uint8_t* b, int sb, int lw, int lh,
                                              int out[4]) {
  const int lg = lw - 2;
  const int ngroups = 1 << (lg + lh);
  int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  for (int g = lane_id(); g < ngroups; g += MBK_WS) {
    const int row = g >> lg, col = (g & ((1 << lg) - 1)) << 2;
    const uint32_t c = ld4u(a + row * sa + col);
    const uint8_t* r = b + row * sb + col;
    s0 += vsadu4(c, ld4u(r - sb));
    s1 += vsadu4(c, ld4u(r + sb));
    s2 += vsadu4(c, ld4u(r - 1));
    s3 += vsadu4(c, ld4u(r + 1));
  }
  out[0] = warp_sum(s0);
  out[1] = warp_sum(s1);
  out[2] = warp_sum(s2);
  out[3] = warp_sum(s3);
}

// |Hadamard4x4(a - b)| summed, (sum+1)>>1, for ONE 4x4 block, by one thread (sample.cpp:48-96).
MBK_HD int satd4x4_thread(const uint8_t* a, int sa, const uint8_t* b, int sb) {
  int t[4][4];
#pragma unroll
  for (int y = 0; y < 4; y++) {
    const uint32_t wa = ld4u(a + y * sa), wb = ld4u(b + y * sb);
    const int d0 = (int)(wa & 0xff) - (int)(wb & 0xff), d1 = (int)((wa >> 8) & 0xff) - (int)((wb >> 8) & 0xff);
    const int d2 = (int)((wa >> 16) & 0xff) - (int)((wb >> 16) & 0xff), d3 = (int)(wa >> 24) - (int)(wb >> 24);
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

// as satd4x4_thread, the prediction being the rounded average of two byte planes of stride sp
MBK_HD int satd4x4_avg_thread(const uint8_t* a, int sa, const uint8_t* p0, const uint8_t* p1, int sp) {
  int t[4][4];
#pragma unroll
  for (int y = 0; y < 4; y++) {
    const uint32_t wa = ld4u(a + y * sa), wb = vavgu4(ld4u(p0 + y * sp), ld4u(p1 + y * sp));
    const int d0 = (int)(wa & 0xff) - (int)(wb & 0xff), d1 = (int)((wa >> 8) & 0xff) - (int)((wb >> 8) & 0xff);
    const int d2 = (int)((wa >> 16) & 0xff) - (int)((wb >> 16) & 0xff), d3 = (int)(wa >> 24) - (int)(wb >> 24);
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
MBK_STAGE int warp_satd_avg(const uint8_t* a, int sa, const uint8_t* p0, const uint8_t* p1, int sp, int lw, int lh) {
  const int lbx = lw - 2;
  const int nblk = 1 << (lbx + lh - 2);
  int s = 0;
  for (int l = lane_id(); l < nblk; l += MBK_WS) {
    const int by = l >> lbx, bx = l & ((1 << lbx) - 1);
    const int o = 4 * by * sp + 4 * bx;
    s += satd4x4_avg_thread(a + 4 * by * sa + 4 * bx, sa, p0 + o, p1 + o, sp);
  }
  return warp_sum(s);
}

// SATD of a block: one lane per 4x4 sub-block (16 lanes busy for 16x16), warp total by REDUX.
MBK_HD int warp_satd_inl(const uint8_t* a, int sa, const uint8_t* b, int sb, int lw, int lh) {
  const int lbx = lw - 2;                    // log2(4x4 blocks per row)
  const int nblk = 1 << (lbx + lh - 2);
  int s = 0;
  for (int l = lane_id(); l < nblk; l += MBK_WS) {
    const int by = l >> lbx, bx = l & ((1 << lbx) - 1);
    s += satd4x4_thread(a + 4 * by * sa + 4 * bx, sa, b + 4 * by * sb + 4 * bx, sb);
  }
  return warp_sum(s);
}
// the shared (real-call) copy; the skip test of stage A, which every macroblock runs, uses the inlined form
MBK_FN int warp_satd(const uint8_t* a, int sa, const uint8_t* b, int sb, int lw, int lh) { return warp_satd_inl(a, sa, b, sb, lw, lh); }

}  // namespace mbk
