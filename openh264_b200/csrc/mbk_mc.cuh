// mbk_mc.cuh — warp-cooperative H.264 sub-pel interpolation.
// Replaces (semantics of) McLuma_c / McHorVer{20,02,22}_c / the 12 averaging combos / PixelAvg_c /
// McChroma_c (codec/common/src/mc.cpp:150-380).
#pragma once
#include "mbk_common.cuh"

namespace mbk {

// 6-tap (1,-5,20,20,-5,1) centred between p[0] and p[step]
MBK_HD int tap6(const uint8_t* p, int step) {
  return (int)(p[-2 * step] + p[3 * step]) - 5 * (int)(p[-step] + p[2 * step]) + 20 * (int)(p[0] + p[step]);
}
MBK_HD int half_h(const uint8_t* p) { return clip255((tap6(p, 1) + 16) >> 5); }
MBK_HD int half_v(const uint8_t* p, int s) { return clip255((tap6(p, s) + 16) >> 5); }
// centre sample: vertical pass unrounded (fits int16, as mc.cpp:218), then horizontal pass, (x+512)>>10
MBK_HD int half_c(const uint8_t* p, int s) {
  const int c0 = tap6(p - 2, s), c1 = tap6(p - 1, s), c2 = tap6(p, s), c3 = tap6(p + 1, s), c4 = tap6(p + 2, s),
            c5 = tap6(p + 3, s);
  return clip255(((c0 + c5) - 5 * (c1 + c4) + 20 * (c2 + c3) + 512) >> 10);
}

// one luma sample at quarter-sample phase (fx, fy); p points at the integer sample
MBK_HD int luma_qpel_sample(const uint8_t* p, int s, int fx, int fy) {
  if ((fx | fy) == 0) return p[0];
  if (fy == 0) {
    const int b = half_h(p);
    return fx == 2 ? b : (p[fx == 3] + b + 1) >> 1;
  }
  if (fx == 0) {
    const int h = half_v(p, s);
    return fy == 2 ? h : (p[fy == 3 ? s : 0] + h + 1) >> 1;
  }
  if (fx == 2 && fy == 2) return half_c(p, s);
  if (fx == 2) return (half_h(p + (fy == 3 ? s : 0)) + half_c(p, s) + 1) >> 1;
  if (fy == 2) return (half_v(p + (fx == 3), s) + half_c(p, s) + 1) >> 1;
  return (half_h(p + (fy == 3 ? s : 0)) + half_v(p + (fx == 3), s) + 1) >> 1;
}

// w x h luma prediction; src already offset by the integer part of the MV (as McLuma_c expects)
MBK_STAGE void warp_mc_luma(const uint8_t* src, int ss, uint8_t* dst, int ds, int mvx, int mvy, int w,
                                             int h) {
  const int fx = mvx & 3, fy = mvy & 3;
  const int sh = 31 - clz32((uint32_t)w);
  const bool p2 = (w & (w - 1)) == 0;
  if ((fx | fy) == 0 && p2 && w >= 4) {           // integer vector: word copy (one unaligned 4-byte load per group)
    const int gsh = sh - 2;
    for (int g = lane_id(); g < (w * h) >> 2; g += MBK_WS) {
      const int y = g >> gsh, x = (g & ((1 << gsh) - 1)) << 2;
      const uint32_t v = ld4u(src + y * ss + x);
      uint8_t* d = dst + y * ds + x;
      if (((reinterpret_cast<uintptr_t>(dst) | (uintptr_t)ds) & 3) == 0) *reinterpret_cast<uint32_t*>(d) = v;     // every caller's case
      else { d[0] = (uint8_t)v; d[1] = (uint8_t)(v >> 8); d[2] = (uint8_t)(v >> 16); d[3] = (uint8_t)(v >> 24); }
    }
    return;
  }
  for (int i = lane_id(); i < w * h; i += MBK_WS) {
    const int y = p2 ? i >> sh : i / w, x = i - y * w;
    dst[y * ds + x] = (uint8_t)luma_qpel_sample(src + y * ss + x, ss, fx, fy);
  }
}

// half-sample planes used by the fractional refinement (pfLumaHalfpelHor/Ver/Cen, mc.h:46-49)
MBK_FN void warp_halfpel(int which, const uint8_t* src, int ss, uint8_t* dst, int ds, int w, int h) {
  for (int i = lane_id(); i < w * h; i += MBK_WS) {
    const int y = i / w, x = i - y * w;
    const uint8_t* p = src + y * ss + x;
    dst[y * ds + x] = (uint8_t)(which == 0 ? half_h(p) : which == 1 ? half_v(p, ss) : half_c(p, ss));
  }
}

// bilinear eighth-sample chroma (mc.cpp:349-380)
MBK_FN void warp_mc_chroma(const uint8_t* src, int ss, uint8_t* dst, int ds, int mvx, int mvy, int w,
                                               int h) {
  const int dx = mvx & 7, dy = mvy & 7;
  const int A = (8 - dx) * (8 - dy), B = dx * (8 - dy), Cc = (8 - dx) * dy, D = dx * dy;
  const int sh = 31 - clz32((uint32_t)w);        // chroma widths are 2, 4, 8
  for (int i = lane_id(); i < w * h; i += MBK_WS) {
    const int y = i >> sh, x = i - (y << sh);
    const uint8_t* p = src + y * ss + x;
    int v;
    if ((dx | dy) == 0) v = p[0];
    else v = (A * p[0] + B * p[1] + Cc * p[ss] + D * p[ss + 1] + 32) >> 6;
    dst[y * ds + x] = (uint8_t)v;
  }
}

MBK_FN void warp_pixel_avg(uint8_t* dst, int ds, const uint8_t* a, int sa, const uint8_t* b, int sb,
                                               int w, int h) {
  for (int i = lane_id(); i < w * h; i += MBK_WS) {
    const int y = i / w, x = i - y * w;
    dst[y * ds + x] = (uint8_t)((a[y * sa + x] + b[y * sb + x] + 1) >> 1);
  }
}

}  // namespace mbk
