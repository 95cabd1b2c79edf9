// mbk_deblock.cuh — H.264 in-loop deblocking edge filters, one thread per pixel line across the edge.
// Replaces (semantics of) DeblockLumaLt4_c / DeblockLumaEq4_c / DeblockChromaLt4_c / DeblockChromaEq4_c
// (codec/common/src/deblocking_common.cpp:5-168).  sx = step across the edge.
#pragma once
#include "mbk_common.cuh"

namespace mbk {

// picture samples may have been written by the warp of the previous MB row on another SM during the same
// launch: read them around L1 on the device
// CG = false: the samples sit in a staged tile (shared memory / host memory): plain loads
template <bool CG>
MBK_HD int ldpix(const uint8_t* p) {
#ifdef __CUDA_ARCH__
  if (CG) return __ldcg(p);
#endif
  return *p;
}

// bS < 4 luma, one line; tc0 < 0 means "not filtered"
template <bool CG = true>
MBK_HD void deblock_luma_lt4_line(uint8_t* pix, int sx, int alpha, int beta, int tc0) {
  if (tc0 < 0) return;
  const int p0 = ldpix<CG>(pix + (-sx)), p1 = ldpix<CG>(pix + (-2 * sx)), p2 = ldpix<CG>(pix + (-3 * sx)), q0 = ldpix<CG>(pix + (0)), q1 = ldpix<CG>(pix + (sx)), q2 = ldpix<CG>(pix + (2 * sx));
  if (!(iabs(p0 - q0) < alpha && iabs(p1 - p0) < beta && iabs(q1 - q0) < beta)) return;
  int tc = tc0;
  if (iabs(p2 - p0) < beta) { pix[-2 * sx] = (uint8_t)(p1 + clip3((p2 + ((p0 + q0 + 1) >> 1) - 2 * p1) >> 1, -tc0, tc0)); tc++; }
  if (iabs(q2 - q0) < beta) { pix[sx] = (uint8_t)(q1 + clip3((q2 + ((p0 + q0 + 1) >> 1) - 2 * q1) >> 1, -tc0, tc0)); tc++; }
  const int delta = clip3((4 * (q0 - p0) + (p1 - q1) + 4) >> 3, -tc, tc);
  pix[-sx] = (uint8_t)clip255(p0 + delta);
  pix[0] = (uint8_t)clip255(q0 - delta);
}
// bS == 4 luma, one line
template <bool CG = true>
MBK_HD void deblock_luma_eq4_line(uint8_t* pix, int sx, int alpha, int beta) {
  const int p0 = ldpix<CG>(pix + (-sx)), p1 = ldpix<CG>(pix + (-2 * sx)), p2 = ldpix<CG>(pix + (-3 * sx)), q0 = ldpix<CG>(pix + (0)), q1 = ldpix<CG>(pix + (sx)), q2 = ldpix<CG>(pix + (2 * sx));
  const int d = iabs(p0 - q0);
  if (!(d < alpha && iabs(p1 - p0) < beta && iabs(q1 - q0) < beta)) return;
  if (d < (alpha >> 2) + 2) {
    if (iabs(p2 - p0) < beta) {
      const int p3 = ldpix<CG>(pix + (-4 * sx));
      pix[-sx] = (uint8_t)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
      pix[-2 * sx] = (uint8_t)((p2 + p1 + p0 + q0 + 2) >> 2);
      pix[-3 * sx] = (uint8_t)((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
    } else {
      pix[-sx] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
    }
    if (iabs(q2 - q0) < beta) {
      const int q3 = ldpix<CG>(pix + (3 * sx));
      pix[0] = (uint8_t)((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
      pix[sx] = (uint8_t)((p0 + q0 + q1 + q2 + 2) >> 2);
      pix[2 * sx] = (uint8_t)((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
    } else {
      pix[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
    }
  } else {
    pix[-sx] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
    pix[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
  }
}
// bS < 4 chroma, one line of one plane; filtered only when tc > 0
template <bool CG = true>
MBK_HD void deblock_chroma_lt4_line(uint8_t* pix, int sx, int alpha, int beta, int tc) {
  if (tc <= 0) return;
  const int p0 = ldpix<CG>(pix + (-sx)), p1 = ldpix<CG>(pix + (-2 * sx)), q0 = ldpix<CG>(pix + (0)), q1 = ldpix<CG>(pix + (sx));
  if (iabs(p0 - q0) < alpha && iabs(p1 - p0) < beta && iabs(q1 - q0) < beta) {
    const int delta = clip3((4 * (q0 - p0) + (p1 - q1) + 4) >> 3, -tc, tc);
    pix[-sx] = (uint8_t)clip255(p0 + delta);
    pix[0] = (uint8_t)clip255(q0 - delta);
  }
}
template <bool CG = true>
MBK_HD void deblock_chroma_eq4_line(uint8_t* pix, int sx, int alpha, int beta) {
  const int p0 = ldpix<CG>(pix + (-sx)), p1 = ldpix<CG>(pix + (-2 * sx)), q0 = ldpix<CG>(pix + (0)), q1 = ldpix<CG>(pix + (sx));
  if (iabs(p0 - q0) < alpha && iabs(p1 - p0) < beta && iabs(q1 - q0) < beta) {
    pix[-sx] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
    pix[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
  }
}

}  // namespace mbk
