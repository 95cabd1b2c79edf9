// enc_deblock.cuh — in-loop deblocking of one macroblock by one warp (H.264 8.7).
// Restates DeblockingMbAvcbase / DeblockingBSCalc_c / DeblockingInterMb / DeblockingIntraMb
// (codec/encoder/core/src/deblocking.cpp:357-655) on top of the edge filters of mbk_deblock.cuh.
// Order inside an MB: luma vertical edges 0..3, luma horizontal edges 0..3 (chroma: edges 0 and 2); MBs in
// raster order (the device runs them as a wavefront with the same left / top / top-right dependencies).
#pragma once
#include "enc_types.h"
#include "mbk_deblock.cuh"

namespace mbk {

#ifdef __CUDACC__
extern __constant__ uint8_t c_alpha[52];
extern __constant__ uint8_t c_beta[52];
extern __constant__ uint8_t c_tc0[52][3];
#endif
extern uint8_t h_alpha[52];
extern uint8_t h_beta[52];
extern uint8_t h_tc0[52][3];
MBK_HD int tbl_alpha(int i) { return MBK_TBL(c_alpha, h_alpha)[i]; }
MBK_HD int tbl_beta(int i) { return MBK_TBL(c_beta, h_beta)[i]; }
MBK_HD int tbl_tc0(int i, int bs) { return bs == 0 ? -1 : MBK_TBL(c_tc0, h_tc0)[i][bs - 1]; }

// boundary strength of the 4 segments of one edge; dir 0 = vertical edge (left neighbour), 1 = horizontal
MBK_HD void edge_bs(const MbInfo* cur, const MbInfo* nb /*other MB for edge 0, else == cur*/, int dir, int edge, int bs[4]) {
  const bool mb_edge = edge == 0;
  for (int i = 0; i < 4; i++) {
    // q block in cur, p block in nb (raster 4x4 indices)
    const int q = dir == 0 ? i * 4 + edge : edge * 4 + i;
    const int p = dir == 0 ? (mb_edge ? i * 4 + 3 : q - 1) : (mb_edge ? 12 + i : q - 4);
    if (MBT_IS_INTRA(cur->mb_type) || MBT_IS_INTRA(nb->mb_type)) { bs[i] = mb_edge ? 4 : 3; continue; }
    if (!mb_edge && cur->mb_type == MBT_PSKIP) { bs[i] = 0; continue; }
    if (cur->nnz[q] | nb->nnz[p]) { bs[i] = 2; continue; }
    const int dx = cur->mv[q][0] - nb->mv[p][0], dy = cur->mv[q][1] - nb->mv[p][1];
    bs[i] = (iabs(dx) >= 4 || iabs(dy) >= 4) ? 1 : 0;
  }
}

MBK_HD void deblock_one_mb(const EncFrameParams& p, const EncFramePtrs& f, int mbx, int mby) {
  const int idx = mby * p.mb_w + mbx;
  const MbInfo* cur = f.mbi + idx;
  uint8_t* y = f.rec[0] + (size_t)(mby * 16) * p.rec_stride_y + mbx * 16;
  uint8_t* u = f.rec[1] + (size_t)(mby * 8) * p.rec_stride_c + mbx * 8;
  uint8_t* v = f.rec[2] + (size_t)(mby * 8) * p.rec_stride_c + mbx * 8;
  for (int dir = 0; dir < 2; dir++) {
    const bool have_nb = dir == 0 ? mbx > 0 : mby > 0;
    const MbInfo* nbm = dir == 0 ? cur - 1 : cur - p.mb_w;
    for (int edge = 0; edge < 4; edge++) {
      if (edge == 0 && !have_nb) continue;
      const MbInfo* other = edge == 0 ? nbm : cur;
      int bs[4];
      edge_bs(cur, other, dir, edge, bs);
      if ((bs[0] | bs[1] | bs[2] | bs[3]) == 0) continue;
      const int qp_y = edge == 0 ? (cur->qp + other->qp + 1) >> 1 : cur->qp;
      const int qp_c = edge == 0 ? (cur->qp_c + other->qp_c + 1) >> 1 : cur->qp_c;
      const int sx_y = dir == 0 ? 1 : p.rec_stride_y, sy_y = dir == 0 ? p.rec_stride_y : 1;
      // luma: 16 lines, one lane each
      {
        const int a = tbl_alpha(qp_y), b = tbl_beta(qp_y);      // slice alpha/beta offsets are 0
        if (a | b) {
          uint8_t* e = y + (dir == 0 ? 4 * edge : 4 * edge * p.rec_stride_y);
          for (int l = lane_id(); l < 16; l += MBK_WS) {
            uint8_t* px = e + l * sy_y;
            if (bs[0] == 4) deblock_luma_eq4_line(px, sx_y, a, b);
            else deblock_luma_lt4_line(px, sx_y, a, b, tbl_tc0(qp_y, bs[l >> 2]));
          }
        }
      }
      // chroma: edges 0 and 2 only, 8 lines per plane
      if (!(edge & 1)) {
        const int a = tbl_alpha(qp_c), b = tbl_beta(qp_c);
        if (a | b) {
          const int sx_c = dir == 0 ? 1 : p.rec_stride_c, sy_c = dir == 0 ? p.rec_stride_c : 1;
          const int off = dir == 0 ? 2 * edge : 2 * edge * p.rec_stride_c;
          for (int l = lane_id(); l < 16; l += MBK_WS) {
            uint8_t* px = (l < 8 ? u : v) + off + (l & 7) * sy_c;
            if (bs[0] == 4) deblock_chroma_eq4_line(px, sx_c, a, b);
            else deblock_chroma_lt4_line(px, sx_c, a, b, tbl_tc0(qp_c, bs[(l & 7) >> 1]) + 1);
          }
        }
      }
      warp_sync();
    }
  }
}

}  // namespace mbk
