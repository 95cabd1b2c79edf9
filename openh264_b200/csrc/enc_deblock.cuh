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
// motion of 4x4 block `blk` (raster) of a decoded macroblock for the filter: reference picture (slot, -1: list not used) and
// vector per list.  List 0 lives in MbInfo (dec_mb.cuh), list 1 of a B macroblock in its DecMbAuxB record (vectors in coding order).
MBK_HD void dbk_block_motion(const MbInfo* m, const DecMbAuxB* b, int blk, int r[2], int mv[2][2]) {
  r[0] = m->i4_mode[blk];
  mv[0][0] = r[0] < 0 ? 0 : m->mv[blk][0]; mv[0][1] = r[0] < 0 ? 0 : m->mv[blk][1];
  r[1] = -1; mv[1][0] = mv[1][1] = 0;
  if (b) {
    const int bx = blk & 3, by = blk >> 2, z = ((by >> 1) * 2 + (bx >> 1)) * 4 + (by & 1) * 2 + (bx & 1);
    r[1] = b->ref_idx[z >> 2];
    if (r[1] >= 0) { mv[1][0] = b->mv[z][0]; mv[1][1] = b->mv[z][1]; }
  }
}
MBK_HD bool dbk_mv_far(const int a[2], const int b[2]) { return iabs(a[0] - b[0]) >= 4 || iabs(a[1] - b[1]) >= 4; }

template <bool BSL = true>
MBK_HD void edge_bs(const MbInfo* cur, const MbInfo* nb /*other MB for edge 0, else == cur*/, int dir, int edge, int bs[4],
                    bool ref_ids = false /* decoder: i4_mode holds the reference picture of every 4x4 block */,
                    const DecMbAuxB* cur_b = nullptr, const DecMbAuxB* nb_b = nullptr /* list 1 of a B macroblock, else NULL */) {
  const bool mb_edge = edge == 0;
  for (int i = 0; i < 4; i++) {
    // q block in cur, p block in nb (raster 4x4 indices)
    const int q = dir == 0 ? i * 4 + edge : edge * 4 + i;
    const int p = dir == 0 ? (mb_edge ? i * 4 + 3 : q - 1) : (mb_edge ? 12 + i : q - 4);
    if (MBT_IS_INTRA(cur->mb_type) || MBT_IS_INTRA(nb->mb_type)) { bs[i] = mb_edge ? 4 : 3; continue; }
    if (!mb_edge && (cur->mb_type == MBT_PSKIP || (BSL && cur->mb_type == MBT_BSKIP))) { bs[i] = 0; continue; }
    if (cur->nnz[q] | nb->nnz[p]) { bs[i] = 2; continue; }
    if (BSL && (cur_b || nb_b)) {
      // a B macroblock on either side (DeblockingBSliceBsMarginalMBAvcbase / IN_SMB_EDGE_MV, deblocking.cpp:544, :95): strength 1
      // unless both blocks use the same set of reference pictures and their vectors — matched by picture — are close
      int rq[2], rp[2], mq[2][2], mp[2][2];
      dbk_block_motion(cur, cur_b, q, rq, mq);
      dbk_block_motion(nb, nb_b, p, rp, mp);
      int v = 1;
      if ((rq[0] == rp[0] && rq[1] == rp[1]) || (rq[0] == rp[1] && rq[1] == rp[0])) {
        if (rq[0] != rq[1]) v = rq[0] == rp[0] ? (dbk_mv_far(mq[0], mp[0]) || dbk_mv_far(mq[1], mp[1])) : (dbk_mv_far(mq[0], mp[1]) || dbk_mv_far(mq[1], mp[0]));
        else v = (dbk_mv_far(mq[0], mp[0]) || dbk_mv_far(mq[1], mp[1])) && (dbk_mv_far(mq[0], mp[1]) || dbk_mv_far(mq[1], mp[0]));
      }
      bs[i] = v;
      continue;
    }
    if (ref_ids && cur->i4_mode[q] != nb->i4_mode[p]) { bs[i] = 1; continue; }      // different reference pictures
    const int dx = cur->mv[q][0] - nb->mv[p][0], dy = cur->mv[q][1] - nb->mv[p][1];
    bs[i] = (iabs(dx) >= 4 || iabs(dy) >= 4) ? 1 : 0;
  }
}

// Working set of one macroblock's deblocking (shared memory on the device): the macroblock's samples plus the 4
// columns / rows of the left / top macroblock its edge-0 filters read and modify, and the three MbInfo records.
// Staging turns 8 dependent read-modify-write round trips to L2 per macroblock into one load phase, the edge
// filters in shared memory, one store phase: the kernel is a dependency wavefront, its time is the per-MB latency.
enum { DBK_PY = 32, DBK_PC = 16 };
struct alignas(16) DbkTile {
  uint8_t y[20 * DBK_PY];          // rows -4..15, cols -4..15 (sample (0,0) at y[4 * DBK_PY + 4])
  uint8_t c[2][12 * DBK_PC];       // rows -4..7, cols -4..7
  MbInfo m[3];                     // cur, left, top
};
struct alignas(16) DbkTileB : DbkTile {
  DecMbAuxB b[3];                  // decoder, pictures with B slices: list 1 of those of the three that are B macroblocks
};

MBK_HD uint32_t dbk_ld32(const uint8_t* p) {
#ifdef __CUDA_ARCH__
  return __ldcg(reinterpret_cast<const uint32_t*>(p));   // written by other warps of this launch
#else
  return *reinterpret_cast<const uint32_t*>(p);
#endif
}

// BSL: the picture may hold B macroblocks or macroblocks with the 8x8 transform (decoder only: Main / High tools).  The encoder and B-free decoder pictures run the BSL = false instantiation,
// whose code and shared-memory footprint are those of the single-list filter.
template <bool BSL, typename Tile>
MBK_HD void deblock_one_mb_t(const EncFrameParams& p, const EncFramePtrs& f, int mbx, int mby, Tile& t) {
  const int idx = mby * p.mb_w + mbx;
  uint8_t* y = f.rec[0] + (size_t)(mby * 16) * p.rec_stride_y + mbx * 16;
  uint8_t* u = f.rec[1] + (size_t)(mby * 8) * p.rec_stride_c + mbx * 8;
  uint8_t* v = f.rec[2] + (size_t)(mby * 8) * p.rec_stride_c + mbx * 8;
  const int sy = p.rec_stride_y, sc = p.rec_stride_c;
  // ---- load phase: 100 luma words, 2 x 36 chroma words, 3 x 30 record words; all loads independent ----
  warp_sync();
  for (int i = lane_id(); i < 100; i += MBK_WS) {
    const int r = i / 5, w = i - r * 5;
    *reinterpret_cast<uint32_t*>(t.y + r * DBK_PY + 4 * w) = dbk_ld32(y + (ptrdiff_t)(r - 4) * sy + 4 * w - 4);
  }
  for (int i = lane_id(); i < 72; i += MBK_WS) {
    const int pl = i / 36, j = i - pl * 36, r = j / 3, w = j - r * 3;
    *reinterpret_cast<uint32_t*>(t.c[pl] + r * DBK_PC + 4 * w) = dbk_ld32((pl ? v : u) + (ptrdiff_t)(r - 4) * sc + 4 * w - 4);
  }
  {
    constexpr int kW = (int)(sizeof(MbInfo) / 4);
    const int nidx[3] = {idx, mbx > 0 ? idx - 1 : idx, mby > 0 ? idx - p.mb_w : idx};
    for (int i = lane_id(); i < 3 * kW; i += MBK_WS) {
      const int k = i / kW, w = i - k * kW;
      // through L2 like the samples: the records of macroblocks that are still being coded share cache lines with these
      reinterpret_cast<uint32_t*>(&t.m[k])[w] = dbk_ld32(reinterpret_cast<const uint8_t*>(f.mbi + nidx[k]) + 4 * w);
    }
  }
  warp_sync();
  [[maybe_unused]] const DecMbAuxB* lb[3] = {nullptr, nullptr, nullptr};
  if constexpr (BSL) if (f.dec_aux_b) {                                         // a picture with B slices: list 1 of the B macroblocks among the three
    constexpr int kWb = (int)(sizeof(DecMbAuxB) / 4);
    const int nidx[3] = {idx, mbx > 0 ? idx - 1 : idx, mby > 0 ? idx - p.mb_w : idx};
    for (int k = 0; k < 3; k++) {
      if (!MBT_IS_B(t.m[k].mb_type)) continue;
      for (int i = lane_id(); i < kWb; i += MBK_WS)
        reinterpret_cast<uint32_t*>(&t.b[k])[i] = reinterpret_cast<const uint32_t*>(f.dec_aux_b + nidx[k])[i];
      lb[k] = &t.b[k];
    }
    warp_sync();
  }
  const MbInfo* cur = &t.m[0];
  uint8_t* ty = t.y + 4 * DBK_PY + 4;
  // decoder: per-slice control travels in MbInfo::p16x16_mv (dec_mb.cuh): disable_deblocking_filter_idc, FilterOffsetA / B and the
  // slice number (idc 2: edges between slices stay unfiltered).  The encoder path codes one slice: its control is per picture.
  int idc = p.dbk_idc, off_a = p.dbk_off_a, off_b = p.dbk_off_b;
  if (p.dec_mode) {
    const int v = (uint16_t)cur->p16x16_mv[1];
    idc = v & 3; off_a = ((v >> 2) & 31) - 16; off_b = ((v >> 7) & 31) - 16;
  }
  for (int dir = 0; dir < 2 && idc != 1; dir++) {
    bool have_nb = dir == 0 ? mbx > 0 : mby > 0;
    if (have_nb && idc == 2 && t.m[1 + dir].p16x16_mv[0] != cur->p16x16_mv[0]) have_nb = false;
    const MbInfo* nbm = dir == 0 ? &t.m[1] : &t.m[2];
    for (int edge = 0; edge < 4; edge++) {
      if (edge == 0 && !have_nb) continue;
      if constexpr (BSL) if (cur->t8x8 && (edge & 1)) continue;      // 8x8 transform: the inner 4x4 edges are no transform edges (chroma has none there)
      const MbInfo* other = edge == 0 ? nbm : cur;
      int bs[4];
      if constexpr (BSL) edge_bs<true>(cur, other, dir, edge, bs, p.dec_mode != 0, lb[0], edge == 0 ? lb[1 + dir] : lb[0]);
      else edge_bs<false>(cur, other, dir, edge, bs, p.dec_mode != 0);
      if ((bs[0] | bs[1] | bs[2] | bs[3]) == 0) continue;
      const int qp_y = edge == 0 ? (cur->qp + other->qp + 1) >> 1 : cur->qp;
      const int qp_c = edge == 0 ? (cur->qp_c + other->qp_c + 1) >> 1 : cur->qp_c;
      // lanes 0..15: one luma line each; lanes 16..31 (device) / the same loop (host): the 8 + 8 chroma lines
      for (int l = lane_id(); l < 32; l += MBK_WS) {
        if (l < 16) {
          const int ia = clip3(qp_y + off_a, 0, 51), ib = clip3(qp_y + off_b, 0, 51);
          const int a = tbl_alpha(ia), b = tbl_beta(ib);
          if (a | b) {
            uint8_t* px = ty + (dir == 0 ? 4 * edge + l * DBK_PY : 4 * edge * DBK_PY + l);
            const int sx = dir == 0 ? 1 : DBK_PY;
            if (bs[0] == 4) deblock_luma_eq4_line<false>(px, sx, a, b);
            else deblock_luma_lt4_line<false>(px, sx, a, b, tbl_tc0(ia, bs[l >> 2]));
          }
        } else if (!(edge & 1)) {                                  // chroma: edges 0 and 2 only
          const int ia = clip3(qp_c + off_a, 0, 51), ib = clip3(qp_c + off_b, 0, 51);
          const int a = tbl_alpha(ia), b = tbl_beta(ib);
          if (a | b) {
            const int k = l - 16, ln = k & 7;
            uint8_t* px = t.c[k >> 3] + 4 * DBK_PC + 4 + (dir == 0 ? 2 * edge + ln * DBK_PC : 2 * edge * DBK_PC + ln);
            const int sx = dir == 0 ? 1 : DBK_PC;
            if (bs[0] == 4) deblock_chroma_eq4_line<false>(px, sx, a, b);
            else deblock_chroma_lt4_line<false>(px, sx, a, b, tbl_tc0(ia, bs[ln >> 1]) + 1);
          }
        }
      }
      warp_sync();
    }
  }
  // ---- store phase: the macroblock, the 4 columns to its left (if any), the 3 luma rows / 1 chroma row above ----
  const int w0 = mbx > 0 ? 0 : 1;
  for (int i = lane_id(); i < 80; i += MBK_WS) {
    const int r = i / 5, w = i - r * 5;
    if (w >= w0) *reinterpret_cast<uint32_t*>(y + (ptrdiff_t)r * sy + 4 * w - 4) = *reinterpret_cast<const uint32_t*>(t.y + (r + 4) * DBK_PY + 4 * w);
  }
  for (int i = lane_id(); i < 48; i += MBK_WS) {
    const int pl = i / 24, j = i - pl * 24, r = j / 3, w = j - r * 3;
    if (w >= w0) *reinterpret_cast<uint32_t*>((pl ? v : u) + (ptrdiff_t)r * sc + 4 * w - 4) = *reinterpret_cast<const uint32_t*>(t.c[pl] + (r + 4) * DBK_PC + 4 * w);
  }
  if (mby > 0) {
    for (int i = lane_id(); i < 12 + 4; i += MBK_WS) {
      if (i < 12) {
        const int r = i / 4 - 3, w = i & 3;
        *reinterpret_cast<uint32_t*>(y + (ptrdiff_t)r * sy + 4 * w) = *reinterpret_cast<const uint32_t*>(t.y + (r + 4) * DBK_PY + 4 + 4 * w);
      } else {
        const int pl = (i - 12) >> 1, w = (i - 12) & 1;
        *reinterpret_cast<uint32_t*>((pl ? v : u) - (ptrdiff_t)sc + 4 * w) = *reinterpret_cast<const uint32_t*>(t.c[pl] + 3 * DBK_PC + 4 + 4 * w);
      }
    }
  }
  warp_sync();
}
MBK_HD void deblock_one_mb(const EncFrameParams& p, const EncFramePtrs& f, int mbx, int mby, DbkTile& t) { deblock_one_mb_t<false>(p, f, mbx, mby, t); }
MBK_HD void deblock_one_mb_b(const EncFrameParams& p, const EncFramePtrs& f, int mbx, int mby, DbkTileB& t) { deblock_one_mb_t<true>(p, f, mbx, mby, t); }

}  // namespace mbk
