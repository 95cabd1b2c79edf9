// mbk_me.cuh — warp-cooperative integer-pel motion search for one block.
// Replaces (semantics of) WelsMotionEstimateSearch / WelsMotionEstimateInitialPoint /
// WelsDiamondSearch / WelsMeSadCostSelect / CalculateSatdCost
// (codec/encoder/core/src/svc_motion_estimate.cpp:170-181, 222-284, 300-386, 286-291).
// The decision order (candidate order, strict '<', up/down/left/right, <=16 iterations, the
// out-of-range "spin") is observable in the bitstream and is reproduced exactly.
#pragma once
#include "mbk_common.cuh"
#include "mbk_sad.cuh"

namespace mbk {

struct MeIn {
  const uint8_t* enc; int enc_stride;      // current block
  const uint8_t* ref; int ref_stride;      // co-located block in the (padded) reference plane
  int blk;                                 // BLK_*
  int mvp_x, mvp_y;                        // predictor, quarter-pel
  int min_x, min_y, max_x, max_y;          // integer-pel window (sMvStartMin/Max)
  int n_mvc; const int16_t* mvc;           // candidates (x,y pairs, quarter-pel), readable by all lanes
  uint32_t sad_pred;                       // early-stop threshold
  int lambda;                              // g_kiQpCostTable[qp]
  bool calc_satd;
  // optional copy of a part of the reference plane in shared memory (stride win_w): sample (0,0) of the copy is
  // the plane sample at integer offset (win_dx, win_dy) from `ref`; null = none.  Same bytes as the plane, so
  // which of the two a SAD reads cannot change a result.
  const uint8_t* win; int win_w, win_h, win_dx, win_dy;
};
struct MeOut {
  int mv_x, mv_y;                          // quarter-pel
  uint32_t sad_cost, satd_cost;
  const uint8_t* ref_best;                 // matched block in the reference plane
};

// the w x h block at integer offset (mx, my) from the co-located block, readable with `m` samples of margin all
// round: out of the shared-memory copy when it covers that, else out of the plane
MBK_HD const uint8_t* me_block(const MeIn& in, int mx, int my, int w, int h, int m, int* stride) {
  if (in.win) {
    const int x = mx - in.win_dx, y = my - in.win_dy;
    if (x >= m && y >= m && x + w + m <= in.win_w && y + h + m <= in.win_h) { *stride = in.win_w; return in.win + y * in.win_w + x; }
  }
  *stride = in.ref_stride;
  return in.ref + my * in.ref_stride + mx;
}

#ifdef B2H264_ME_CALL                      // profiling variant: the search as a real call
#define MBK_ME MBK_FN
#else
#define MBK_ME MBK_STAGE                // one call site per kernel (me_partition / k_me_search)
#endif
MBK_ME void warp_me_search(const MeIn& in, MeOut& out) {
  const int lw = blk_lw(in.blk), lh = blk_lh(in.blk), w = 1 << lw, h = 1 << lh;
  const int px = in.mvp_x, py = in.mvp_y;
  int rs;

  int mx = clip3((2 + px) >> 2, in.min_x, in.max_x);
  int my = clip3((2 + py) >> 2, in.min_y, in.max_y);
  const uint8_t* r = me_block(in, mx, my, w, h, 0, &rs);
  int best = warp_sad(in.enc, in.enc_stride, r, rs, lw, lh) + mvd_cost(in.lambda, mx * 4 - px, my * 4 - py);
  for (int i = 0; i < in.n_mvc; i++) {
    const int cx = clip3((2 + in.mvc[2 * i]) >> 2, in.min_x, in.max_x);
    const int cy = clip3((2 + in.mvc[2 * i + 1]) >> 2, in.min_y, in.max_y);
    if (cx == mx && cy == my) continue;
    r = me_block(in, cx, cy, w, h, 0, &rs);
    const int c = warp_sad(in.enc, in.enc_stride, r, rs, lw, lh) + mvd_cost(in.lambda, cx * 4 - px, cy * 4 - py);
    if (c < best) { best = c; mx = cx; my = cy; }
  }
  if (!(best < (int)in.sad_pred)) {
    int dx = mx * 4 - px, dy = my * 4 - py;
    for (int iter = 0; iter < 16; iter++) {
      const int tx = (dx + px) >> 2, ty = (dy + py) >> 2;
      if (!(tx >= in.min_x && tx < in.max_x && ty >= in.min_y && ty < in.max_y)) continue;
      int s[4];
      r = me_block(in, tx, ty, w, h, 1, &rs);
      warp_sad_four(in.enc, in.enc_stride, r, rs, lw, lh, s);
      const int cu = s[0] + mvd_cost(in.lambda, dx, dy - 4), cd = s[1] + mvd_cost(in.lambda, dx, dy + 4);
      const int cl = s[2] + mvd_cost(in.lambda, dx - 4, dy), cr = s[3] + mvd_cost(in.lambda, dx + 4, dy);
      int sx = 0, sy = 0;
      bool moved = false;
      if (cu < best) { best = cu; sx = 0; sy = -1; moved = true; }
      if (cd < best) { best = cd; sx = 0; sy = 1; moved = true; }
      if (cl < best) { best = cl; sx = -1; sy = 0; moved = true; }
      if (cr < best) { best = cr; sx = 1; sy = 0; moved = true; }
      if (!moved) break;
      dx += 4 * sx; dy += 4 * sy;
    }
    mx = (dx + px) >> 2; my = (dy + py) >> 2;
  }
  out.mv_x = mx * 4; out.mv_y = my * 4;
  out.sad_cost = (uint32_t)best;
  out.satd_cost = (uint32_t)best;
  out.ref_best = in.ref + my * in.ref_stride + mx;
  if (in.calc_satd) {
    r = me_block(in, mx, my, w, h, 0, &rs);
    out.satd_cost = (uint32_t)(warp_satd(in.enc, in.enc_stride, r, rs, lw, lh) +
                               mvd_cost(in.lambda, out.mv_x - px, out.mv_y - py));
  }
}

}  // namespace mbk
