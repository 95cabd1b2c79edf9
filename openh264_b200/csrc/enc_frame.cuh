// enc_frame.cuh — macroblock entry point shared by the CUDA wavefront kernel (enc_kernels.cu) and the
// host emulation build (tests/emu).  One call = one macroblock = one warp.
#pragma once
#include "enc_mb.cuh"
#ifdef B2H264_WITH_INTER
#include "enc_inter.cuh"
#endif

namespace mbk {

MBK_HD void encode_one_mb(const EncFrameParams& p, const EncFramePtrs& f, MbScratch& s, int mbx, int mby) {
  MbCtx c;
  c.p = p; c.f = f; c.mbx = mbx; c.mby = mby;
  c.nb = (mbx > 0 ? NB_LEFT : 0) | (mby > 0 ? NB_TOP : 0) | (mbx > 0 && mby > 0 ? NB_TOPLEFT : 0) |
         (mby > 0 && mbx < p.mb_w - 1 ? NB_TOPRIGHT : 0);
  c.qp = p.qp;
  c.qp_c = tbl_chroma_qp(p.qp);          // chroma_qp_index_offset = 0
  c.lambda = tbl_lambda(p.qp);
  phase_mark(s, 31);
  // clear the staged records (levels need no clearing: every block that gets written to the bitstream is
  // produced by this MB's own coding path; only the header part must be deterministic)
  {
    uint32_t* a = reinterpret_cast<uint32_t*>(&s.info);
    for (int i = lane_id(); i < (int)(sizeof(MbInfo) / 4); i += MBK_WS) a[i] = 0;
    uint32_t* b = reinterpret_cast<uint32_t*>(&s.out);
    for (int i = lane_id(); i < MBOUT_HEADER_WORDS; i += MBK_WS) b[i] = 0;
  }
  mb_load_all(c, s);
  phase_mark(s, 0);
  if (p.is_idr) {
    intra_mb_md_enc(c, s, 0x7fffffff);
    if (lane_id() == 0) { s.info.ref_idx = -1; c.f.sad_cost[mby * p.mb_w + mbx] = 0; }   // pSadCost[0] = 0 (:2038)
  }
#ifdef B2H264_WITH_INTER
  else {
    inter_mb_md_enc(c, s);
  }
#endif
  warp_sync();
  phase_mark(s, 10);
  mb_store_recon(c, s);
  mb_publish(c, s);
  phase_mark(s, 11);
}

}  // namespace mbk
