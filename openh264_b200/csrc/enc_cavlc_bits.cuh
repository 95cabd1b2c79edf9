// enc_cavlc_bits.cuh — EXACT number of CAVLC bits of one macroblock, computed on the device from the staged record
// (SURVEY.md section 8f rank 2).  Mirrors, length for length, what the host writer emits for macroblock_layer():
// h264_bitstream.cpp write_slice / write_block, i.e. the reference's WelsSpatialWriteMbSyn
// (codec/encoder/core/src/svc_set_mb_syn_cavlc.cpp:260) and WriteBlockResidualCavlc (set_mb_syn_cavlc.cpp:109).
// It is the quantity the reference's rate control reads back from the bitstream position per macroblock
// (ratectl.cpp:1239-1278) and the overflow check of svc_encode_slice.cpp:1863-1867 needs; having it on the device
// removes the entropy-in-the-loop coupling (no bits are emitted here).  mb_skip_run is NOT part of the figure (it
// belongs to the run that ends at the next coded macroblock); mb_qp_delta is counted as se(0): constant QP.
// One lane per residual block (27 blocks + 1 lane for the header syntax), one warp reduction.
#pragma once
#include "enc_mb.cuh"

namespace mbk {

// code LENGTHS of Rec. H.264 Tables 9-5, 9-7, 9-8, 9-9, 9-10 and the me(v) codeNum of Table 9-4; filled by
// b2h264_init from the same generated tables the host writer uses (cavlc_tables.h), so the two cannot drift apart
struct CavlcLen {
  uint8_t coeff_token[5][17][4];
  uint8_t total_zeros[16][16];
  uint8_t total_zeros_cdc[4][4];
  uint8_t run_before[8][15];
  uint8_t nc_class[18];
  uint8_t cbp_intra[48], cbp_inter[48];
};
extern CavlcLen h_cavlc_len;
#ifdef __CUDACC__
extern __device__ CavlcLen d_cavlc_len;
#endif
MBK_HD const CavlcLen& cavlc_len() { return MBK_TBL(d_cavlc_len, h_cavlc_len); }

MBK_HD int ue_len(uint32_t v) { return 2 * (31 - clz32(v + 1)) + 1; }
MBK_HD int se_len(int v) { return ue_len(v > 0 ? (uint32_t)(2 * v - 1) : (uint32_t)(-2 * v)); }

// bits of one residual block: lv = scan-ordered levels, max_coef 16 / 15 / 4, nc < 0 for chroma DC
MBK_HD int cavlc_block_bits(const int16_t* lv, int max_coef, int nc, bool active) {
  const CavlcLen& T = cavlc_len();
  int16_t level[16];
  uint8_t run[16];
  int total = 0, total_zeros = 0;
  int i = active ? max_coef - 1 : -1;
  while (i >= 0 && lv[i] == 0) i--;
  while (i >= 0) {
    int zeros = 0;
    level[total] = lv[i--];
    while (i >= 0 && lv[i] == 0) { zeros++; i--; }
    total_zeros += zeros;
    run[total++] = (uint8_t)zeros;
  }
  int t1 = 0;
  for (int k = 0; k < total && k < 3; k++) {
    if (level[k] == 1 || level[k] == -1) t1++;
    else break;
  }
  const int cls = nc < 0 ? 4 : T.nc_class[nc];
  int bits = T.coeff_token[cls][total][t1];
  if (total == 0) return bits;
  bits += t1;
  int suffix_len = (total > 10 && t1 < 3) ? 1 : 0;
  for (int k = t1; k < total; k++) {
    const int val = level[k];
    int code = val > 0 ? 2 * (val - 1) : -2 * val - 1;       // level_code
    if (k == t1 && t1 < 3) code -= 2;
    int prefix = code >> suffix_len, suffix_size = suffix_len;
    if (prefix >= 14 && prefix < 30 && suffix_len == 0) { prefix = 14; suffix_size = 4; }
    else if (prefix >= 15) { prefix = 15; suffix_size = 12; }
    bits += prefix + 1 + suffix_size;
    if (suffix_len == 0) suffix_len = 1;
    if (iabs(val) > (3 << (suffix_len - 1)) && suffix_len < 6) suffix_len++;
  }
  if (total < max_coef) bits += nc >= 0 ? T.total_zeros[total][total_zeros] : T.total_zeros_cdc[total][total_zeros];
  int zeros_left = total_zeros;
  for (int k = 0; k + 1 < total && zeros_left > 0; k++) {
    bits += T.run_before[zeros_left > 7 ? 7 : zeros_left][run[k]];
    zeros_left -= run[k];
  }
  return bits;
}

MBK_HD int nc_of(int a, int b) {          // a / b = neighbour counts or -1 when unavailable
  if (a >= 0 && b >= 0) return (a + b + 1) >> 1;
  if (a >= 0) return a;
  if (b >= 0) return b;
  return 0;
}

// bits of macroblock_layer() of the record staged in s.out / s.info (after mb_publish filled mb_type / cbp / nnz);
// 0 for a P_SKIP macroblock.  is_idr selects the mb_type numbering of I slices.
MBK_FN int mb_cavlc_bits(const MbCtx& c, MbScratch& s) {
  const MbOut& m = s.out;
  if (m.mb_type == MBT_PSKIP) return 0;
  const CavlcLen& T = cavlc_len();
  const int cbp_l = m.cbp & 15, cbp_c = m.cbp >> 4;
  const bool coded = m.cbp > 0 || m.mb_type == MBT_I16x16;
  const int8_t* L = (c.nb & NB_LEFT) ? s.nbi[3].nnz : nullptr;
  const int8_t* Tn = (c.nb & NB_TOP) ? s.nbi[1].nnz : nullptr;
  int bits = 0;
  for (int l = lane_id(); l < 28; l += MBK_WS) {
    if (l == 27) {                                        // mb_type, prediction syntax, coded_block_pattern, mb_qp_delta
      const int off = c.p.is_idr ? 0 : 5;
      switch (m.mb_type) {
        case MBT_I4x4:
          bits += ue_len((uint32_t)off);
          for (int k = 0; k < 16; k++) bits += m.prev_i4_flag[k] ? 1 : 4;
          bits += ue_len(m.chroma_mode);
          break;
        case MBT_I16x16:
          bits += ue_len((uint32_t)(1 + off + m.i16_mode + (cbp_c << 2) + (cbp_l ? 12 : 0))) + ue_len(m.chroma_mode);
          break;
        case MBT_P16x16: bits += 1 + se_len(m.mvd[0][0]) + se_len(m.mvd[0][1]); break;
        case MBT_P16x8:
        case MBT_P8x16:
          bits += 3 + se_len(m.mvd[0][0]) + se_len(m.mvd[0][1]) + se_len(m.mvd[1][0]) + se_len(m.mvd[1][1]);
          break;
        case MBT_P8x8:
          bits += ue_len(4) + 4;                         // P_8x8ref0, four sub_mb_type ue(0)
          for (int k = 0; k < 4; k++) bits += se_len(m.mvd[k][0]) + se_len(m.mvd[k][1]);
          break;
        default: break;
      }
      if (m.mb_type == MBT_I4x4) bits += ue_len(T.cbp_intra[m.cbp]);
      else if (m.mb_type != MBT_I16x16) bits += ue_len(T.cbp_inter[m.cbp]);
      if (coded) bits += 1;                               // mb_qp_delta = 0
    } else if (!coded) {
      continue;
    } else if (l < 16) {                                  // luma block l (coding order)
      if (!(cbp_l & (1 << (l >> 2)))) continue;
      const int bx = blk_x(l), by = blk_y(l);
      const int a = bx > 0 ? m.nnz[by * 4 + bx - 1] : (L ? L[by * 4 + 3] : -1);
      const int b = by > 0 ? m.nnz[(by - 1) * 4 + bx] : (Tn ? Tn[12 + bx] : -1);
      bits += cavlc_block_bits(m.luma[l], m.mb_type == MBT_I16x16 ? 15 : 16, nc_of(a, b), m.nnz[by * 4 + bx] > 0);
    } else if (l < 24) {                                  // chroma AC
      if (cbp_c != 2) continue;
      const int uv = (l - 16) >> 2, j = (l - 16) & 3, bx = j & 1, by = j >> 1, base = 16 + 4 * uv;
      const int a = bx > 0 ? m.nnz[base + by * 2] : (L ? L[base + by * 2 + 1] : -1);
      const int b = by > 0 ? m.nnz[base + bx] : (Tn ? Tn[base + 2 + bx] : -1);
      bits += cavlc_block_bits(m.chroma_ac[4 * uv + j], 15, nc_of(a, b), m.nnz[base + j] > 0);
    } else if (l < 26) {                                  // chroma DC
      if (cbp_c) bits += cavlc_block_bits(m.chroma_dc[l - 24], 4, -1, true);
    } else {                                              // l == 26: Intra16x16 luma DC
      if (m.mb_type == MBT_I16x16)                       // context of block (0,0): left MB's block (3,0), top MB's block (0,3)
        bits += cavlc_block_bits(m.luma_dc, 16, nc_of(L ? L[3] : -1, Tn ? Tn[12] : -1), true);
    }
  }
  return warp_sum(bits);
}

}  // namespace mbk
