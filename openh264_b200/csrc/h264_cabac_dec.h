// h264_cabac_dec.h — CABAC arithmetic DECODER and the binarisations of the macroblock layer (Rec. H.264 9.3.1.2, 9.3.2, 9.3.3.2),
// the inverse of h264_cabac.cpp.  Used by h264_parse.cpp for streams with entropy_coding_mode_flag = 1 (I and P slices, frame
// macroblocks, 4x4 transform).  The reference's counterpart: codec/decoder/core/src/cabac_decoder.cpp (DecodeBinCabac,
// DecodeBypassCabac, DecodeTerminateCabac) and parse_mb_syn_cabac.cpp; this file follows the Recommendation's bit-serial
// description (9-bit offset register), the context variables are the same tables the encoder uses (cabac_tables.h).
// Context index increments that depend on neighbouring macroblocks are computed by the caller (h264_parse.cpp keeps the
// per-macroblock context records) and passed in.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#include "cabac_tables.h"

namespace b2h264 {

class CabacDecoder {
 public:
  CabacDecoder(const uint8_t* p, size_t nbytes) : p_(p), nbytes_(nbytes) {}
  bool ok() const { return !overrun_; }
  size_t pos() const { return byte_ * 8 - (size_t)avail_; }      // bits consumed so far

  void init_contexts(int slice_qp, int table /* 0 = I slice, 1 + cabac_init_idc */) {
    const int qp = slice_qp < 0 ? 0 : slice_qp > 51 ? 51 : slice_qp;
    for (int i = 0; i < 460; i++) {
      int pre = ((kCabacInit[i][table][0] * qp) >> 4) + kCabacInit[i][table][1];
      pre = pre < 1 ? 1 : pre > 126 ? 126 : pre;
      ctx_[i] = pre <= 63 ? (uint8_t)((63 - pre) << 1) : (uint8_t)(((pre - 64) << 1) | 1);      // pStateIdx << 1 | valMPS
    }
  }
  // 9.3.1.2: at the start of the slice data and after the samples of an I_PCM macroblock (bit_pos: byte aligned)
  void init_engine(size_t bit_pos) {
    byte_ = (bit_pos + 7) >> 3; buf_ = 0; avail_ = 0;
    range_ = 510;
    offset_ = read_bits(9);
  }
  int decision(int ctx) {
    const uint32_t c = ctx_[ctx], st = c >> 1;
    const uint32_t lps = kCabacRangeLps[st][(range_ >> 6) & 3];
    range_ -= lps;
    int bin;
    if (offset_ >= range_) {
      bin = (int)((c & 1) ^ 1);
      offset_ -= range_;
      range_ = lps;
      ctx_[ctx] = (uint8_t)((kCabacNextLps[st] << 1) | ((c & 1) ^ (st == 0)));
    } else {
      bin = (int)(c & 1);
      ctx_[ctx] = (uint8_t)((kCabacNextMps[st] << 1) | (c & 1));
    }
    if (range_ < 256) {
      const int shift = __builtin_clz(range_) - 23;      // RenormD: up to 7 places (range_ >= 2... an LPS range is >= 6)
      range_ <<= shift;
      offset_ = (offset_ << shift) | read_bits(shift);
    }
    return bin;
  }
  int bypass() {
    offset_ = (offset_ << 1) | read_bits(1);
    if (offset_ >= range_) { offset_ -= range_; return 1; }
    return 0;
  }
  // end_of_slice_flag / the I_PCM bin of mb_type.  After a 1 the read position is just behind the encoder's flush (whose last
  // bit is the rbsp_stop_one_bit, 9.3.4.5): nothing more is read
  int terminate() {
    range_ -= 2;
    if (offset_ >= range_) return 1;
    if (range_ < 256) { range_ <<= 1; offset_ = (offset_ << 1) | read_bits(1); }
    return 0;
  }
  // k-th order Exp-Golomb suffix, bypass coded (9.3.2.3); -1: not a valid code
  int64_t exp_golomb_bypass(int k) {
    int64_t v = 0;
    while (bypass()) {
      v += (int64_t)1 << k;
      if (++k > 24) { overrun_ = true; return -1; }
    }
    while (k--) v += (int64_t)bypass() << k;
    return v;
  }

  // ---- syntax elements (ctxIdx: Table 9-34; bins: 9.3.2.5 and Tables 9-36 .. 9-38) ----
  // mb_type of an I slice (prefix_ctx = 3 + inc, rest 3 + 3..7) or the suffix inside a P slice (prefix_ctx = 17, rest 17 + 1..3);
  // returns mb_type 0..25
  int mb_type_intra(int prefix_ctx, bool in_p) {
    if (!decision(prefix_ctx)) return 0;                                  // I_NxN
    if (terminate()) return 25;                                           // I_PCM
    const int base = in_p ? prefix_ctx : 3;                               // P slices 17, B slices 32
    const int c_l = in_p ? base + 1 : base + 3, c_c0 = in_p ? base + 2 : base + 4, c_c1 = in_p ? base + 2 : base + 5;
    const int c_m0 = in_p ? base + 3 : base + 6, c_m1 = in_p ? base + 3 : base + 7;
    const int luma = decision(c_l);
    int chroma = decision(c_c0);
    if (chroma) chroma += decision(c_c1);
    const int m0 = decision(c_m0), m1 = decision(c_m1);                   // 9.3.3.1.2: the same two contexts with or without the extra chroma bin
    return 1 + (m0 * 2 + m1) + 4 * chroma + 12 * luma;
  }
  // P slice: 0 P_L0_16x16, 1 P_L0_L0_16x8, 2 P_L0_L0_8x16, 3 P_8x8, 5.. = 5 + intra type
  int mb_type_p() {
    if (decision(14)) return 5 + mb_type_intra(17, true);
    if (!decision(15)) return decision(16) ? 3 : 0;
    return decision(17) ? 1 : 2;
  }
  // B slice (Table 9-37 b): 0 B_Direct_16x16, 1..22 as Table 7-14, 23.. = 23 + intra type; inc: neighbours that are neither
  // B_Skip nor B_Direct_16x16
  int mb_type_b(int inc) {
    if (!decision(27 + inc)) return 0;
    if (!decision(27 + 3)) return 1 + decision(27 + 5);
    int bits = decision(27 + 4) << 3;
    bits |= decision(27 + 5) << 2;
    bits |= decision(27 + 5) << 1;
    bits |= decision(27 + 5);
    if (bits < 8) return bits + 3;
    if (bits == 13) return 23 + mb_type_intra(32, true);
    if (bits == 14) return 11;
    if (bits == 15) return 22;
    bits = (bits << 1) | decision(27 + 5);
    return bits - 4;
  }
  int sub_mb_type_b() {                                                    // 0..12 as Table 7-18
    if (!decision(36)) return 0;
    if (!decision(37)) return 1 + decision(39);
    int type = 3;
    if (decision(38)) {
      if (decision(39)) return 11 + decision(39);
      type += 4;
    }
    type += 2 * decision(39);
    type += decision(39);
    return type;
  }
  int sub_mb_type_p() {                                                    // 0 8x8, 1 8x4, 2 4x8, 3 4x4
    if (decision(21)) return 0;
    if (!decision(22)) return 1;
    return decision(23) ? 2 : 3;
  }
  int ref_idx(int inc) {                                                   // unary, ctx 54 + inc / 58 / 59
    if (!decision(54 + inc)) return 0;
    int v = 1;
    if (!decision(58)) return v;
    for (v = 2; decision(59); v++) if (v > 32) { overrun_ = true; return 0; }
    return v;
  }
  int mvd(int ctx_base, int sum) {                                         // UEG3, uCoff 9, signed
    int inc = sum > 32 ? 2 : sum > 2 ? 1 : 0;
    int a = 0;
    while (a < 9 && decision(ctx_base + inc)) {
      inc = a == 0 ? 3 : (inc < 6 ? inc + 1 : 6);
      a++;
    }
    if (a == 9) {
      const int64_t s = exp_golomb_bypass(3);
      if (s < 0 || s > 1 << 20) { overrun_ = true; return 0; }
      a += (int)s;
    }
    if (a && bypass()) a = -a;
    return a;
  }
  int intra_chroma_pred_mode(int inc) {
    if (!decision(64 + inc)) return 0;
    if (!decision(67)) return 1;
    return decision(67) ? 3 : 2;
  }
  int mb_qp_delta(int inc) {
    if (!decision(60 + inc)) return 0;
    int v = 1;
    if (decision(62)) for (v = 2; decision(63); v++) if (v > 104) { overrun_ = true; return 0; }
    return (v & 1) ? (v + 1) / 2 : -(v / 2);
  }
  // residual_block_cabac of an 8x8 luma block (ctxBlockCat 5: no coded_block_flag in 4:2:0; frame-coded context maps of Table 9-43):
  // fills lv[0..64) in 8x8 zig-zag order, returns the number of non-zero levels or -1
  int residual_levels8x8(int16_t* lv) {
    static const uint8_t kSig[63] = {0, 1, 2, 3, 4, 5, 5, 4, 4, 3, 3, 4, 4, 4, 5, 5, 4, 4, 4, 4, 3, 3, 6, 7, 7, 7, 8, 9, 10, 9, 8, 7,
                                     7, 6, 11, 12, 13, 11, 6, 7, 8, 9, 14, 10, 9, 8, 6, 11, 12, 13, 11, 6, 9, 14, 10, 9, 11, 12, 13, 11, 14, 10, 12};
    static const uint8_t kLast[63] = {0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2,
                                      3, 3, 3, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 8, 8, 8};
    for (int i = 0; i < 64; i++) lv[i] = 0;
    int sig_pos[64], n = 0, i = 0;
    for (; i < 63; i++) {
      if (decision(402 + kSig[i])) {
        sig_pos[n++] = i;
        if (decision(417 + kLast[i])) break;
      }
    }
    if (i == 63) sig_pos[n++] = 63;
    int eq1 = 0, gt1 = 0;
    const int base = 426;
    for (int k = n - 1; k >= 0; k--) {
      int a = 0;
      if (decision(base + (gt1 ? 0 : (1 + eq1 < 4 ? 1 + eq1 : 4)))) {
        const int ctx = base + 5 + (gt1 < 4 ? gt1 : 4);
        a = 1;
        while (a < 14 && decision(ctx)) a++;
        if (a == 14) {
          const int64_t s = exp_golomb_bypass(0);
          if (s < 0 || s > 1 << 16) { overrun_ = true; return -1; }
          a += (int)s;
        }
        gt1++;
      } else {
        eq1++;
      }
      const int v = a + 1;
      if (v > 32767) { overrun_ = true; return -1; }
      lv[sig_pos[k]] = (int16_t)(bypass() ? -v : v);
    }
    return n;
  }
  // residual_block_cabac without the coded_block_flag (7.3.5.3.3): fills lv[0..max_coef) (scan order), returns the number of
  // non-zero levels or -1
  int residual_levels(int cat, int16_t* lv, int max_coef) {
    static const int kSigOff[5] = {0, 15, 29, 44, 47}, kAbsOff[5] = {0, 10, 20, 30, 39};
    for (int i = 0; i < max_coef; i++) lv[i] = 0;
    int sig_pos[16], n = 0;
    int i = 0;
    for (; i < max_coef - 1; i++) {
      const int inc = cat == 3 ? (i < 2 ? i : 2) : i;
      if (decision(105 + kSigOff[cat] + inc)) {
        sig_pos[n++] = i;
        if (decision(166 + kSigOff[cat] + inc)) break;
      }
    }
    if (i == max_coef - 1) sig_pos[n++] = max_coef - 1;                   // no "last" flag seen: the final position is significant
    int eq1 = 0, gt1 = 0;
    const int base = 227 + kAbsOff[cat];
    for (int k = n - 1; k >= 0; k--) {
      int a = 0;
      if (decision(base + (gt1 ? 0 : (1 + eq1 < 4 ? 1 + eq1 : 4)))) {
        const int lim = 4 - (cat == 3 ? 1 : 0);
        const int ctx = base + 5 + (gt1 < lim ? gt1 : lim);
        a = 1;
        while (a < 14 && decision(ctx)) a++;
        if (a == 14) {
          const int64_t s = exp_golomb_bypass(0);
          if (s < 0 || s > 1 << 16) { overrun_ = true; return -1; }
          a += (int)s;
        }
        gt1++;
      } else {
        eq1++;
      }
      const int v = a + 1;
      if (v > 32767) { overrun_ = true; return -1; }
      lv[sig_pos[k]] = (int16_t)(bypass() ? -v : v);
    }
    return n;
  }

 private:
  // the next n <= 16 bits; bits beyond the end of the payload read as 0 (a few of them are legal look-ahead of the engine)
  uint32_t read_bits(int n) {
    if (avail_ < n) {
      while (avail_ <= 56) {
        uint8_t b = 0;
        if (byte_ < nbytes_) b = p_[byte_];
        else if (++past_end_ > 16) overrun_ = true;
        byte_++;
        buf_ = (buf_ << 8) | b;
        avail_ += 8;
      }
    }
    avail_ -= n;
    return (uint32_t)(buf_ >> avail_) & ((1u << n) - 1u);
  }
  const uint8_t* p_;
  size_t nbytes_, byte_ = 0;
  uint64_t buf_ = 0;
  int avail_ = 0;
  uint8_t ctx_[460];
  uint32_t range_ = 510, offset_ = 0;
  int past_end_ = 0;
  bool overrun_ = false;
};

}  // namespace b2h264
