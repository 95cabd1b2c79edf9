// enc_host.h — host-side stream control for one encoder instance: parameter sets, frame numbering,
// slice serialisation from the device's MbOut records.  Mirrors the small part of
// codec/encoder/core/src/encoder_ext.cpp (WelsEncoderEncodeExt :3441, frame_num / idr_pic_id handling
// :3131,:3253) that the supported configuration exercises: CAMERA_VIDEO_REAL_TIME, 1 spatial + 1
// temporal layer, RC_OFF_MODE, SM_SINGLE_SLICE, CAVLC, 1 reference frame, IDR only at the start
// (uiIntraPeriod = 0) or on ForceIntraFrame.
#pragma once
#include <vector>

#include "enc_types.h"
#include "h264_bitstream.h"

namespace b2h264 {

// Host twin of k_pack_records (enc_kernels.cu): compacts n records into `dst` (worst case n * sizeof(MbOut) bytes) and fills idx[mb] =
// offset of the macroblock's record in 32-byte units, or -1 - qp for a P_SKIP macroblock (the decoder's hand-over keeps the
// quantiser of a skipped macroblock: deblocking needs it; the encoder's reader only tests the sign).  Returns the units written.
// Record: 128-byte head (MbOut bytes [0, 112) + chroma_dc, presence mask of the 24 residual blocks in pad0) + 32 bytes per present block.
int pack_records_compact(const MbOut* mbs, int n, uint8_t* dst, int32_t* idx);

struct StreamCtl {
  StreamParams sp;
  float fps;
  int frame_num = 0;
  int idr_pic_id = 0;
  long frames_coded = 0;
  bool force_idr = true;
  // eSpsPpsIdStrategy: INCREASING_ID (the reference default) gives every emitted SPS/PPS pair the next id
  // (mod MAX_SPS_COUNT 32 / MAX_PPS_COUNT 57, paraset_strategy.cpp:338-369); CONSTANT_ID keeps 0/0
  bool increasing_ids = true;
  int parasets_written = 0;
  bool fast_mode = false;                   // iComplexityMode == LOW_COMPLEXITY
  bool record_mb_bits = false;              // keep the writer's bits per macroblock of the last picture (parity of the device count)
  std::vector<int32_t> last_mb_bits;

  void init(int width, int height, int qp, float fps_, int target_bitrate, int entropy_cabac = 0, int profile_idc = 0);
  bool next_is_idr() const { return force_idr; }
  // iLoopFilterDisableIdc / iLoopFilterAlphaC0Offset / iLoopFilterBetaOffset (after init): with one slice per picture idc 2 is
  // idc 0 (encoder_ext.cpp:1109-1114); false if a value is out of range (encoder_ext.cpp:316-318)
  bool set_loop_filter(int idc, int alpha_c0_offset, int beta_offset) {
    if (idc < 0 || idc > 2 || alpha_c0_offset < -6 || alpha_c0_offset > 6 || beta_offset < -6 || beta_offset > 6) return false;
    sp.dbk_idc = idc == 2 ? 0 : idc; sp.dbk_alpha_div2 = alpha_c0_offset; sp.dbk_beta_div2 = beta_offset;
    return true;
  }
  // geometry of the padded pictures (picture_handle.cpp:60-85: 32-pixel luma padding)
  int rec_stride_y() const { return sp.mb_w * 16 + 64; }
  int rec_stride_c() const { return sp.mb_w * 8 + 32; }
  int rec_rows_y() const { return sp.mb_h * 16 + 64; }
  int rec_rows_c() const { return sp.mb_h * 8 + 32; }
  EncFrameParams frame_params(bool idr, bool ref_is_p) const;
  // serialises the access unit of the frame just coded (SPS+PPS+IDR slice, or P slice)
  void write_access_unit(bool idr, const MbOut* mbs, std::vector<uint8_t>* au);                 // one record per MB
  // compact form (k_pack_records): idx[mb] = offset of the macroblock's record in `packed` in 32-byte units, or -1 for P_SKIP
  void write_access_unit_packed(bool idr, const MbOut* packed, const int32_t* idx, std::vector<uint8_t>* au);
 private:
  void write_au(bool idr, const MbOut* const* recs, std::vector<uint8_t>* au);
  std::vector<const MbOut*> recs_;
  std::vector<uint8_t> rbsp_;               // payload of the NAL being written (reused from picture to picture)
  std::vector<MbOut> expand_;               // compact records of the coded macroblocks, expanded for the slice writer
 public:
  int last_coded_mbs = 0;                   // macroblocks of the last picture that were not P_SKIP (write_access_unit_packed)
};

// copies a w x h I420 picture into MB-aligned planes; rows/cols beyond the picture are 0 (luma) / 0x80
// (chroma) like CWelsPreProcess::Padding (wels_preprocess.cpp:1250)
void pad_source(const uint8_t* yuv, int w, int h, int mb_w, int mb_h, uint8_t* y, uint8_t* u, uint8_t* v);

}  // namespace b2h264
