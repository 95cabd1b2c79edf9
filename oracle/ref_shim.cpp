/*
 * oracle/ref_shim.cpp — plain-C entry points onto the UNMODIFIED reference (cisco/openh264).
 *
 * TEST INFRASTRUCTURE ONLY.  Compiled (only where /root/reference exists) against the reference's
 * own headers and linked to oracle/_ref/libopenh264_ref.so; output oracle/_ref/librefshim.so.
 * Every ref_* function has exactly the signature of the oracle's orc_* twin (h264_oracle.h) and
 * forwards to the reference function through the reference's own function-pointer tables
 * initialised with uiCpuFlag = 0 (the *_c path), so tests can diff oracle vs reference call by
 * call.  Nothing here re-implements any arithmetic.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "wels_func_ptr_def.h"
#include "svc_motion_estimate.h"
#include "sample.h"
#include "encode_mb_aux.h"
#include "decode_mb_aux.h"
#include "md.h"
#include "mc.h"
#include "sad_common.h"
#include "deblocking_common.h"
#include "expand_pic.h"
#include "slice.h"
#include "svc_enc_frame.h"
#include "picture.h"
#include "encoder_context.h"

#include "../oracle/h264_oracle.h"

using namespace WelsEnc;

namespace WelsDec {
void IdctResAddPred_c (uint8_t* pPred, const int32_t kiStride, int16_t* pRs);
void IdctResAddPred8x8_c (uint8_t* pPred, const int32_t kiStride, int16_t* pRs);
}
namespace WelsEnc {
extern const int32_t g_kiQpCostTable[52];
}
/* non-static but undeclared in deblocking_common.h (deblocking_common.cpp:5,39,93,135) */
void DeblockLumaLt4_c (uint8_t* pPix, int32_t iStrideX, int32_t iStrideY, int32_t iAlpha, int32_t iBeta, int8_t* pTc);
void DeblockLumaEq4_c (uint8_t* pPix, int32_t iStrideX, int32_t iStrideY, int32_t iAlpha, int32_t iBeta);
void DeblockChromaLt4_c (uint8_t* pPixCb, uint8_t* pPixCr, int32_t iStrideX, int32_t iStrideY, int32_t iAlpha,
                         int32_t iBeta, int8_t* pTc);
void DeblockChromaEq4_c (uint8_t* pPixCb, uint8_t* pPixCr, int32_t iStrideX, int32_t iStrideY, int32_t iAlpha,
                         int32_t iBeta);

static SWelsFuncPtrList* Funcs() {
  static SWelsFuncPtrList* p = NULL;
  if (!p) {
    p = (SWelsFuncPtrList*) calloc (1, sizeof (SWelsFuncPtrList));
    WelsInitSampleSadFunc (p, 0);
    WelsInitEncodingFuncs (p, 0);
    WelsInitReconstructionFuncs (p, 0);
    WelsCommon::InitMcFunc (&p->sMcFuncs, 0);
    WelsInitMeFunc (p, 0, false);
    InitExpandPictureFunc (&p->sExpandPicFunc, 0);
  }
  return p;
}

extern "C" {

const int16_t* ref_quant_ff (int q) { return g_kiQuantInterFF[q]; }
const int16_t* ref_quant_mf (int q) { return g_kiQuantMF[q]; }
const uint16_t* ref_dequant_coeff (int q) { return WelsCommon::g_kuiDequantCoeff[q]; }
int ref_qp_lambda (int q) { return g_kiQpCostTable[q]; }
int ref_chroma_qp (int q) { return WelsCommon::g_kuiChromaQpTable[q]; }
/* full reference table for all 52 qps, stride = 2*sz+1 (md.cpp:797) */
void ref_mvd_cost_init_all (uint16_t* table, int mvd_sz) { MvdCostInit (table, mvd_sz); }

int32_t ref_sad (int blk, const uint8_t* a, int sa, const uint8_t* b, int sb) {
  return Funcs()->sSampleDealingFuncs.pfSampleSad[blk] ((uint8_t*)a, sa, (uint8_t*)b, sb);
}
void ref_sad_four (int blk, const uint8_t* a, int sa, const uint8_t* b, int sb, int32_t out[4]) {
  Funcs()->sSampleDealingFuncs.pfSample4Sad[blk] ((uint8_t*)a, sa, (uint8_t*)b, sb, out);
}
int32_t ref_satd (int blk, const uint8_t* a, int sa, const uint8_t* b, int sb) {
  return Funcs()->sSampleDealingFuncs.pfSampleSatd[blk] ((uint8_t*)a, sa, (uint8_t*)b, sb);
}
void ref_mc_luma (const uint8_t* src, int ss, uint8_t* dst, int ds, int mvx, int mvy, int w, int h) {
  Funcs()->sMcFuncs.pMcLumaFunc (src, ss, dst, ds, (int16_t)mvx, (int16_t)mvy, w, h);
}
void ref_mc_chroma (const uint8_t* src, int ss, uint8_t* dst, int ds, int mvx, int mvy, int w, int h) {
  Funcs()->sMcFuncs.pMcChromaFunc (src, ss, dst, ds, (int16_t)mvx, (int16_t)mvy, w, h);
}
void ref_pixel_avg (uint8_t* dst, int ds, const uint8_t* a, int sa, const uint8_t* b, int sb, int w, int h) {
  Funcs()->sMcFuncs.pfSampleAveraging (dst, ds, a, sa, b, sb, w, h);
}
/* the three half-sample plane producers used by MeRefineFracPixel (mc.h:46-49) */
void ref_halfpel (int which, const uint8_t* src, int ss, uint8_t* dst, int ds, int w, int h) {
  SMcFunc* m = &Funcs()->sMcFuncs;
  (which == 0 ? m->pfLumaHalfpelHor : which == 1 ? m->pfLumaHalfpelVer : m->pfLumaHalfpelCen) (src, ss, dst, ds, w, h);
}

void ref_dct4x4 (int16_t* d, const uint8_t* p1, int s1, const uint8_t* p2, int s2) {
  Funcs()->pfDctT4 (d, (uint8_t*)p1, s1, (uint8_t*)p2, s2);
}
void ref_dct_four4x4 (int16_t* d, const uint8_t* p1, int s1, const uint8_t* p2, int s2) {
  Funcs()->pfDctFourT4 (d, (uint8_t*)p1, s1, (uint8_t*)p2, s2);
}
void ref_quant4x4 (int16_t* d, const int16_t* ff, const int16_t* mf) { Funcs()->pfQuantization4x4 (d, ff, mf); }
void ref_quant4x4_dc (int16_t* d, int16_t ff, int16_t mf) { Funcs()->pfQuantizationDc4x4 (d, ff, mf); }
void ref_quant_four4x4 (int16_t* d, const int16_t* ff, const int16_t* mf) { Funcs()->pfQuantizationFour4x4 (d, ff, mf); }
void ref_quant_four4x4_max (int16_t* d, const int16_t* ff, const int16_t* mf, int16_t* mx) {
  Funcs()->pfQuantizationFour4x4Max (d, ff, mf, mx);
}
int32_t ref_hadamard_quant2x2_skip (const int16_t* rs, int16_t ff, int16_t mf) {
  return Funcs()->pfQuantizationHadamard2x2Skip ((int16_t*)rs, ff, mf);
}
int32_t ref_hadamard_quant2x2 (int16_t* rs, int16_t ff, int16_t mf, int16_t* dct, int16_t* block) {
  return Funcs()->pfQuantizationHadamard2x2 (rs, ff, mf, dct, block);
}
void ref_hadamard_t4_dc (int16_t* dc, const int16_t* dct) { Funcs()->pfTransformHadamard4x4Dc (dc, (int16_t*)dct); }
void ref_scan4x4_dcac (int16_t* level, const int16_t* dct) { Funcs()->pfScan4x4 (level, (int16_t*)dct); }
void ref_scan4x4_ac (int16_t* level, const int16_t* dct) { Funcs()->pfScan4x4Ac (level, (int16_t*)dct); }
int32_t ref_single_ctr4x4 (const int16_t* d) { return Funcs()->pfCalculateSingleCtr4x4 ((int16_t*)d); }
int32_t ref_nonzero_count (const int16_t* l) { return Funcs()->pfGetNoneZeroCount ((int16_t*)l); }

void ref_ihadamard4x4_dc (int16_t* r) { WelsIHadamard4x4Dc (r); }
void ref_dequant_luma_dc4x4 (int16_t* r, int qp) { WelsDequantLumaDc4x4 (r, qp); }
void ref_dequant_ihadamard4x4 (int16_t* r, uint16_t mf) { Funcs()->pfDequantizationIHadamard4x4 (r, mf); }
void ref_dequant_ihadamard2x2_dc (int16_t* r, uint16_t mf) { WelsDequantIHadamard2x2Dc (r, mf); }
void ref_dequant4x4 (int16_t* r, const uint16_t* mf) { Funcs()->pfDequantization4x4 (r, mf); }
void ref_dequant_four4x4 (int16_t* r, const uint16_t* mf) { Funcs()->pfDequantizationFour4x4 (r, mf); }
void ref_idct4x4_rec (uint8_t* rec, int rs, const uint8_t* pred, int ps, const int16_t* dct) {
  Funcs()->pfIDctT4 (rec, rs, (uint8_t*)pred, ps, (int16_t*)dct);
}
void ref_idct_four4x4_rec (uint8_t* rec, int rs, const uint8_t* pred, int ps, const int16_t* dct) {
  Funcs()->pfIDctFourT4 (rec, rs, (uint8_t*)pred, ps, (int16_t*)dct);
}
void ref_idct_rec_i16x16_dc (uint8_t* rec, int rs, const uint8_t* pred, int ps, const int16_t* dc) {
  Funcs()->pfIDctI16x16Dc (rec, rs, (uint8_t*)pred, ps, (int16_t*)dc);
}
void ref_idct_res_add_pred (uint8_t* pred, int stride, const int16_t* rs) { WelsDec::IdctResAddPred_c (pred, stride, (int16_t*)rs); }
void ref_idct_res_add_pred8x8 (uint8_t* pred, int stride, const int16_t* rs) { WelsDec::IdctResAddPred8x8_c (pred, stride, (int16_t*)rs); }

void ref_deblock_luma_lt4 (uint8_t* pix, int sx, int sy, int alpha, int beta, const int8_t tc[4]) {
  DeblockLumaLt4_c (pix, sx, sy, alpha, beta, (int8_t*)tc);
}
void ref_deblock_luma_eq4 (uint8_t* pix, int sx, int sy, int alpha, int beta) { DeblockLumaEq4_c (pix, sx, sy, alpha, beta); }
void ref_deblock_chroma_lt4 (uint8_t* cb, uint8_t* cr, int sx, int sy, int alpha, int beta, const int8_t tc[4]) {
  DeblockChromaLt4_c (cb, cr, sx, sy, alpha, beta, (int8_t*)tc);
}
void ref_deblock_chroma_eq4 (uint8_t* cb, uint8_t* cr, int sx, int sy, int alpha, int beta) {
  DeblockChromaEq4_c (cb, cr, sx, sy, alpha, beta);
}
/* pad = 32 -> luma expander, pad = 16 -> chroma expander (expand_pic.h:49-50) */
void ref_expand_plane (uint8_t* pic, int stride, int w, int h, int pad) {
  if (pad == 32) Funcs()->sExpandPicFunc.pfExpandLumaPicture (pic, stride, w, h);
  else Funcs()->sExpandPicFunc.pfExpandChromaPicture[0] (pic, stride, w, h);
}

/* WelsMotionEstimateSearch (svc_motion_estimate.cpp:170) driven with the reference's own structs */
void ref_me_search (const uint8_t* cur, int cs, const uint8_t* ref, int rs, const orc_me_job* j, orc_me_result* out) {
  static uint16_t* cost = NULL;
  const int kSz = 1100, kStride = 2 * kSz + 1;
  if (!cost) { cost = (uint16_t*) malloc (sizeof (uint16_t) * 52 * kStride); MvdCostInit (cost, kStride); }
  SWelsFuncPtrList* f = Funcs();
  static SDqLayer* layer = NULL; static SPicture* pic = NULL; static SSlice* slice = NULL;
  if (!layer) {
    layer = (SDqLayer*) calloc (1, sizeof (SDqLayer));
    pic = (SPicture*) calloc (1, sizeof (SPicture));
    slice = (SSlice*) calloc (1, sizeof (SSlice));
    layer->pRefPic = pic;
  }
  layer->iEncStride[0] = cs;
  pic->iLineSize[0] = rs;
  slice->uiMvcNum = j->n_mvc;
  for (int i = 0; i < j->n_mvc; i++) { slice->sMvc[i].iMvX = j->mvc[i][0]; slice->sMvc[i].iMvY = j->mvc[i][1]; }
  slice->sMvStartMin.iMvX = j->mv_min_x; slice->sMvStartMin.iMvY = j->mv_min_y;
  slice->sMvStartMax.iMvX = j->mv_max_x; slice->sMvStartMax.iMvY = j->mv_max_y;
  for (int b = 0; b < BLOCK_SIZE_ALL; b++) f->pfSearchMethod[b] = WelsDiamondSearch;
  f->pfCalculateSatd = j->calc_satd ? CalculateSatdCost : NotCalculateSatdCost;

  SWelsME me;
  memset (&me, 0, sizeof (me));
  me.pMvdCost = cost + j->qp * kStride + kSz;
  me.uSadPredISatd.uiSadPred = j->sad_pred;
  me.uiBlockSize = (uint8_t) j->blk;
  me.pEncMb = (uint8_t*) cur + j->cur_off;
  me.pRefMb = me.pColoRefMb = (uint8_t*) ref + j->ref_off;
  me.sMvp.iMvX = j->mvp_x; me.sMvp.iMvY = j->mvp_y;
  WelsMotionEstimateSearch (f, layer, &me, slice);
  out->mv_x = me.sMv.iMvX; out->mv_y = me.sMv.iMvY;
  out->sad_cost = me.uiSadCost; out->satd_cost = me.uiSatdCost;
  out->ref_off = (int32_t) (me.pRefMb - ref);
}

/* WelsMotionCrossSearch (svc_motion_estimate.cpp:620) with the reference's own structs; io = SWelsME::sMv (integer pels) / uiSadCost / pRefMb */
void ref_me_cross_search (const uint8_t* cur, int cs, const uint8_t* ref, int rs, const orc_cross_job* j, orc_me_result* io) {
  static uint16_t* cost = NULL;
  const int kSz = 1100, kStride = 2 * kSz + 1;
  if (!cost) { cost = (uint16_t*) malloc (sizeof (uint16_t) * 52 * kStride); MvdCostInit (cost, kStride); }
  SWelsFuncPtrList* f = Funcs();
  f->pfVerticalFullSearch = LineFullSearch_c;
  f->pfHorizontalFullSearch = LineFullSearch_c;
  static SSlice* slice = NULL;
  if (!slice) slice = (SSlice*) calloc (1, sizeof (SSlice));
  slice->sMvStartMin.iMvX = j->mv_min_x; slice->sMvStartMin.iMvY = j->mv_min_y;
  slice->sMvStartMax.iMvX = j->mv_max_x; slice->sMvStartMax.iMvY = j->mv_max_y;
  SWelsME me;
  memset (&me, 0, sizeof (me));
  me.pMvdCost = cost + j->qp * kStride + kSz;
  me.uiBlockSize = (uint8_t) j->blk;
  me.pEncMb = (uint8_t*) cur + j->cur_off;
  me.pColoRefMb = (uint8_t*) ref + j->ref_off;
  me.pRefMb = (uint8_t*) ref + io->ref_off;
  me.sMvp.iMvX = j->mvp_x; me.sMvp.iMvY = j->mvp_y;
  me.sMv.iMvX = io->mv_x; me.sMv.iMvY = io->mv_y;
  me.uiSadCost = io->sad_cost;
  me.uiSadCostThreshold = j->sad_cost_threshold;
  me.iCurMeBlockPixX = 64; me.iCurMeBlockPixY = 64;          /* only differences of positions enter the result */
  WelsMotionCrossSearch (f, &me, slice, cs, rs);
  io->mv_x = me.sMv.iMvX; io->mv_y = me.sMv.iMvY;
  io->sad_cost = me.uiSadCost;
  io->ref_off = (int32_t) (me.pRefMb - ref);
}

} // extern "C"

/* ------------------------------------------------------------------------------------------------
 * Frame-level drivers through the reference's own PUBLIC API (codec_api.h): used as the end-to-end
 * oracle (bitstream / YUV parity) and as the CPU baseline ("kind": "reference").
 * ---------------------------------------------------------------------------------------------- */
#include "codec_api.h"
#include <time.h>

static double now_s() { struct timespec t; clock_gettime (CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

extern "C" {

/* Encode n I420 frames (tightly packed w*h*3/2 each) at constant QP, camera mode, single layer, CAVLC,
 * deblock on, no denoise/BGD/AQ/scene-change/LTR, IDR only at frame 0 — the configuration of
 * BASELINE.md configs 2/3.  complexity: 0 low, 1 medium, 2 high.  threads = iMultipleThreadIdc
 * (1 = single thread, single slice; >1 = that many fixed slices + threads: a speed reference only,
 * it changes the bitstream).  Returns total bytes written to out (or <0 on error); seconds spent in
 * EncodeFrame in *enc_seconds; per-frame byte counts in frame_bytes[n]. */
/* entropy coder / profile of the following ref_encode calls (0 / 66 = CAVLC Baseline, the default; 1 / 0 = CABAC with the profile the
 * reference picks itself, i.e. High; 1 / 77 = CABAC Main) */
static int g_entropy_cabac = 0, g_profile_idc = 66, g_intra_period = 0;
void ref_set_entropy (int cabac, int profile_idc) { g_entropy_cabac = cabac; g_profile_idc = profile_idc; }
static int g_dbk_idc = 0, g_dbk_alpha = 0, g_dbk_beta = 0;
void ref_set_loop_filter (int idc, int alpha, int beta) { g_dbk_idc = idc; g_dbk_alpha = alpha; g_dbk_beta = beta; }   /* iLoopFilterDisableIdc / offsets */
void ref_set_intra_period (int n) { g_intra_period = n; }      /* uiIntraPeriod of the following ref_encode calls (0: the default) */

long ref_encode (const uint8_t* yuv, int w, int h, int n, int qp, int complexity, int threads, float fps,
                 uint8_t* out, long out_cap, int32_t* frame_bytes, double* enc_seconds) {
  ISVCEncoder* enc = NULL;
  if (WelsCreateSVCEncoder (&enc) || !enc) return -1;
  SEncParamExt p;
  enc->GetDefaultParams (&p);
  p.iUsageType = CAMERA_VIDEO_REAL_TIME;
  p.iPicWidth = w; p.iPicHeight = h;
  p.iTargetBitrate = 5000000;
  p.iRCMode = RC_OFF_MODE;
  p.fMaxFrameRate = fps;
  p.iTemporalLayerNum = 1;
  p.iSpatialLayerNum = 1;
  p.iComplexityMode = (ECOMPLEXITY_MODE) complexity;
  p.uiIntraPeriod = (unsigned int) g_intra_period;
  p.iNumRefFrame = 1;
  p.iEntropyCodingModeFlag = g_entropy_cabac;
  p.bEnableFrameSkip = false;
  p.bEnableLongTermReference = false;
  p.iMultipleThreadIdc = (unsigned short) threads;
  p.iLoopFilterDisableIdc = g_dbk_idc;
  p.iLoopFilterAlphaC0Offset = g_dbk_alpha;
  p.iLoopFilterBetaOffset = g_dbk_beta;
  p.bEnableDenoise = false;
  p.bEnableBackgroundDetection = false;
  p.bEnableAdaptiveQuant = false;
  p.bEnableFrameCroppingFlag = true;
  p.bEnableSceneChangeDetect = false;
  p.sSpatialLayers[0].iVideoWidth = w;
  p.sSpatialLayers[0].iVideoHeight = h;
  p.sSpatialLayers[0].fFrameRate = fps;
  p.sSpatialLayers[0].iSpatialBitrate = 5000000;
  p.sSpatialLayers[0].iDLayerQp = qp;
  p.sSpatialLayers[0].uiProfileIdc = (EProfileIdc) g_profile_idc;
  if (threads > 1) {
    p.sSpatialLayers[0].sSliceArgument.uiSliceMode = SM_FIXEDSLCNUM_SLICE;
    p.sSpatialLayers[0].sSliceArgument.uiSliceNum = threads;
  } else {
    p.sSpatialLayers[0].sSliceArgument.uiSliceMode = SM_SINGLE_SLICE;
  }
  if (enc->InitializeExt (&p)) { WelsDestroySVCEncoder (enc); return -2; }
  int lvl = WELS_LOG_QUIET;
  enc->SetOption (ENCODER_OPTION_TRACE_LEVEL, &lvl);
  long total = 0;
  double secs = 0;
  const size_t fsz = (size_t) w * h * 3 / 2;
  for (int i = 0; i < n; i++) {
    SSourcePicture pic;
    memset (&pic, 0, sizeof (pic));
    pic.iColorFormat = videoFormatI420;
    pic.iPicWidth = w; pic.iPicHeight = h;
    pic.iStride[0] = w; pic.iStride[1] = pic.iStride[2] = w / 2;
    pic.pData[0] = (uint8_t*) yuv + i * fsz;
    pic.pData[1] = pic.pData[0] + (size_t) w * h;
    pic.pData[2] = pic.pData[1] + (size_t) w * h / 4;
    pic.uiTimeStamp = (long long) (i * 1000.0 / fps);
    SFrameBSInfo info;
    memset (&info, 0, sizeof (info));
    const double t0 = now_s();
    const int rc = enc->EncodeFrame (&pic, &info);
    secs += now_s() - t0;
    if (rc) { WelsDestroySVCEncoder (enc); return -3; }
    int fb = 0;
    if (info.eFrameType != videoFrameTypeSkip) {
      for (int l = 0; l < info.iLayerNum; l++) {
        int sz = 0;
        for (int k = 0; k < info.sLayerInfo[l].iNalCount; k++) sz += info.sLayerInfo[l].pNalLengthInByte[k];
        if (total + sz > out_cap) { WelsDestroySVCEncoder (enc); return -4; }
        memcpy (out + total, info.sLayerInfo[l].pBsBuf, sz);
        total += sz; fb += sz;
      }
    }
    if (frame_bytes) frame_bytes[i] = fb;
  }
  if (enc_seconds) *enc_seconds = secs;
  enc->Uninitialize();
  WelsDestroySVCEncoder (enc);
  return total;
}

/* Decode an Annex-B stream; writes cropped I420 frames back to back into out. Returns frame count
 * (or <0); *w,*h = picture size; seconds in the decode calls in *dec_seconds. */
int ref_decode (const uint8_t* bs, long len, uint8_t* out, long out_cap, int* w, int* h, double* dec_seconds) {
  ISVCDecoder* dec = NULL;
  if (WelsCreateDecoder (&dec) || !dec) return -1;
  SDecodingParam dp;
  memset (&dp, 0, sizeof (dp));
  dp.sVideoProperty.eVideoBsType = VIDEO_BITSTREAM_AVC;
  dp.eEcActiveIdc = ERROR_CON_DISABLE;
  if (dec->Initialize (&dp)) { WelsDestroyDecoder (dec); return -2; }
  int lvl = WELS_LOG_QUIET;
  dec->SetOption (DECODER_OPTION_TRACE_LEVEL, &lvl);
  long pos = 0, outpos = 0;
  int frames = 0;
  double secs = 0;
  auto emit = [&] (uint8_t** d, SBufferInfo& bi) {
    if (bi.iBufferStatus != 1) return;
    const int W = bi.UsrData.sSystemBuffer.iWidth, H = bi.UsrData.sSystemBuffer.iHeight;
    *w = W; *h = H;
    if (outpos + (long) W * H * 3 / 2 > out_cap) return;
    for (int pl = 0; pl < 3; pl++) {
      const int pw = pl ? W / 2 : W, ph = pl ? H / 2 : H, st = bi.UsrData.sSystemBuffer.iStride[pl ? 1 : 0];
      for (int y = 0; y < ph; y++) { memcpy (out + outpos, d[pl] + (size_t) y * st, pw); outpos += pw; }
    }
    frames++;
  };
  while (pos < len) {
    /* next access-unit chunk: up to the next start code that follows at least one slice NAL */
    long end = pos + 4;
    while (end + 4 <= len && ! (bs[end] == 0 && bs[end + 1] == 0 && bs[end + 2] == 0 && bs[end + 3] == 1)) end++;
    if (end + 4 > len) end = len;
    uint8_t* d[3] = {0, 0, 0};
    SBufferInfo bi;
    memset (&bi, 0, sizeof (bi));
    const double t0 = now_s();
    dec->DecodeFrameNoDelay (bs + pos, (int) (end - pos), d, &bi);
    secs += now_s() - t0;
    emit (d, bi);
    pos = end;
  }
  for (;;) {
    int32_t remain = 0;
    dec->GetOption (DECODER_OPTION_NUM_OF_FRAMES_REMAINING_IN_BUFFER, &remain);
    if (remain <= 0) break;
    uint8_t* d[3] = {0, 0, 0};
    SBufferInfo bi;
    memset (&bi, 0, sizeof (bi));
    dec->FlushFrame (d, &bi);
    if (bi.iBufferStatus != 1) break;
    emit (d, bi);
  }
  if (dec_seconds) *dec_seconds = secs;
  dec->Uninitialize();
  WelsDestroyDecoder (dec);
  return frames;
}

}  // extern "C"

/* persistent-instance variant of ref_encode for throughput runs (bench.py --impl reference / cpu_baseline):
 * one encoder per stream lives across calls so that only the first picture is an IDR */
extern "C" {
void* ref_enc_open (int w, int h, int qp, int complexity, int threads, float fps) {
  ISVCEncoder* enc = NULL;
  if (WelsCreateSVCEncoder (&enc) || !enc) return NULL;
  SEncParamExt p;
  enc->GetDefaultParams (&p);
  p.iUsageType = CAMERA_VIDEO_REAL_TIME;
  p.iPicWidth = w; p.iPicHeight = h; p.iTargetBitrate = 5000000; p.iRCMode = RC_OFF_MODE; p.fMaxFrameRate = fps;
  p.iTemporalLayerNum = 1; p.iSpatialLayerNum = 1; p.iComplexityMode = (ECOMPLEXITY_MODE) complexity;
  p.uiIntraPeriod = 0; p.iNumRefFrame = 1; p.iEntropyCodingModeFlag = 0; p.bEnableFrameSkip = false;
  p.bEnableLongTermReference = false; p.iMultipleThreadIdc = (unsigned short) threads; p.iLoopFilterDisableIdc = 0;
  p.bEnableDenoise = false; p.bEnableBackgroundDetection = false; p.bEnableAdaptiveQuant = false;
  p.bEnableFrameCroppingFlag = true; p.bEnableSceneChangeDetect = false;
  p.sSpatialLayers[0].iVideoWidth = w; p.sSpatialLayers[0].iVideoHeight = h; p.sSpatialLayers[0].fFrameRate = fps;
  p.sSpatialLayers[0].iSpatialBitrate = 5000000; p.sSpatialLayers[0].iDLayerQp = qp;
  p.sSpatialLayers[0].uiProfileIdc = PRO_BASELINE;
  p.sSpatialLayers[0].sSliceArgument.uiSliceMode = threads > 1 ? SM_FIXEDSLCNUM_SLICE : SM_SINGLE_SLICE;
  if (threads > 1) p.sSpatialLayers[0].sSliceArgument.uiSliceNum = threads;
  if (enc->InitializeExt (&p)) { WelsDestroySVCEncoder (enc); return NULL; }
  int lvl = WELS_LOG_QUIET;
  enc->SetOption (ENCODER_OPTION_TRACE_LEVEL, &lvl);
  return enc;
}
/* encodes n pictures (tightly packed I420); returns bytes produced (discarded) or <0 */
long ref_enc_frames (void* h, const uint8_t* yuv, int w, int h_, int n, long long* ts_ms) {
  ISVCEncoder* enc = (ISVCEncoder*) h;
  long total = 0;
  const size_t fsz = (size_t) w * h_ * 3 / 2;
  for (int i = 0; i < n; i++) {
    SSourcePicture pic;
    memset (&pic, 0, sizeof (pic));
    pic.iColorFormat = videoFormatI420; pic.iPicWidth = w; pic.iPicHeight = h_;
    pic.iStride[0] = w; pic.iStride[1] = pic.iStride[2] = w / 2;
    pic.pData[0] = (uint8_t*) yuv + i * fsz; pic.pData[1] = pic.pData[0] + (size_t) w * h_; pic.pData[2] = pic.pData[1] + (size_t) w * h_ / 4;
    pic.uiTimeStamp = (*ts_ms) += 33;
    SFrameBSInfo info;
    memset (&info, 0, sizeof (info));
    if (enc->EncodeFrame (&pic, &info)) return -1;
    for (int l = 0; l < info.iLayerNum; l++)
      for (int k = 0; k < info.sLayerInfo[l].iNalCount; k++) total += info.sLayerInfo[l].pNalLengthInByte[k];
  }
  return total;
}
void ref_enc_close (void* h) { ISVCEncoder* enc = (ISVCEncoder*) h; enc->Uninitialize(); WelsDestroySVCEncoder (enc); }
}
