// wels_decoder.cpp — layer 3: an ISVCDecoder (codec/api/wels/codec_api.h:346-468) over the layer-2 batched decoder
// (include/b2h264_codec.h: b2h264_dec_*).  Behavioural model: CWelsDecoder (codec/decoder/plus/src/welsDecoderExt.cpp):
// Initialize(SDecodingParam*) first (else dsInitialOptExpected, :739-744); DecodeFrameNoDelay / DecodeFrame2 take one
// access unit (or parameter sets alone) with Annex-B start codes and hand back pointers into DECODER-OWNED picture
// memory that stay valid until the next decode call, with the SBufferInfo contract of codec_def.h:197-205; a NULL / 0
// input flushes (nothing is ever buffered here: the supported stream class has no picture reordering, so every access
// unit with a slice yields its picture in the same call and NUM_OF_FRAMES_REMAINING_IN_BUFFER is always 0).
// The picture size comes from the stream: the GPU decoder is (re)created when an SPS announces a new size.
// Stream class: what layer 2 decodes (include/b2h264_codec.h: I and P slices with CAVLC or CABAC, no 8x8 transform, progressive);
// anything else is refused with dsBitstreamError and a reason on stderr — there is no CPU decoder in this library.
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "b2h264_codec.h"
#include "broker.h"
#include "codec_api.h"

namespace {

class B2Decoder : public ISVCDecoder {
 public:
  ~B2Decoder() override { Uninitialize(); }

  long EXTAPI Initialize(const SDecodingParam* p) override {
    if (!p) return cmInitParaError;
    if (p->bParseOnly) { fprintf(stderr, "[b2h264] ISVCDecoder: bParseOnly is not supported\n"); return cmUnsupportedData; }
    Uninitialize();
    par_ = *p;
    inited_ = true;
    return cmResultSuccess;
  }

  long EXTAPI Uninitialize() override {
    drop_slot();
    w_ = h_ = 0;
    inited_ = false;
    return cmResultSuccess;
  }

  DECODING_STATE EXTAPI DecodeFrame(const unsigned char* src, const int len, unsigned char** dst, int* stride, int& w, int& h) override {
    SBufferInfo bi;
    memset(&bi, 0, sizeof(bi));
    const DECODING_STATE st = DecodeFrame2(src, len, dst, &bi);
    if (bi.iBufferStatus == 1) {
      if (stride) { stride[0] = bi.UsrData.sSystemBuffer.iStride[0]; stride[1] = bi.UsrData.sSystemBuffer.iStride[1]; }
      w = bi.UsrData.sSystemBuffer.iWidth; h = bi.UsrData.sSystemBuffer.iHeight;
    }
    return st;
  }

  DECODING_STATE EXTAPI DecodeFrameNoDelay(const unsigned char* src, const int len, unsigned char** dst, SBufferInfo* info) override {
    // the reference: DecodeFrame2(src) then DecodeFrame2(NULL) OR-ing the results (welsDecoderExt.cpp:720-725); here
    // the first call already delivers the picture and the flush has nothing left
    return DecodeFrame2(src, len, dst, info);
  }

  DECODING_STATE EXTAPI DecodeFrame2(const unsigned char* src, const int len, unsigned char** dst, SBufferInfo* info) override {
    if (!inited_) return dsInitialOptExpected;
    if (!info || !dst) return dsInvalidArgument;
    const unsigned long long ts = info->uiInBsTimeStamp;
    info->iBufferStatus = 0;
    dst[0] = dst[1] = dst[2] = nullptr;
    if (!src || len <= 0) { pending_.clear(); return dsErrorFree; }   // flush: no picture is ever held back (an incomplete one is dropped)
    // Applications feed whole access units or, like the reference's console decoder, one NAL unit per call.  A picture may
    // be coded as several slices: units are collected until they cover the picture (layer 2 answers -105 while they do not).
    pending_.insert(pending_.end(), src, src + len);
    int32_t w = 0, h = 0, has_slice = 0;
    int rc = b2h264_dec_probe(pending_.data(), (int32_t)pending_.size(), &w, &h, &has_slice);
    if (rc) { pending_.clear(); return refuse(rc); }
    vcl_ = has_slice;
    if (w > 0 && h > 0 && (w != w_ || h != h_)) {                 // a (new) SPS: a stream slot of the shared decoder of that size
      drop_slot();
      pool_ = b2wels::Broker::get().attach_decoder(w, h, &slot_);
      if (!pool_ || slot_ < 0) { pool_.reset(); slot_ = -1; pending_.clear(); return dsOutOfMemory; }
      pic_ = pool_->picture(slot_);
      w_ = w; h_ = h;
    }
    if (!pool_) { pending_.clear(); return dsNoParamSets; }
    // objects of one picture size are streams of one batched GPU decoder: the units of the callers that arrive together are
    // decoded by one launch (openh264_b200/wels/broker.h); a lone decoder is served at once
    rc = pool_->decode(slot_, pending_.data(), (int32_t)pending_.size());
    const int got0 = rc == 1;
    if (rc >= 0) rc = 0;
    else if (rc <= -1000) rc = -1000 - rc;                       // CUDA error of the call
    if (rc == -105) return dsErrorFree;                           // more slices of this picture to come
    pending_.clear();
    if (rc) return refuse(rc);
    if (got0) {
      frames_++;
      info->iBufferStatus = 1;
      info->uiOutYuvTimeStamp = ts;
      info->UsrData.sSystemBuffer.iWidth = w_;
      info->UsrData.sSystemBuffer.iHeight = h_;
      info->UsrData.sSystemBuffer.iFormat = videoFormatI420;
      info->UsrData.sSystemBuffer.iStride[0] = w_;
      info->UsrData.sSystemBuffer.iStride[1] = w_ / 2;
      dst[0] = info->pDst[0] = pic_;
      dst[1] = info->pDst[1] = pic_ + (size_t)w_ * h_;
      dst[2] = info->pDst[2] = dst[1] + (size_t)(w_ / 2) * (h_ / 2);
    }
    return dsErrorFree;
  }

  DECODING_STATE EXTAPI FlushFrame(unsigned char** dst, SBufferInfo* info) override {
    if (!inited_) return dsInitialOptExpected;
    if (info) info->iBufferStatus = 0;
    if (dst) dst[0] = dst[1] = dst[2] = nullptr;
    return dsErrorFree;
  }

  DECODING_STATE EXTAPI DecodeParser(const unsigned char*, const int, SParserBsInfo*) override {
    fprintf(stderr, "[b2h264] ISVCDecoder::DecodeParser (parse-only mode) is not supported\n");
    return dsInvalidArgument;
  }

  DECODING_STATE EXTAPI DecodeFrameEx(const unsigned char* src, const int len, unsigned char* pdst, int dst_stride, int& dst_len, int& w,
                                      int& h, int& fmt) override {
    // the reference's implementation of this entry point is an empty stub that reports success (welsDecoderExt.cpp)
    (void)src; (void)len; (void)pdst; (void)dst_stride; (void)dst_len; (void)w; (void)h; (void)fmt;
    return dsErrorFree;
  }

  long EXTAPI SetOption(DECODER_OPTION id, void* v) override {
    if (!inited_ && id != DECODER_OPTION_TRACE_LEVEL && id != DECODER_OPTION_TRACE_CALLBACK && id != DECODER_OPTION_TRACE_CALLBACK_CONTEXT)
      return dsInitialOptExpected;
    if (!v) return cmInitParaError;
    switch (id) {
      case DECODER_OPTION_END_OF_STREAM: eos_ = *(bool*)v; return cmResultSuccess;
      case DECODER_OPTION_ERROR_CON_IDC: ec_ = *(int*)v; return cmResultSuccess;      // no concealment here: errors are refused
      case DECODER_OPTION_TRACE_LEVEL:
      case DECODER_OPTION_TRACE_CALLBACK:
      case DECODER_OPTION_TRACE_CALLBACK_CONTEXT:
      case DECODER_OPTION_STATISTICS_LOG_INTERVAL:
        return cmResultSuccess;
      case DECODER_OPTION_NUM_OF_THREADS: return cmResultSuccess;                     // the GPU batch replaces decoder threads
      default: return cmInitParaError;
    }
  }

  long EXTAPI GetOption(DECODER_OPTION id, void* v) override {
    if (!inited_) return dsInitialOptExpected;
    if (!v) return cmInitParaError;
    switch (id) {
      case DECODER_OPTION_END_OF_STREAM: *(int*)v = eos_; return cmResultSuccess;
      case DECODER_OPTION_VCL_NAL: *(int*)v = vcl_; return cmResultSuccess;
      case DECODER_OPTION_TEMPORAL_ID: *(int*)v = 0; return cmResultSuccess;
      case DECODER_OPTION_ERROR_CON_IDC: *(int*)v = ec_; return cmResultSuccess;
      case DECODER_OPTION_NUM_OF_FRAMES_REMAINING_IN_BUFFER: *(int*)v = 0; return cmResultSuccess;
      case DECODER_OPTION_NUM_OF_THREADS: *(int*)v = 0; return cmResultSuccess;
      case DECODER_OPTION_IS_REF_PIC: *(int*)v = 1; return cmResultSuccess;
      case DECODER_OPTION_PROFILE: *(int*)v = 66; return cmResultSuccess;
      default: return cmInitParaError;
    }
  }

 private:
  DECODING_STATE refuse(int rc) {
    const char* what = rc == -101 ? "truncated access unit" : rc == -102 ? "stream feature outside the supported class (I / P slices, CAVLC or "
                       "CABAC, 4x4 transform, progressive, no FMO / ASO)" : rc == -103 ? "invalid syntax"
                       : rc == -104 ? "slice before its parameter sets" : rc == -2 ? "picture size changed without an SPS" : "CUDA / internal error";
    fprintf(stderr, "[b2h264] ISVCDecoder: %s (%d)\n", what, rc);
    return rc == -104 ? dsNoParamSets : rc > 0 ? dsOutOfMemory : dsBitstreamError;
  }

  SDecodingParam par_;
  bool inited_ = false, eos_ = false;
  int ec_ = 0, vcl_ = 0;
  void drop_slot() {
    if (pool_) { b2wels::Broker::get().detach_decoder(pool_, slot_); pool_.reset(); }
    slot_ = -1; pic_ = nullptr;
  }
  std::shared_ptr<b2wels::DecPool> pool_;
  int slot_ = -1;
  uint8_t* pic_ = nullptr;                // the slot's page-locked output picture (owned by the pool)
  int w_ = 0, h_ = 0;
  long frames_ = 0;
  std::vector<uint8_t> pending_;         // NAL units of a picture whose slices have not all arrived yet
};

}  // namespace

extern "C" {

long WelsCreateDecoder(ISVCDecoder** pp) {
  if (!pp) return 1;
  *pp = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    fprintf(stderr, "[b2h264] WelsCreateDecoder: no CUDA device — this library has no CPU path\n");
    return 1;
  }
  *pp = new B2Decoder();
  return 0;
}
void WelsDestroyDecoder(ISVCDecoder* p) { delete static_cast<B2Decoder*>(p); }

int WelsGetDecoderCapability(SDecoderCapability* c) {
  if (!c) return 1;
  memset(c, 0, sizeof(*c));
  // the same report as the reference (welsDecoderExt.cpp:1404-1417: Baseline, level 3.2 limits)
  c->iProfileIdc = 66;
  c->iProfileIop = 0xE0;
  c->iLevelIdc = 32;
  c->iMaxMbps = 216000;
  c->iMaxFs = 5120;
  c->iMaxCpb = 20000;
  c->iMaxDpb = 20480;
  c->iMaxBr = 20000;
  c->bRedPicCap = 0;
  return 0;
}

}  // extern "C"
