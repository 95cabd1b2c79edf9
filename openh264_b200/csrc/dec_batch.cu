// dec_batch.cu — layer-2 batched decoder: S independent streams, one access unit per stream per call.
// Host: Annex-B parsing (h264_parse.cpp, like the reference parses on the host: codec/decoder/core/src/au_parser.cpp,
// parse_mb_syn_cavlc.cpp), device: reconstruction of every macroblock by one warp (dec_mb.cuh, k_decode_mbs),
// in-loop deblocking and border replication (the encoder's kernels).  Replaces, for the supported stream class,
// CWelsDecoder::DecodeFrameNoDelay -> WelsDecodeBs -> DecodeCurrentAccessUnit -> WelsTargetSliceConstruction /
// WelsTargetMbConstruction (codec/decoder/core/src/decoder_core.cpp, decode_slice.cpp:~100) and
// WelsDeblockingFilterSlice (deblocking.cpp).  Streams outside the class are rejected with the parser's error
// code; there is no CPU reconstruction path in this library.
// Synchronous per call (parse -> H2D -> kernels -> D2H); streams without a slice in their access unit sit the call out.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>

#include <vector>

#include "../../include/b2h264_codec.h"
#include "b2h264_internal.h"
#include "enc_host.h"
#include "enc_launch.h"
#include "h264_parse.h"
#include "host_pool.h"

using namespace b2h264;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)

struct b2h264_dec {
  b2h264_dec_config cfg;
  int S = 0, mb_w = 0, mb_h = 0, n_mb = 0;
  StreamCtl geo;                          // picture geometry (strides, padded rows)
  std::vector<ParserState> parser;
  std::vector<ParsedPicture> parsed;      // per stream: the access unit of the running call
  std::vector<int> parse_rc;
  Pool* pool = nullptr;                   // the streams are parsed side by side (the parser is the decoder's host cost)
  int slots = 2;                          // picture slots per stream (num_ref_frames + 1 of the most demanding stream so far)
  std::vector<int> out_cx, out_cy;        // per active stream: luma samples cropped at the left / top of the output
  std::vector<int> out_slot;              // per active stream: slot of the picture decoded in this call
  std::vector<int> act;
  std::vector<uint8_t> act_ref;           // is the picture of act[i] a reference picture
  cudaStream_t st = nullptr;
  uint8_t* d_pic = nullptr;               // S x slots padded pictures: picture (s, slot) at (s * slots + slot) * pic_bytes
  MbInfo* d_mbi = nullptr;
  MbOut* d_recs = nullptr;
  MbOut* h_recs = nullptr;                // pinned: COMPACT records (pack_records_compact), one worst-case slot per stream
  uint8_t* d_pack = nullptr;              // device copy of the compact records (same slots)
  int32_t* h_idx = nullptr;               // pinned: per macroblock offset of its compact record or -1 - qp (skipped)
  int32_t* d_idx = nullptr;
  std::vector<int> units;                 // 32-byte units packed per active stream
  DecMbAux* d_aux = nullptr;
  DecMbAux* h_aux = nullptr;              // pinned
  DecMbAuxB* d_aux_b = nullptr;           // list 1 of B macroblocks: allocated when the first picture with B slices arrives
  DecMbAuxB* h_aux_b = nullptr;           // pinned
  std::vector<int> last_poc, last_idr, last_reorder;   // per stream: picture order count / IDR flag / reorder depth of the last decoded picture
  StreamFrame* d_sf = nullptr;
  StreamFrame* h_sf = nullptr;            // pinned
  int* d_ws = nullptr;
  size_t pic_bytes = 0, pic_y_bytes = 0, pic_c_bytes = 0;
  int last_error_stream = -1;

  uint8_t* plane0(int slot, int s, int pl) const {
    const int sty = geo.rec_stride_y(), stc = geo.rec_stride_c();
    uint8_t* base = d_pic + ((size_t)s * slots + slot) * pic_bytes;
    if (pl == 0) return base + (size_t)32 * sty + 32;
    return base + pic_y_bytes + (size_t)(pl - 1) * pic_c_bytes + (size_t)16 * stc + 16;
  }
};

extern "C" {

int b2h264_dec_create(const b2h264_dec_config* cfg, b2h264_dec** out) {
  if (!cfg || !out) return -1;
  if (cfg->width < 16 || cfg->height < 16 || (cfg->width & 1) || (cfg->height & 1) || cfg->n_streams < 1) return -2;
  int rc = b2h264_init(cfg->device);
  if (rc) return rc;
  if ((rc = enc_upload_deblock_tables())) return rc;
  b2h264_dec* d = new b2h264_dec();
  d->cfg = *cfg;
  d->S = cfg->n_streams;
  d->geo.init(cfg->width, cfg->height, 26, 30.0f, 0);
  d->mb_w = d->geo.sp.mb_w; d->mb_h = d->geo.sp.mb_h; d->n_mb = d->mb_w * d->mb_h;
  d->parser.resize(d->S);
  d->parsed.resize(d->S);
  d->parse_rc.assign(d->S, 0);
  { int nt = usable_cores(); if (nt > d->S) nt = d->S; d->pool = new Pool(nt < 1 ? 1 : nt); }
  d->pic_y_bytes = (size_t)d->geo.rec_stride_y() * d->geo.rec_rows_y();
  d->pic_c_bytes = (size_t)d->geo.rec_stride_c() * d->geo.rec_rows_c();
  d->pic_bytes = d->pic_y_bytes + 2 * d->pic_c_bytes;
  const size_t S = d->S;
  CK(cudaStreamCreateWithFlags(&d->st, cudaStreamNonBlocking));
  CK(cudaMalloc(&d->d_pic, S * d->slots * d->pic_bytes + 256));
  CK(cudaMemset(d->d_pic, 0, S * d->slots * d->pic_bytes + 256));
  CK(cudaMalloc(&d->d_mbi, S * d->n_mb * sizeof(MbInfo)));
  CK(cudaMemset(d->d_mbi, 0, S * d->n_mb * sizeof(MbInfo)));
  CK(cudaMalloc(&d->d_recs, S * d->n_mb * sizeof(MbOut)));
  CK(cudaMallocHost(&d->h_recs, S * d->n_mb * sizeof(MbOut)));
  CK(cudaMalloc(&d->d_pack, S * d->n_mb * sizeof(MbOut)));
  CK(cudaMallocHost(&d->h_idx, S * d->n_mb * sizeof(int32_t)));
  CK(cudaMalloc(&d->d_idx, S * d->n_mb * sizeof(int32_t)));
  d->units.assign(d->S, 0);
  d->last_poc.assign(d->S, 0); d->last_idr.assign(d->S, 0); d->last_reorder.assign(d->S, 0);
  CK(cudaMalloc(&d->d_aux, S * d->n_mb * sizeof(DecMbAux)));
  CK(cudaMallocHost(&d->h_aux, S * d->n_mb * sizeof(DecMbAux)));
  CK(cudaMalloc(&d->d_sf, S * sizeof(StreamFrame)));
  CK(cudaMallocHost(&d->h_sf, S * sizeof(StreamFrame)));
  CK(cudaMalloc(&d->d_ws, dec_sched_ints((int)S, d->n_mb) * sizeof(int)));
  *out = d;
  return 0;
}

void b2h264_dec_destroy(b2h264_dec* d) {
  if (!d) return;
  cudaSetDevice(d->cfg.device);
  if (d->st) cudaStreamSynchronize(d->st);
  cudaFree(d->d_pic);
  cudaFree(d->d_mbi); cudaFree(d->d_recs); cudaFree(d->d_aux); cudaFreeHost(d->h_aux); cudaFree(d->d_sf); cudaFree(d->d_ws);
  if (d->d_aux_b) cudaFree(d->d_aux_b);
  if (d->h_aux_b) cudaFreeHost(d->h_aux_b);
  cudaFreeHost(d->h_recs); cudaFreeHost(d->h_sf); cudaFree(d->d_pack); cudaFreeHost(d->h_idx); cudaFree(d->d_idx);
  if (d->st) cudaStreamDestroy(d->st);
  delete d->pool;
  delete d;
}

// got_picture (may be NULL): per stream 1 = a picture was decoded into yuv[s], 0 = the access unit carried no slice
// (parameter sets only / au[s] == NULL).  With got_picture == NULL an access unit without a slice is an error.
// status (may be NULL): per-stream outcome for callers that batch UNRELATED streams (the ISVCDecoder broker): 1 = picture, 0 = no slice in
// the unit, < 0 = this stream's error (-101 .. -105, -2) — such a stream sits the batch out and the others are decoded; without
// `status` the first stream error ends the call (the streams of one application stand or fall together).
static int dec_decode_impl(b2h264_dec* d, const uint8_t* const* au, const int32_t* au_bytes, uint8_t* const* yuv, int32_t* got_picture, int32_t* status) {
  if (!d || !au || !au_bytes || !yuv) return -1;
  CK(cudaSetDevice(d->cfg.device));
  const int S = d->S;
  int deblock = 1;
  static const bool timing = getenv("B2H264_DEC_TIMING") != nullptr;          // debug: host phases of the call on stderr
  const auto t0 = std::chrono::steady_clock::now();
  auto ms_since = [&](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count(); };
  d->act.clear(); d->act_ref.clear(); d->out_slot.clear(); d->out_cx.clear(); d->out_cy.clear();
  // 1. parse: every stream's access unit on the pool (streams are independent; a stream's parser state is its own)
  {
    std::function<void(int)> job = [&](int s) {
      ParsedPicture& pp = d->parsed[s];               // a fresh picture that keeps the record arrays' storage
      std::vector<MbOut> m = std::move(pp.mbs);
      std::vector<DecMbAux> a = std::move(pp.aux);
      std::vector<DecMbAuxB> ab = std::move(pp.aux_b);
      std::vector<CabacMbInfo> ci = std::move(pp.cabac_info);
      pp = ParsedPicture();
      pp.mbs = std::move(m); pp.aux = std::move(a); pp.aux_b = std::move(ab); pp.cabac_info = std::move(ci);
      d->parse_rc[s] = (!au[s] || au_bytes[s] <= 0) ? PARSE_NO_PICTURE : parse_access_unit(au[s], (size_t)au_bytes[s], &d->parser[s], &pp);
    };
    d->pool->run(S, job);
  }
  const double t_parse = ms_since(t0);
  // 2. the batch: descriptors and records of the streams that carry a picture, in stream order
#define STREAM_FAIL(code) { d->last_error_stream = s; if (status) { status[s] = (code); continue; } return (code); }
  if (status) for (int s = 0; s < S; s++) status[s] = 0;
  for (int s = 0; s < S; s++) {
    if (got_picture) got_picture[s] = 0;
    if (!au[s] || au_bytes[s] <= 0) {
      if (!got_picture && !status) { d->last_error_stream = s; return -1; }
      continue;
    }
    ParsedPicture& pic = d->parsed[s];
    const int rc = d->parse_rc[s];
    if (rc == PARSE_NO_PICTURE && (got_picture || status)) continue;
    if (rc != PARSE_OK) STREAM_FAIL(-100 + (rc == PARSE_NO_PICTURE ? PARSE_INVALID : rc))   // -101 truncated, -102 unsupported, -103 invalid, -104 no parameter sets, -105 picture incomplete
    const StreamParams& sp = d->parser[s].sp;
    if (sp.mb_w != d->mb_w || sp.mb_h != d->mb_h || sp.width != d->cfg.width || sp.height != d->cfg.height) STREAM_FAIL(-2)
    if ((int)pic.mbs.size() != d->n_mb) STREAM_FAIL(-103)
    if ((int)pic.aux.size() != d->n_mb) STREAM_FAIL(-103)
    if (pic.has_b && (int)pic.aux_b.size() != d->n_mb) STREAM_FAIL(-103)
    if (pic.n_slots > d->slots) {                   // a stream with more reference frames: widen every stream's slot array, keeping the pictures
      if (pic.n_slots > 17) STREAM_FAIL(-103)
      uint8_t* np = nullptr;
      CK(cudaMalloc(&np, (size_t)S * pic.n_slots * d->pic_bytes + 256));
      CK(cudaMemsetAsync(np, 0, (size_t)S * pic.n_slots * d->pic_bytes + 256, d->st));
      for (int q = 0; q < S; q++)
        CK(cudaMemcpyAsync(np + (size_t)q * pic.n_slots * d->pic_bytes, d->d_pic + (size_t)q * d->slots * d->pic_bytes,
                           (size_t)d->slots * d->pic_bytes, cudaMemcpyDeviceToDevice, d->st));
      CK(cudaStreamSynchronize(d->st));
      cudaFree(d->d_pic);
      d->d_pic = np; d->slots = pic.n_slots;
      for (int q = 0; q < (int)d->act.size(); q++) {      // descriptors built before the move
        StreamFrame& G = d->h_sf[q];
        const int qs = d->act[q];
        for (int pl = 0; pl < 3; pl++) { G.f.rec[pl] = d->plane0(d->out_slot[q], qs, pl); G.f.ref[pl] = G.f.dpb0[pl] = d->plane0(0, qs, pl); }
      }
    }
    if (d->act.empty()) deblock = 0;
    if (pic.any_deblock) deblock = 1;               // the filter kernel runs if any slice of any stream wants it (per-MB control inside)
    const int i = (int)d->act.size();
    d->act.push_back(s); d->act_ref.push_back(pic.is_ref ? 1 : 0); d->out_slot.push_back(pic.cur_slot); d->out_cx.push_back(pic.crop_left); d->out_cy.push_back(pic.crop_top);
    StreamFrame& F = d->h_sf[i];
    memset(&F, 0, sizeof(F));
    F.p.mb_w = d->mb_w; F.p.mb_h = d->mb_h;
    F.p.rec_stride_y = d->geo.rec_stride_y(); F.p.rec_stride_c = d->geo.rec_stride_c();
    F.p.dec_cqp_off = pic.chroma_qp_offset;
    F.p.qp = pic.ss.qp; F.p.is_idr = pic.ss.idr; F.p.ref_is_p = !pic.ss.idr; F.p.mv_range = 64; F.p.dec_mode = 1;
    if (pic.cur_slot >= d->slots) STREAM_FAIL(-103)
    for (int pl = 0; pl < 3; pl++) { F.f.rec[pl] = d->plane0(pic.cur_slot, s, pl); F.f.ref[pl] = F.f.dpb0[pl] = d->plane0(0, s, pl); }
    F.f.dpb_stride = (int64_t)d->pic_bytes;
    F.f.mbi = d->d_mbi + (size_t)s * d->n_mb;
    d->last_poc[s] = pic.poc; d->last_idr[s] = (pic.ss.idr ? 1 : 0) | (pic.has_b ? 2 : 0); d->last_reorder[s] = pic.max_reorder;
    if (pic.has_b) {                                // list-1 records of the picture's B macroblocks travel beside the aux records
      if (!d->d_aux_b) {
        CK(cudaMalloc(&d->d_aux_b, (size_t)S * d->n_mb * sizeof(DecMbAuxB)));
        CK(cudaMallocHost(&d->h_aux_b, (size_t)S * d->n_mb * sizeof(DecMbAuxB)));
      }
      F.f.dec_aux_b = d->d_aux_b + (size_t)i * d->n_mb;
    }
  }
  const int n = (int)d->act.size();
  if (n == 0) return 0;
  {
    std::function<void(int)> job = [&](int i) {          // compact records + index table into the stream's pinned slot
      const ParsedPicture& pic = d->parsed[d->act[i]];
      d->units[i] = pack_records_compact(pic.mbs.data(), d->n_mb, reinterpret_cast<uint8_t*>(d->h_recs + (size_t)i * d->n_mb), d->h_idx + (size_t)i * d->n_mb);
      memcpy(d->h_aux + (size_t)i * d->n_mb, pic.aux.data(), (size_t)d->n_mb * sizeof(DecMbAux));
      if (pic.has_b) memcpy(d->h_aux_b + (size_t)i * d->n_mb, pic.aux_b.data(), (size_t)d->n_mb * sizeof(DecMbAuxB));
    };
    d->pool->run(n, job);
  }
  const double t_pack = ms_since(t0);
  for (int i = 0; i < n; i++)                            // only what was packed crosses PCIe
    if (d->units[i] > 0)
      CK(cudaMemcpyAsync(d->d_pack + (size_t)i * d->n_mb * sizeof(MbOut), d->h_recs + (size_t)i * d->n_mb, (size_t)d->units[i] * 32, cudaMemcpyHostToDevice, d->st));
  CK(cudaMemcpyAsync(d->d_idx, d->h_idx, (size_t)n * d->n_mb * sizeof(int32_t), cudaMemcpyHostToDevice, d->st));
  { const int rcu = dec_launch_unpack(d->d_pack, (size_t)d->n_mb * sizeof(MbOut), d->d_idx, d->d_recs, n, d->n_mb, d->st); if (rcu) return rcu; }
  CK(cudaMemcpyAsync(d->d_aux, d->h_aux, (size_t)n * d->n_mb * sizeof(DecMbAux), cudaMemcpyHostToDevice, d->st));
  for (int i = 0; i < n; i++)
    if (d->parsed[d->act[i]].has_b)
      CK(cudaMemcpyAsync(d->d_aux_b + (size_t)i * d->n_mb, d->h_aux_b + (size_t)i * d->n_mb, (size_t)d->n_mb * sizeof(DecMbAuxB), cudaMemcpyHostToDevice, d->st));
  CK(cudaMemcpyAsync(d->d_sf, d->h_sf, (size_t)n * sizeof(StreamFrame), cudaMemcpyHostToDevice, d->st));
  int b_slices = 0;
  for (int i = 0; i < n; i++) b_slices |= (d->parsed[d->act[i]].has_b || d->parsed[d->act[i]].has_t8) ? 1 : 0;   // Main / High tools: the wider filter kernel
  const int rc = dec_launch_frame(d->d_sf, n, d->mb_w, d->mb_h, d->d_ws, d->d_recs, d->d_aux, deblock, d->st, b_slices);
  if (rc) return rc;
  const double t_launch = ms_since(t0);
  if (timing) cudaStreamSynchronize(d->st);
  const double t_kernels = ms_since(t0);
  const int w = d->cfg.width, h = d->cfg.height;
  for (int i = 0; i < n; i++) {
    const int s = d->act[i], rec = d->out_slot[i];
    uint8_t* dst = yuv[s];
    for (int pl = 0; pl < 3; pl++) {
      const int pw = pl ? w / 2 : w, ph = pl ? h / 2 : h;
      const int stp = pl ? d->geo.rec_stride_c() : d->geo.rec_stride_y();
      const uint8_t* from = d->plane0(rec, s, pl) + (size_t)(d->out_cy[i] >> (pl ? 1 : 0)) * stp + (d->out_cx[i] >> (pl ? 1 : 0));
      CK(cudaMemcpy2DAsync(dst, pw, from, stp, pw, ph, cudaMemcpyDeviceToHost, d->st));
      dst += (size_t)pw * ph;
    }
    if (got_picture) got_picture[s] = 1;
    if (status) status[s] = 1;
  }
  CK(cudaStreamSynchronize(d->st));
  if (timing) {
    long long units = 0;
    for (int i = 0; i < n; i++) units += d->units[i];
    fprintf(stderr, "b2h264_dec_decode: %d pictures: parse %.2f ms, pack %.2f, launches issued %.2f, kernels done %.2f, pictures on the host %.2f; %.2f MB of records\n", n, t_parse,
            t_pack - t_parse, t_launch - t_pack, t_kernels - t_launch, ms_since(t0) - t_kernels, units * 32 / 1e6);
  }
  return 0;
}

#undef STREAM_FAIL

int b2h264_dec_decode2(b2h264_dec* d, const uint8_t* const* au, const int32_t* au_bytes, uint8_t* const* yuv, int32_t* got_picture) {
  return dec_decode_impl(d, au, au_bytes, yuv, got_picture, nullptr);
}
int b2h264_dec_decode3(b2h264_dec* d, const uint8_t* const* au, const int32_t* au_bytes, uint8_t* const* yuv, int32_t* status) {
  if (!status) return -1;
  return dec_decode_impl(d, au, au_bytes, yuv, nullptr, status);
}
int b2h264_dec_decode(b2h264_dec* d, const uint8_t* const* au, const int32_t* au_bytes, uint8_t* const* yuv) {
  return dec_decode_impl(d, au, au_bytes, yuv, nullptr, nullptr);
}

// Looks at an access unit without a decoder: *has_slice = a coded slice NAL is present; if it carries an SPS of the
// supported class, *width / *height = the cropped picture size (else left untouched).  0, or a parser error (-10x).
int b2h264_dec_probe(const uint8_t* au, int32_t au_bytes, int32_t* width, int32_t* height, int32_t* has_slice) {
  if (!au || au_bytes <= 0) return -1;
  int w = 0, h = 0, sl = 0;
  const int rc = probe_access_unit(au, (size_t)au_bytes, &w, &h, &sl);
  if (has_slice) *has_slice = sl;
  if (rc != PARSE_OK) return -100 + rc;
  if (w > 0 && width) *width = w;
  if (h > 0 && height) *height = h;
  return 0;
}

// Output order: the decoder hands every picture back in DECODING order; *poc is its picture order count inside its coded video sequence
// (*flags: bit 0 = IDR picture, a new sequence starts; bit 1 = the picture holds B slices), *reorder_depth how many later-decoded pictures may precede it on output (0: the stream class has no
// reordering — Baseline — and decoding order is output order).  Layer 3 (ISVCDecoder) reorders with these.
int b2h264_dec_last_picture_order(b2h264_dec* d, int stream, int32_t* poc, int32_t* flags, int32_t* reorder_depth) {
  if (!d || stream < 0 || stream >= d->S) return -1;
  if (poc) *poc = d->last_poc[stream];
  if (flags) *flags = d->last_idr[stream];
  if (reorder_depth) *reorder_depth = d->last_reorder[stream];
  return 0;
}

// a stream slot starts over (a new ISVCDecoder object takes it): parser state, parameter sets and reference pictures are forgotten
int b2h264_dec_reset_stream(b2h264_dec* d, int stream) {
  if (!d || stream < 0 || stream >= d->S) return -1;
  d->parser[stream] = ParserState();
  d->parsed[stream] = ParsedPicture();
  return 0;
}

void* b2h264_host_alloc(size_t bytes) {
  void* p = nullptr;
  return cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) == cudaSuccess ? p : nullptr;
}
void b2h264_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"
