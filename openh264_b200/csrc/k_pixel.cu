// k_pixel.cu — batched pixel-domain kernels behind include/b2h264.h: SAD/SATD, sub-pel MC, deblocking
// edge filters, border expansion, integer motion search, and the MC+SAD roofline unit.
// One warp per job; 8 warps per CTA; grid sized to cover n jobs.
#include "b2h264_internal.h"
#include "mbk_deblock.cuh"
#include "mbk_mc.cuh"
#include "mbk_me.cuh"
#include "mbk_sad.cuh"

using namespace mbk;

#define WARPS_PER_CTA 8
#define JOB_GRID(n) dim3(((n) + WARPS_PER_CTA - 1) / WARPS_PER_CTA), dim3(32 * WARPS_PER_CTA)
__device__ __forceinline__ int warp_job() { return blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5); }

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sad(const uint8_t* __restrict__ a, int sa, const int32_t* __restrict__ a_off,
                                             const uint8_t* __restrict__ b, int sb, const int32_t* __restrict__ b_off, int blk,
                                             int n, int32_t* sad, int32_t* satd, int32_t* sad4) {
  const int j = warp_job();
  if (j >= n) return;
  const uint8_t* pa = a + a_off[j];
  const uint8_t* pb = b + b_off[j];
  const int lw = blk_lw(blk), lh = blk_lh(blk);
  if (sad) { const int v = warp_sad(pa, sa, pb, sb, lw, lh); if (lane_id() == 0) sad[j] = v; }
  if (satd) { const int v = warp_satd(pa, sa, pb, sb, lw, lh); if (lane_id() == 0) satd[j] = v; }
  if (sad4) {
    int s[4];
    warp_sad_four(pa, sa, pb, sb, lw, lh, s);
    if (lane_id() < 4) sad4[4 * j + lane_id()] = s[lane_id()];   // s[] is warp-uniform
  }
}

__global__ void __launch_bounds__(256) k_mc_luma(const uint8_t* __restrict__ src, int ss, const int32_t* __restrict__ off,
                                                 const int16_t* __restrict__ mv, int w, int h, int n, uint8_t* dst) {
  const int j = warp_job();
  if (j >= n) return;
  warp_mc_luma(src + off[j], ss, dst + 256 * j, 16, mv[2 * j], mv[2 * j + 1], w, h);
}
__global__ void __launch_bounds__(256) k_mc_chroma(const uint8_t* __restrict__ src, int ss, const int32_t* __restrict__ off,
                                                   const int16_t* __restrict__ mv, int w, int h, int n, uint8_t* dst) {
  const int j = warp_job();
  if (j >= n) return;
  warp_mc_chroma(src + off[j], ss, dst + 64 * j, 8, mv[2 * j], mv[2 * j + 1], w, h);
}
__global__ void __launch_bounds__(256) k_halfpel(int which, const uint8_t* __restrict__ src, int ss,
                                                 const int32_t* __restrict__ off, int w, int h, int n, uint8_t* dst) {
  const int j = warp_job();
  if (j >= n) return;
  warp_halfpel(which, src + off[j], ss, dst + 289 * j, 17, w, h);
}
__global__ void __launch_bounds__(256) k_pixel_avg(const uint8_t* __restrict__ a, int sa, const int32_t* __restrict__ a_off,
                                                   const uint8_t* __restrict__ b, int sb, const int32_t* __restrict__ b_off,
                                                   int w, int h, int n, uint8_t* dst) {
  const int j = warp_job();
  if (j >= n) return;
  warp_pixel_avg(dst + 256 * j, 16, a + a_off[j], sa, b + b_off[j], sb, w, h);
}

// ------------------------------------------------------------------------------------------------
// deblocking: one warp per edge job, one lane per pixel line (16 luma lines / 8 chroma lines x 2 planes)
__global__ void __launch_bounds__(256) k_deblock_luma(uint8_t* pic, const b2h264_edge_job* __restrict__ jobs, int n) {
  const int j = warp_job();
  if (j >= n) return;
  const b2h264_edge_job e = jobs[j];
  const int l = lane_id();
  if (l < 16) {
    uint8_t* p = pic + e.off + l * e.sy;
    if (e.strong) deblock_luma_eq4_line(p, e.sx, e.alpha, e.beta);
    else deblock_luma_lt4_line(p, e.sx, e.alpha, e.beta, e.tc[l >> 2]);
  }
}
__global__ void __launch_bounds__(256) k_deblock_chroma(uint8_t* cb, uint8_t* cr, const b2h264_edge_job* __restrict__ jobs,
                                                        int n) {
  const int j = warp_job();
  if (j >= n) return;
  const b2h264_edge_job e = jobs[j];
  const int l = lane_id();
  if (l < 16) {
    uint8_t* p = (l < 8 ? cb : cr) + e.off + (l & 7) * e.sy;
    if (e.strong) deblock_chroma_eq4_line(p, e.sx, e.alpha, e.beta);
    else deblock_chroma_lt4_line(p, e.sx, e.alpha, e.beta, e.tc[(l & 7) >> 1]);
  }
}

// ------------------------------------------------------------------------------------------------
// border replication: phase 1 left/right of every row, phase 2 top/bottom bands (incl. corners)
__global__ void k_expand_lr(uint8_t* pic, int stride, int w, int h, int pad) {
  const int y = blockIdx.x * blockDim.y + threadIdx.y;
  if (y >= h) return;
  uint8_t* row = pic + (size_t)y * stride;
  const uint8_t l = row[0], r = row[w - 1];
  for (int x = threadIdx.x; x < pad; x += blockDim.x) { row[x - pad] = l; row[w + x] = r; }
}
__global__ void k_expand_tb(uint8_t* pic, int stride, int w, int h, int pad) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x - pad;   // covers [-pad, w+pad)
  if (x >= w + pad) return;
  const uint8_t t = pic[x], b = pic[(size_t)(h - 1) * stride + x];
  for (int y = 1 + blockIdx.y; y <= pad; y += gridDim.y) {
    pic[x - (ptrdiff_t)y * stride] = t;
    pic[(size_t)(h - 1 + y) * stride + x] = b;
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_me_search(const uint8_t* __restrict__ cur, int cs, const uint8_t* __restrict__ ref,
                                                   int rs, const b2h264_me_job* __restrict__ jobs, int n,
                                                   b2h264_me_result* out) {
  const int j = warp_job();
  if (j >= n) return;
  const b2h264_me_job* jb = jobs + j;
  MeIn in;
  in.enc = cur + jb->cur_off; in.enc_stride = cs;
  in.ref = ref + jb->ref_off; in.ref_stride = rs;
  in.blk = jb->blk;
  in.mvp_x = jb->mvp_x; in.mvp_y = jb->mvp_y;
  in.min_x = jb->mv_min_x; in.min_y = jb->mv_min_y; in.max_x = jb->mv_max_x; in.max_y = jb->mv_max_y;
  in.n_mvc = jb->n_mvc; in.mvc = &jb->mvc[0][0];
  in.sad_pred = jb->sad_pred;
  in.lambda = c_lambda[jb->qp];
  in.calc_satd = jb->calc_satd != 0;
  MeOut o;
  warp_me_search(in, o);
  if (lane_id() == 0) {
    b2h264_me_result r;
    r.mv_x = (int16_t)o.mv_x; r.mv_y = (int16_t)o.mv_y;
    r.sad_cost = o.sad_cost; r.satd_cost = o.satd_cost;
    r.ref_off = (int32_t)(o.ref_best - ref);
    out[j] = r;
  }
}

// ------------------------------------------------------------------------------------------------
// MC + SAD roofline unit.  One warp per macroblock.  The warp stages the reference window that all
// of the macroblock's k candidates can touch ([-R-2, 16+R+3) in both axes, R = 8 integer pels) and
// the current macroblock into shared memory once (coalesced 4-byte loads of 128-B-aligned rows from
// HBM/L2), then evaluates every candidate from shared memory: quarter-pel interpolation fused with
// the SAD accumulation, no intermediate prediction buffer.  HBM traffic per MB = cur 256 B + its
// share of the reference plane (each reference byte is fetched from DRAM once; window overlaps
// between neighbouring MBs hit L2) + 4k bytes of costs  ->  ~512 B + 4k per MB algorithmic.
#define MCS_R 8
#define MCS_WIN (16 + 2 * MCS_R + 5)       // 37 rows/cols touched
#define MCS_WSTRIDE 44                     // smem row pitch: 37 + alignment slack, multiple of 4
__global__ void __launch_bounds__(256) k_mc_sad(const uint8_t* __restrict__ cur, int cs, const uint8_t* __restrict__ ref,
                                                int rs, int mb_w, int mb_h, const int16_t* __restrict__ mv, int k,
                                                int32_t* __restrict__ cost) {
  __shared__ __align__(16) uint8_t s_win[WARPS_PER_CTA][MCS_WIN * MCS_WSTRIDE + 16];
  __shared__ __align__(16) uint8_t s_cur[WARPS_PER_CTA][256];
  const int m = warp_job();
  if (m >= mb_w * mb_h) return;
  const int wi = threadIdx.x >> 5, l = lane_id();
  const int mbx = m % mb_w, mby = m / mb_w;
  uint8_t* win = s_win[wi];
  uint8_t* cm = s_cur[wi];
  // current MB: 16 rows x 16 B = 64 words, 2 per lane
  for (int i = l; i < 64; i += 32) {
    const int row = i >> 2, c4 = (i & 3) << 2;
    *reinterpret_cast<uint32_t*>(cm + row * 16 + c4) =
        *reinterpret_cast<const uint32_t*>(cur + (size_t)(mby * 16 + row) * cs + mbx * 16 + c4);
  }
  // reference window: origin (mbx*16 - R - 2, mby*16 - R - 2); load whole aligned words covering it
  const int ox = mbx * 16 - MCS_R - 2, oy = mby * 16 - MCS_R - 2;
  const uint8_t* wsrc = ref + (ptrdiff_t)oy * rs + ox;
  const int mis = (int)(reinterpret_cast<uintptr_t>(wsrc) & 3);     // same for every row when rs % 4 == 0
  const int words = (mis + MCS_WIN + 3) >> 2;                      // <= 11
  for (int i = l; i < MCS_WIN * words; i += 32) {
    const int row = i / words, wd = i - row * words;
    *reinterpret_cast<uint32_t*>(win + row * MCS_WSTRIDE + 4 * wd) =
        *reinterpret_cast<const uint32_t*>(wsrc - mis + (ptrdiff_t)row * rs + 4 * wd);
  }
  __syncwarp();
  const uint8_t* w0 = win + mis + (MCS_R + 2) * MCS_WSTRIDE + (MCS_R + 2);   // pixel (0,0) of the co-located block
  for (int c = 0; c < k; c++) {
    const int mvx = mv[(size_t)(m * k + c) * 2], mvy = mv[(size_t)(m * k + c) * 2 + 1];
    const uint8_t* p = w0 + (mvy >> 2) * MCS_WSTRIDE + (mvx >> 2);
    const int fx = mvx & 3, fy = mvy & 3;
    int s = 0;
#pragma unroll 2
    for (int i = l; i < 256; i += 32) {
      const int y = i >> 4, x = i & 15;
      s += iabs((int)cm[i] - luma_qpel_sample(p + y * MCS_WSTRIDE + x, MCS_WSTRIDE, fx, fy));
    }
    s = warp_sum(s);
    if (l == 0) cost[(size_t)m * k + c] = s;
  }
}

// ================================================================================================
// C-ABI
// ================================================================================================
extern "C" {

int b2h264_k_sad(const uint8_t* a, int sa, const int32_t* a_off, const uint8_t* b, int sb, const int32_t* b_off, int blk,
                 int n, int32_t* sad, int32_t* satd, int32_t* sad4, void* stream) {
  if (n <= 0) return 0;
  k_sad<<<JOB_GRID(n), 0, (cudaStream_t)stream>>>(a, sa, a_off, b, sb, b_off, blk, n, sad, satd, sad4);
  return b2h264_launched();
}
int b2h264_k_mc_luma(const uint8_t* src, int ss, const int32_t* off, const int16_t* mv, int w, int h, int n, uint8_t* dst,
                     void* stream) {
  if (n <= 0) return 0;
  k_mc_luma<<<JOB_GRID(n), 0, (cudaStream_t)stream>>>(src, ss, off, mv, w, h, n, dst);
  return b2h264_launched();
}
int b2h264_k_mc_chroma(const uint8_t* src, int ss, const int32_t* off, const int16_t* mv, int w, int h, int n,
                       uint8_t* dst, void* stream) {
  if (n <= 0) return 0;
  k_mc_chroma<<<JOB_GRID(n), 0, (cudaStream_t)stream>>>(src, ss, off, mv, w, h, n, dst);
  return b2h264_launched();
}
int b2h264_k_halfpel(int which, const uint8_t* src, int ss, const int32_t* off, int w, int h, int n, uint8_t* dst,
                     void* stream) {
  if (n <= 0) return 0;
  k_halfpel<<<JOB_GRID(n), 0, (cudaStream_t)stream>>>(which, src, ss, off, w, h, n, dst);
  return b2h264_launched();
}
int b2h264_k_pixel_avg(const uint8_t* a, int sa, const int32_t* a_off, const uint8_t* b, int sb, const int32_t* b_off,
                       int w, int h, int n, uint8_t* dst, void* stream) {
  if (n <= 0) return 0;
  k_pixel_avg<<<JOB_GRID(n), 0, (cudaStream_t)stream>>>(a, sa, a_off, b, sb, b_off, w, h, n, dst);
  return b2h264_launched();
}
int b2h264_k_deblock_luma(uint8_t* pic, const b2h264_edge_job* jobs, int n, void* stream) {
  if (n <= 0) return 0;
  k_deblock_luma<<<JOB_GRID(n), 0, (cudaStream_t)stream>>>(pic, jobs, n);
  return b2h264_launched();
}
int b2h264_k_deblock_chroma(uint8_t* cb, uint8_t* cr, const b2h264_edge_job* jobs, int n, void* stream) {
  if (n <= 0) return 0;
  k_deblock_chroma<<<JOB_GRID(n), 0, (cudaStream_t)stream>>>(cb, cr, jobs, n);
  return b2h264_launched();
}
int b2h264_k_expand_plane(uint8_t* pic, int stride, int w, int h, int pad, void* stream) {
  k_expand_lr<<<dim3((h + 7) / 8), dim3(32, 8), 0, (cudaStream_t)stream>>>(pic, stride, w, h, pad);
  int rc = b2h264_launched();
  if (rc) return rc;
  k_expand_tb<<<dim3((w + 2 * pad + 127) / 128, 4), dim3(128), 0, (cudaStream_t)stream>>>(pic, stride, w, h, pad);
  return b2h264_launched();
}
int b2h264_k_me_search(const uint8_t* cur, int cs, const uint8_t* ref, int rs, const b2h264_me_job* jobs, int n,
                       b2h264_me_result* out, void* stream) {
  if (n <= 0) return 0;
  k_me_search<<<JOB_GRID(n), 0, (cudaStream_t)stream>>>(cur, cs, ref, rs, jobs, n, out);
  return b2h264_launched();
}
int b2h264_k_mc_sad(const uint8_t* cur, int cs, const uint8_t* ref, int rs, int mb_w, int mb_h, const int16_t* mv, int k,
                    int32_t* cost, void* stream) {
  const int n = mb_w * mb_h;
  if (n <= 0 || k <= 0) return 0;
  if ((rs & 3) || (cs & 3)) return cudaErrorInvalidValue;
  k_mc_sad<<<JOB_GRID(n), 0, (cudaStream_t)stream>>>(cur, cs, ref, rs, mb_w, mb_h, mv, k, cost);
  return b2h264_launched();
}

}  // extern "C"
