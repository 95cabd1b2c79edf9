// enc_kernels.cu — frame-level CUDA kernels of the encoder path (sm_100a):
//   k_pad_source     source I420 -> MB-aligned current picture
//   k_encode_mbs     macroblock mode decision + coding + reconstruction, ONE WARP PER MACROBLOCK; the MBs of
//                    all streams of the batch are scheduled by dependency (ready list, persistent warps)
//   k_deblock_mbs    in-loop deblocking, same scheduling
//   k_expand_*       border replication of the new reference picture (ExpandReferencingPicture)
// Why a wavefront: MB(x,y) needs the FINAL state of (x-1,y), (x,y-1), (x+1,y-1) — motion-vector / SAD
// predictors, intra neighbours, skip context (SURVEY.md §7 hard part 1); bit-exactness forbids breaking
// that chain, so parallelism comes from rows (2-MB lag) x independent streams of the batch.
#include <stdlib.h>

#include "b2h264_internal.h"
#define B2H264_WITH_INTER 1
#include "enc_deblock.cuh"
#include "enc_frame.cuh"
#include "enc_launch.h"

using namespace mbk;

namespace mbk {
__constant__ uint8_t c_alpha[52];
__constant__ uint8_t c_beta[52];
__constant__ uint8_t c_tc0[52][3];
}  // namespace mbk

#ifndef ENC_WPC
#define ENC_WPC 16         // warps per CTA of the macroblock kernels (one CTA per SM at 128 registers)
#endif

__device__ __forceinline__ int ld_volatile(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

// ---- source padding (CWelsPreProcess::Padding, wels_preprocess.cpp:1250) --------------------------------
__global__ void k_pad_source(const StreamFrame* __restrict__ sf, const uint8_t* const* __restrict__ src, int w, int h) {
  const StreamFrame& F = sf[blockIdx.z];
  const uint8_t* yuv = src[blockIdx.z];
  const int W = F.p.mb_w * 16, H = F.p.mb_h * 16;
  const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  uchar4 v = make_uchar4(0, 0, 0, 0);
  if (y < h) {
    const uint8_t* r = yuv + (size_t)y * w;
    if (x + 3 < w) v = *reinterpret_cast<const uchar4*>(r + x);          // w % 4 == 0 is required at create time
  }
  *reinterpret_cast<uchar4*>(const_cast<uint8_t*>(F.f.cur[0]) + (size_t)y * F.p.cur_stride_y + x) = v;
  if (y < H / 2 && x < W / 2) {
    const int cw = w / 2, ch = h / 2;
    for (int pl = 0; pl < 2; pl++) {
      uchar4 c = make_uchar4(0x80, 0x80, 0x80, 0x80);
      const uint8_t* p = yuv + (size_t)w * h + (size_t)pl * cw * ch + (size_t)y * cw;
      if (y < ch) {
        uint8_t t[4];
        for (int i = 0; i < 4; i++) t[i] = (x + i < cw) ? p[x + i] : 0x80;
        c = make_uchar4(t[0], t[1], t[2], t[3]);
      }
      *reinterpret_cast<uchar4*>(const_cast<uint8_t*>(F.f.cur[1 + pl]) + (size_t)y * F.p.cur_stride_c + x) = c;
    }
  }
}

// ---- dependency-driven macroblock scheduling -------------------------------------------------------------
// MB(x,y) may start when (x-1,y) and (x+1,y-1) [(x,y-1) on the last column] are final; those two imply the
// top and top-left neighbours.  Every finished MB notifies its right neighbour and its bottom-left
// neighbour; the second notification pushes the MB onto a global ready list.  Persistent warps pop the
// list: no warp ever holds an SM slot while a row it depends on is still busy (a row-per-warp wavefront
// spent >80% of its issue slots spinning, profiles/r01_encode_rows_ncu.txt).
struct Sched {
  int* dep;        // per (stream, mb): notifications received so far
  int* queue;      // ready list, capacity = n_streams * n_mb, -1 = not yet pushed
  int* head;       // next entry to pop
  int* tail;       // next free entry
};

__device__ __forceinline__ void sched_push(const Sched& q, int id) {
  const int slot = atomicAdd(q.tail, 1);
  *reinterpret_cast<volatile int*>(q.queue + slot) = id;
}
__device__ __forceinline__ void sched_notify(const Sched& q, int id, int need) {
  if (atomicAdd(q.dep + id, 1) + 1 == need) sched_push(q, id);
}

template <class Body>
__device__ __forceinline__ void run_mbs(const StreamFrame* sf, int n_streams, const Sched q, Body body) {
  const int lane = threadIdx.x & 31;
  const int mb_w = sf[0].p.mb_w, mb_h = sf[0].p.mb_h, n_mb = mb_w * mb_h, total = n_streams * n_mb;
  for (;;) {
    int id = 0;
    if (lane == 0) {
      const int t = atomicAdd(q.head, 1);
      id = -1;
      if (t < total) {
        while ((id = ld_volatile(q.queue + t)) < 0) __nanosleep(100);
      }
    }
    id = __shfl_sync(MBK_FULL, id, 0);
    if (id < 0) break;
    __threadfence();
    const int si = id / n_mb, mb = id - si * n_mb, y = mb / mb_w, x = mb - y * mb_w;
    body(sf[si], x, y);
    __threadfence();
    __syncwarp();
    if (lane == 0) {
      if (x + 1 < mb_w) sched_notify(q, id + 1, 1 + (y > 0));                    // right neighbour: we are its left
      if (y + 1 < mb_h) {
        if (x > 0) sched_notify(q, id + mb_w - 1, 1 + (x - 1 > 0));              // bottom-left: we are its top-right
        if (x == mb_w - 1) sched_notify(q, id + mb_w, 1 + (x > 0));              // last column: we are its top
      }
    }
  }
}

// CTA-synchronous variant: the CTA takes up to ENC_WPC macroblocks that are ALREADY ready, its warps start
// them together and meet again before the next batch.  Warps that start together run the same code at the
// same time, so one instruction-cache fill serves all of them.
template <class Body>
__device__ __forceinline__ void run_mbs_cta(const StreamFrame* sf, int n_streams, const Sched q, Body body) {
  __shared__ int s_base, s_n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int mb_w = sf[0].p.mb_w, mb_h = sf[0].p.mb_h, n_mb = mb_w * mb_h, total = n_streams * n_mb;
  for (;;) {
    if (threadIdx.x == 0) {
      int h, n;
      for (;;) {
        h = ld_volatile(q.head);
        if (h >= total) { n = -1; break; }
        const int t = ld_volatile(q.tail);
        n = min(ENC_WPC, t - h);
        if (n > 0 && atomicCAS(q.head, h, h + n) == h) break;
        __nanosleep(200);
      }
      s_base = h; s_n = n;
    }
    __syncthreads();
    const int base = s_base, n = s_n;
    if (n < 0) break;
    if (warp < n) {
      int id = 0;
      if (lane == 0) while ((id = ld_volatile(q.queue + base + warp)) < 0) {}
      id = __shfl_sync(MBK_FULL, id, 0);
      __threadfence();
      const int si = id / n_mb, mb = id - si * n_mb, y = mb / mb_w, x = mb - y * mb_w;
      body(sf[si], x, y);
      __threadfence();
      __syncwarp();
      if (lane == 0) {
        if (x + 1 < mb_w) sched_notify(q, id + 1, 1 + (y > 0));
        if (y + 1 < mb_h) {
          if (x > 0) sched_notify(q, id + mb_w - 1, 1 + (x - 1 > 0));
          if (x == mb_w - 1) sched_notify(q, id + mb_w, 1 + (x > 0));
        }
      }
    }
    __syncthreads();
  }
}

__global__ void k_sched_init(Sched q, int n_streams, int n_mb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_streams) q.queue[i] = i * n_mb;        // MB (0,0) of every stream is ready
  if (i == 0) { *q.head = 0; *q.tail = n_streams; }
}

// optional per-MB-type cycle statistics (debug): [type*2] = cycles, [type*2+1] = count
__device__ unsigned long long g_enc_stats[16];
#ifdef B2H264_PHASE_STATS
namespace mbk { __device__ unsigned long long g_phase[32]; }
extern "C" int b2h264_debug_phase_stats(unsigned long long* out32, int reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out32, mbk::g_phase, sizeof(mbk::g_phase));
  if (e == cudaSuccess && reset) { unsigned long long z[32] = {0}; e = cudaMemcpyToSymbol(mbk::g_phase, z, sizeof(z)); }
  return (int)e;
}
#endif

__global__ void __launch_bounds__(32 * ENC_WPC) k_encode_mbs(const StreamFrame* __restrict__ sf, int n_streams, Sched q, int stats) {
  extern __shared__ __align__(16) uint8_t smem[];
  MbScratch& s = reinterpret_cast<MbScratch*>(smem)[threadIdx.x >> 5];
#ifdef B2H264_WARP_ASYNC
  run_mbs(sf,
#else
  run_mbs_cta(sf,
#endif
 n_streams, q, [&](const StreamFrame& F, int x, int y) {
    const long long t0 = stats ? clock64() : 0;
    encode_one_mb(F.p, F.f, s, x, y);
    if (stats && (threadIdx.x & 31) == 0) {
      const int t = s.info.mb_type & 7;
      atomicAdd(&g_enc_stats[2 * t], (unsigned long long)(clock64() - t0));
      atomicAdd(&g_enc_stats[2 * t + 1], 1ull);
    }
  });
}

__global__ void __launch_bounds__(32 * ENC_WPC) k_deblock_mbs(const StreamFrame* __restrict__ sf, int n_streams, Sched q) {
  run_mbs(sf, n_streams, q, [&](const StreamFrame& F, int x, int y) { deblock_one_mb(F.p, F.f, x, y); });
}

// ---- border replication, all planes of all streams in two launches -------------------------------------------
__global__ void k_expand_lr_batch(const StreamFrame* __restrict__ sf) {
  const StreamFrame& F = sf[blockIdx.z / 3];
  const int pl = blockIdx.z % 3;
  const int pad = pl ? 16 : 32, st = pl ? F.p.rec_stride_c : F.p.rec_stride_y;
  const int w = (pl ? 8 : 16) * F.p.mb_w, h = (pl ? 8 : 16) * F.p.mb_h;
  const int y = blockIdx.x * blockDim.y + threadIdx.y;
  if (y >= h) return;
  uint8_t* row = F.f.rec[pl] + (size_t)y * st;
  const uint8_t l = row[0], r = row[w - 1];
  for (int x = threadIdx.x; x < pad; x += blockDim.x) { row[x - pad] = l; row[w + x] = r; }
}
__global__ void k_expand_tb_batch(const StreamFrame* __restrict__ sf) {
  const StreamFrame& F = sf[blockIdx.z / 3];
  const int pl = blockIdx.z % 3;
  const int pad = pl ? 16 : 32, st = pl ? F.p.rec_stride_c : F.p.rec_stride_y;
  const int w = (pl ? 8 : 16) * F.p.mb_w, h = (pl ? 8 : 16) * F.p.mb_h;
  const int x = blockIdx.x * blockDim.x + threadIdx.x - pad;
  if (x >= w + pad) return;
  uint8_t* pic = F.f.rec[pl];
  const uint8_t t = pic[x], b = pic[(size_t)(h - 1) * st + x];
  for (int y = 1 + blockIdx.y; y <= pad; y += gridDim.y) {
    pic[x - (ptrdiff_t)y * st] = t;
    pic[(size_t)(h - 1 + y) * st + x] = b;
  }
}

// ================================================================================================
// H.264 Tables 8-16 / 8-17 (host copies; the device reads the constant-memory twins)
namespace mbk {
uint8_t h_alpha[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 5, 6, 7, 8, 9, 10, 12, 13, 15, 17, 20, 22, 25, 28,
                       32, 36, 40, 45, 50, 56, 63, 71, 80, 90, 101, 113, 127, 144, 162, 182, 203, 226, 255, 255};
uint8_t h_beta[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 6, 6, 7, 7, 8, 8,
                      9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 14, 14, 15, 15, 16, 16, 17, 17, 18, 18};
uint8_t h_tc0[52][3] = {
    {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0},
    {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 1}, {0, 0, 1}, {0, 0, 1}, {0, 0, 1}, {0, 1, 1},
    {0, 1, 1}, {1, 1, 1}, {1, 1, 1}, {1, 1, 1}, {1, 1, 1}, {1, 1, 2}, {1, 1, 2}, {1, 1, 2}, {1, 1, 2}, {1, 2, 3}, {1, 2, 3},
    {2, 2, 3}, {2, 2, 4}, {2, 3, 4}, {2, 3, 4}, {3, 3, 5}, {3, 4, 6}, {3, 4, 6}, {4, 5, 7}, {4, 5, 8}, {4, 6, 9}, {5, 7, 10},
    {6, 8, 11}, {6, 8, 13}, {7, 10, 14}, {8, 11, 16}, {9, 12, 18}, {10, 13, 20}, {11, 15, 23}, {13, 17, 25}};
}  // namespace mbk

int enc_upload_deblock_tables() {
  cudaError_t e;
  if ((e = cudaMemcpyToSymbol(mbk::c_alpha, mbk::h_alpha, sizeof(mbk::h_alpha))) != cudaSuccess) return (int)e;
  if ((e = cudaMemcpyToSymbol(mbk::c_beta, mbk::h_beta, sizeof(mbk::h_beta))) != cudaSuccess) return (int)e;
  if ((e = cudaMemcpyToSymbol(mbk::c_tc0, mbk::h_tc0, sizeof(mbk::h_tc0))) != cudaSuccess) return (int)e;
  return 0;
}

static int g_enc_blocks = 0;
static int enc_grid_blocks() {
  if (g_enc_blocks) return g_enc_blocks;
  cudaFuncSetAttribute(k_encode_mbs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(MbScratch) * ENC_WPC));
  int dev = 0, sms = 0, per_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_encode_mbs, 32 * ENC_WPC, sizeof(MbScratch) * ENC_WPC);
  if (per_sm < 1) per_sm = 1;
  g_enc_blocks = sms * per_sm;          // persistent: exactly what the chip can hold
  return g_enc_blocks;
}

// scheduler workspace layout (ints): [0..3] head/tail for encode, deblock; then dep[2][total]; then queue[2][total]
size_t enc_sched_ints(int n_streams, int n_mb) { return 8 + 4 * (size_t)n_streams * n_mb; }

static Sched make_sched(int* ws, int which, int total) {
  Sched q;
  q.head = ws + 2 * which; q.tail = ws + 2 * which + 1;
  q.dep = ws + 8 + (size_t)which * total;
  q.queue = ws + 8 + (size_t)(2 + which) * total;
  return q;
}

int enc_launch_frame(const StreamFrame* d_sf, const uint8_t* const* d_src, int n_streams, int w, int h, int mb_w, int mb_h,
                     int* d_ws, cudaStream_t st) {
  int rc;
  if (d_src) {
    dim3 b(32, 8), g((mb_w * 4 + 31) / 32, (mb_h * 16 + 7) / 8, n_streams);
    k_pad_source<<<g, b, 0, st>>>(d_sf, d_src, w, h);
    if ((rc = b2h264_launched())) return rc;
  }
  const int total = n_streams * mb_w * mb_h;
  cudaMemsetAsync(d_ws + 8, 0, 2 * (size_t)total * sizeof(int), st);                               // dep counters
  cudaMemsetAsync(d_ws + 8 + 2 * (size_t)total, 0xff, 2 * (size_t)total * sizeof(int), st);        // ready lists
  const Sched qe = make_sched(d_ws, 0, total), qd = make_sched(d_ws, 1, total);
  k_sched_init<<<(n_streams + 127) / 128, 128, 0, st>>>(qe, n_streams, mb_w * mb_h);
  k_sched_init<<<(n_streams + 127) / 128, 128, 0, st>>>(qd, n_streams, mb_w * mb_h);
  int blocks = enc_grid_blocks();
  const int need = (total + ENC_WPC - 1) / ENC_WPC;
  if (blocks > need) blocks = need;
  static const int stats = getenv("B2H264_ENC_STATS") ? 1 : 0;
  k_encode_mbs<<<blocks, 32 * ENC_WPC, sizeof(MbScratch) * ENC_WPC, st>>>(d_sf, n_streams, qe, stats);
  if ((rc = b2h264_launched())) return rc;
  return 0;
}

int enc_launch_deblock_expand(const StreamFrame* d_sf, int n_streams, int mb_w, int mb_h, int* d_ws, cudaStream_t st) {
  int rc;
  const int total = n_streams * mb_w * mb_h;
  int blocks = enc_grid_blocks() * 3;
  const int need = (total + ENC_WPC - 1) / ENC_WPC;
  if (blocks > need) blocks = need;
  k_deblock_mbs<<<blocks, 32 * ENC_WPC, 0, st>>>(d_sf, n_streams, make_sched(d_ws, 1, total));
  if ((rc = b2h264_launched())) return rc;
  k_expand_lr_batch<<<dim3((mb_h * 16 + 7) / 8, 1, 3 * n_streams), dim3(32, 8), 0, st>>>(d_sf);
  if ((rc = b2h264_launched())) return rc;
  k_expand_tb_batch<<<dim3((mb_w * 16 + 64 + 127) / 128, 4, 3 * n_streams), dim3(128), 0, st>>>(d_sf);
  return b2h264_launched();
}

size_t enc_scratch_bytes() { return sizeof(MbScratch); }

extern "C" int b2h264_debug_enc_stats(unsigned long long* out16, int reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out16, g_enc_stats, sizeof(g_enc_stats));
  if (e != cudaSuccess) return (int)e;
  if (reset) { unsigned long long z[16] = {0}; e = cudaMemcpyToSymbol(g_enc_stats, z, sizeof(z)); }
  return (int)e;
}
