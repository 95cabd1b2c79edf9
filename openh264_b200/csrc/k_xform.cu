// k_xform.cu — batched transform-domain kernels behind include/b2h264.h (one THREAD per 4x4 job unit):
// forward DCT, quantisation, Hadamard DC paths, zig-zag/score, dequantisation, IDCT + reconstruction.
#include "b2h264_internal.h"
#include "mbk_xform.cuh"

using namespace mbk;

#define TPB 128
#define TGRID(n) dim3(((n) + TPB - 1) / TPB), dim3(TPB)
__device__ __forceinline__ int tjob() { return blockIdx.x * TPB + threadIdx.x; }

__device__ __forceinline__ void ld16(int16_t d[16], const int16_t* p) {
#pragma unroll
  for (int i = 0; i < 16; i++) d[i] = p[i];
}
__device__ __forceinline__ void st16(int16_t* p, const int16_t d[16]) {
#pragma unroll
  for (int i = 0; i < 16; i++) p[i] = d[i];
}

// job = one 4x4 block; 4 consecutive jobs form one pfDctFourT4 call (z order inside the 8x8)
__global__ void k_dct_four4x4(const uint8_t* __restrict__ p1, int s1, const int32_t* __restrict__ o1,
                              const uint8_t* __restrict__ p2, int s2, const int32_t* __restrict__ o2, int n4, int16_t* dct) {
  const int t = tjob();
  if (t >= n4) return;
  const int j = t >> 2, k = t & 3, ox = (k & 1) * 4, oy = (k >> 1) * 4;
  int16_t d[16];
  dct4x4(d, p1 + o1[j] + oy * s1 + ox, s1, p2 + o2[j] + oy * s2 + ox, s2);
  st16(dct + 16 * t, d);
}
__global__ void k_quant_four4x4(int16_t* dct, int qp, int intra, int n4, int16_t* max4) {
  const int t = tjob();
  if (t >= n4) return;
  int16_t d[16];
  ld16(d, dct + 16 * t);
  const int16_t mx = quant4x4_max(d, c_quant_ff[qp + (intra ? 6 : 0)], c_quant_mf[qp]);
  st16(dct + 16 * t, d);
  if (max4) max4[t] = mx;
}
__global__ void k_quant4x4_dc(int16_t* dct, int ff, int mf, int n) {
  const int t = tjob();
  if (t >= n) return;
  int16_t d[16];
  ld16(d, dct + 16 * t);
  quant4x4_dc(d, ff, mf);
  st16(dct + 16 * t, d);
}
__global__ void k_hadamard_quant2x2(int16_t* rs, int ff, int mf, int n, int16_t* dct, int32_t* nz, int32_t* skip) {
  const int t = tjob();
  if (t >= n) return;
  int16_t* r = rs + 64 * t;
  const int16_t in[4] = {r[0], r[16], r[32], r[48]};
  if (skip) skip[t] = hadamard_quant2x2_skip(in, (int16_t)ff, (int16_t)mf);
  int16_t out[4];
  const int z = hadamard_quant2x2(in, (int16_t)ff, (int16_t)mf, out);
  r[0] = r[16] = r[32] = r[48] = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) dct[4 * t + i] = out[i];
  if (nz) nz[t] = z;
}
__global__ void k_hadamard_t4_dc(const int16_t* __restrict__ dct, int n, int16_t* dc) {
  const int t = tjob();
  if (t >= n) return;
  int16_t in[16], out[16];
#pragma unroll
  for (int k = 0; k < 16; k++) in[k] = dct[256 * t + 16 * k];
  hadamard_t4_dc(out, in);
  st16(dc + 16 * t, out);
}
__global__ void k_scan4x4(const int16_t* __restrict__ dct, int n, int16_t* dcac, int16_t* ac, int32_t* ctr_nzc) {
  const int t = tjob();
  if (t >= n) return;
  int16_t d[16], l[16];
  ld16(d, dct + 16 * t);
  scan4x4_dcac(l, d);
  if (dcac) st16(dcac + 16 * t, l);
  if (ctr_nzc) { ctr_nzc[2 * t] = single_ctr4x4(l); ctr_nzc[2 * t + 1] = nonzero_count(l); }
  if (ac) { scan4x4_ac(l, d); st16(ac + 16 * t, l); }
}
__global__ void k_dequant_four4x4(int16_t* res, int qp, int n4) {
  const int t = tjob();
  if (t >= n4) return;
  int16_t d[16];
  ld16(d, res + 16 * t);
  dequant4x4(d, c_dequant[qp]);
  st16(res + 16 * t, d);
}
__global__ void k_dequant_ihadamard4x4(int16_t* res, int mf, int n) {
  const int t = tjob();
  if (t >= n) return;
  int16_t d[16];
  ld16(d, res + 16 * t);
  dequant_ihadamard4x4(d, (uint16_t)mf);
  st16(res + 16 * t, d);
}
__global__ void k_dequant_luma_dc(int16_t* res, int qp, int n) {
  const int t = tjob();
  if (t >= n) return;
  int16_t d[16];
  ld16(d, res + 16 * t);
  // WelsIHadamard4x4Dc walks rows/cols from the last to the first; the butterflies are independent
  // per row/column so the order is not observable
  ihadamard4x4(d);
  dequant_luma_dc4x4(d, qp);
  st16(res + 16 * t, d);
}
__global__ void k_dequant_ihadamard2x2(int16_t* res, int mf, int n) {
  const int t = tjob();
  if (t >= n) return;
  int16_t d[4] = {res[4 * t], res[4 * t + 1], res[4 * t + 2], res[4 * t + 3]};
  dequant_ihadamard2x2_dc(d, (uint16_t)mf);
#pragma unroll
  for (int i = 0; i < 4; i++) res[4 * t + i] = d[i];
}
__global__ void k_idct_four4x4_rec(const uint8_t* __restrict__ pred, int ps, const int32_t* __restrict__ off,
                                   const int16_t* __restrict__ dct, int n4, uint8_t* rec) {
  const int t = tjob();
  if (t >= n4) return;
  const int j = t >> 2, k = t & 3, ox = (k & 1) * 4, oy = (k >> 1) * 4;
  int16_t d[16];
  ld16(d, dct + 16 * t);
  idct4x4_rec(rec + 64 * j + oy * 8 + ox, 8, pred + off[j] + oy * ps + ox, ps, d);
}
__global__ void k_idct_rec_i16x16_dc(const uint8_t* __restrict__ pred, int ps, const int32_t* __restrict__ off,
                                     const int16_t* __restrict__ dc, int n, uint8_t* rec) {
  // one thread per pixel row of a job
  const int t = tjob();
  if (t >= 16 * n) return;
  const int j = t >> 4, y = t & 15;
  const uint8_t* p = pred + off[j] + y * ps;
#pragma unroll
  for (int x = 0; x < 16; x++)
    rec[256 * j + 16 * y + x] = (uint8_t)clip255(p[x] + ((dc[16 * j + (y & 12) + (x >> 2)] + 32) >> 6));
}
__global__ void k_idct_res_add_pred4(uint8_t* pic, int stride, const int32_t* __restrict__ off,
                                     const int16_t* __restrict__ rs, int n) {
  const int t = tjob();
  if (t >= n) return;
  int16_t d[16];
  ld16(d, rs + 16 * t);
  idct_res_add_pred(pic + off[t], stride, d);
}
// 8x8: one warp-quarter (8 threads) per job: row pass by thread r on row r, transpose through shared
// memory, column pass by thread c on column c, then add to the picture.
__global__ void k_idct_res_add_pred8(uint8_t* pic, int stride, const int32_t* __restrict__ off,
                                     const int16_t* __restrict__ rs, int n) {
  __shared__ int16_t tmp[TPB / 8][64];
  const int t = tjob();
  const int j = t >> 3, r = t & 7, slot = threadIdx.x >> 3;
  const bool live = j < n;
  int16_t in[8], out[8];
  if (live) {
#pragma unroll
    for (int x = 0; x < 8; x++) in[x] = rs[64 * j + 8 * r + x];
    idct8_1d(in, out);
#pragma unroll
    for (int x = 0; x < 8; x++) tmp[slot][8 * r + x] = out[x];
  }
  __syncwarp();
  if (live) {
#pragma unroll
    for (int y = 0; y < 8; y++) in[y] = tmp[slot][8 * y + r];
    idct8_1d(in, out);
    uint8_t* p = pic + off[j] + r;
#pragma unroll
    for (int y = 0; y < 8; y++) p[y * stride] = (uint8_t)clip255(p[y * stride] + ((32 + out[y]) >> 6));
  }
}

extern "C" {
#define LAUNCH(kern, count, ...)                                               \
  do {                                                                         \
    if ((count) <= 0) return 0;                                                \
    kern<<<TGRID(count), 0, (cudaStream_t)stream>>>(__VA_ARGS__);              \
    return b2h264_launched();                                                  \
  } while (0)

int b2h264_k_dct_four4x4(const uint8_t* p1, int s1, const int32_t* o1, const uint8_t* p2, int s2, const int32_t* o2, int n,
                         int16_t* dct, void* stream) {
  LAUNCH(k_dct_four4x4, 4 * n, p1, s1, o1, p2, s2, o2, 4 * n, dct);
}
int b2h264_k_quant_four4x4(int16_t* dct, int qp, int intra, int n, int16_t* max4, void* stream) {
  if (qp < 0 || qp > 51) return cudaErrorInvalidValue;
  LAUNCH(k_quant_four4x4, 4 * n, dct, qp, intra, 4 * n, max4);
}
int b2h264_k_quant4x4_dc(int16_t* dct, int ff, int mf, int n, void* stream) { LAUNCH(k_quant4x4_dc, n, dct, ff, mf, n); }
int b2h264_k_hadamard_quant2x2(int16_t* rs, int ff, int mf, int n, int16_t* dct, int32_t* nz, int32_t* skip, void* stream) {
  LAUNCH(k_hadamard_quant2x2, n, rs, ff, mf, n, dct, nz, skip);
}
int b2h264_k_hadamard_t4_dc(const int16_t* dct, int n, int16_t* dc, void* stream) { LAUNCH(k_hadamard_t4_dc, n, dct, n, dc); }
int b2h264_k_scan4x4(const int16_t* dct, int n, int16_t* dcac, int16_t* ac, int32_t* ctr_nzc, void* stream) {
  LAUNCH(k_scan4x4, n, dct, n, dcac, ac, ctr_nzc);
}
int b2h264_k_dequant_four4x4(int16_t* res, int qp, int n, void* stream) {
  if (qp < 0 || qp > 51) return cudaErrorInvalidValue;
  LAUNCH(k_dequant_four4x4, 4 * n, res, qp, 4 * n);
}
int b2h264_k_dequant_ihadamard4x4(int16_t* res, int mf, int n, void* stream) { LAUNCH(k_dequant_ihadamard4x4, n, res, mf, n); }
int b2h264_k_dequant_luma_dc(int16_t* res, int qp, int n, void* stream) {
  if (qp < 0 || qp > 11) return cudaErrorInvalidValue;
  LAUNCH(k_dequant_luma_dc, n, res, qp, n);
}
int b2h264_k_dequant_ihadamard2x2(int16_t* res, int mf, int n, void* stream) { LAUNCH(k_dequant_ihadamard2x2, n, res, mf, n); }
int b2h264_k_idct_four4x4_rec(const uint8_t* pred, int ps, const int32_t* off, const int16_t* dct, int n, uint8_t* rec,
                              void* stream) {
  LAUNCH(k_idct_four4x4_rec, 4 * n, pred, ps, off, dct, 4 * n, rec);
}
int b2h264_k_idct_rec_i16x16_dc(const uint8_t* pred, int ps, const int32_t* off, const int16_t* dc, int n, uint8_t* rec,
                                void* stream) {
  LAUNCH(k_idct_rec_i16x16_dc, 16 * n, pred, ps, off, dc, n, rec);
}
int b2h264_k_idct_res_add_pred(uint8_t* pic, int stride, const int32_t* off, const int16_t* rs, int size, int n,
                               void* stream) {
  if (size == 4) LAUNCH(k_idct_res_add_pred4, n, pic, stride, off, rs, n);
  if (size == 8) LAUNCH(k_idct_res_add_pred8, 8 * n, pic, stride, off, rs, n);
  return cudaErrorInvalidValue;
}
}  // extern "C"
