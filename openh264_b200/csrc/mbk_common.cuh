// mbk_common.cuh — shared device helpers for the macroblock kernels (sm_100a).
//
// Execution model used throughout: ONE WARP OWNS ONE MACROBLOCK (or one block-level job).
// Every warp-level routine below must be called by all 32 lanes of a converged warp; results
// are returned to all lanes.  Pixel pointers are generic (shared or global).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MBK_FULL 0xffffffffu

namespace mbk {

// quantiser tables live in constant memory, filled once by b2h264 init (closed forms, see tables.cu)
extern __constant__ int16_t c_quant_ff[58][8];   // g_kiQuantInterFF (intra = row qp+6)
extern __constant__ int16_t c_quant_mf[52][8];   // g_kiQuantMF
extern __constant__ uint16_t c_dequant[52][8];   // g_kuiDequantCoeff
extern __constant__ uint8_t c_lambda[52];        // g_kiQpCostTable
extern __constant__ uint8_t c_chroma_qp[52];     // g_kuiChromaQpTable

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int iabs(int v) { return v < 0 ? -v : v; }
__device__ __forceinline__ int clip3(int v, int lo, int hi) { return min(max(v, lo), hi); }
__device__ __forceinline__ int clip255(int v) { return min(max(v, 0), 255); }

// block-size ids follow the reference (encoder/core/inc/wels_const.h:139-148)
enum { BLK_16x16 = 0, BLK_16x8, BLK_8x16, BLK_8x8, BLK_4x4, BLK_8x4, BLK_4x8 };
__device__ __forceinline__ int blk_lw(int blk) { return (0x2323344 >> (blk * 4)) & 0xf; }  // log2(width)
__device__ __forceinline__ int blk_lh(int blk) { return (0x3223434 >> (blk * 4)) & 0xf; }  // log2(height)
__device__ __forceinline__ int blk_w(int blk) { return 1 << blk_lw(blk); }
__device__ __forceinline__ int blk_h(int blk) { return 1 << blk_lh(blk); }

// 4 consecutive bytes at an arbitrary byte address as a little-endian word: two aligned 32-bit
// loads + funnel shift (never reads past the aligned word that holds byte p+3).
__device__ __forceinline__ uint32_t ld4u(const uint8_t* p) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
  const uint32_t sh = (uint32_t)(a & 3) * 8;
  const uint32_t lo = w[0];
  const uint32_t hi = sh ? w[1] : 0u;
  return __funnelshift_r(lo, hi, sh);
}

// bits of the signed Exp-Golomb code of v (encoder/core/inc/svc_enc_golomb.h:84-95)
__device__ __forceinline__ int se_bits(int v) {
  const uint32_t code = v > 0 ? (uint32_t)(2 * v - 1) : (uint32_t)(-2 * v);
  return 2 * (31 - __clz(code + 1)) + 1;
}
// COST_MVD(table, dx, dy) with table[d] = (uint16)(lambda * bits(se(d)))  (md.cpp:797-824)
__device__ __forceinline__ int mvd_cost(int lambda, int dx, int dy) {
  return ((lambda * se_bits(dx)) & 0xffff) + ((lambda * se_bits(dy)) & 0xffff);
}

}  // namespace mbk
