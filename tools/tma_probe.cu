// tma_probe.cu — stand-alone probe of the bulk tensor copy the encode kernel uses (rank-3 map over stacked byte planes,
// 48 x 48 x 1 box), in the variants that matter: descriptor as __grid_constant__ parameter addressed directly / through a
// pointer that went through shared memory / in global memory; issued from the kernel body or from a non-inlined function.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/_build/tma_probe tools/tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#define CKD(x) do { CUresult r_ = (x); if (r_ != CUDA_SUCCESS) { printf("driver error %d at %s\n", (int)r_, #x); return 1; } } while (0)
#define CKR(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("runtime error %s at %s\n", cudaGetErrorString(e_), #x); return 1; } } while (0)

struct Ctx { const void* tmap; unsigned long long* bar; };

__device__ __noinline__ void issue(const Ctx* c, uint8_t* win, int x, int y, int z) {
  if ((threadIdx.x & 31) == 0) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(win), bar = (uint32_t)__cvta_generic_to_shared(c->bar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(bar), "r"(48 * 48) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(c->tmap), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
  }
  __syncwarp();
}

__device__ int g_bw = 48, g_bh = 48, g_rank = 3;
__global__ void k_probe(const __grid_constant__ CUtensorMap tm, const void* tm_global, int variant, int x, int y, int z, uint8_t* out, int rank, int bw, int bh) {
  __shared__ __align__(128) uint8_t win[4][64 * 64];
  __shared__ __align__(8) unsigned long long bar[4];
  __shared__ Ctx ctx[4];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar[w])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    ctx[w].tmap = (variant & 1) ? tm_global : (const void*)&tm;
    ctx[w].bar = &bar[w];
  }
  __syncthreads();
  if ((variant & 2) && rank == 3 && bw == 48 && bh == 48) issue(&ctx[w], win[w], x + 16 * w, y, z);               // through a real function, pointer out of shared memory
  else if (lane == 0) {
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(win[w]), b = (uint32_t)__cvta_generic_to_shared(&bar[w]);
    const void* t = (variant & 1) ? tm_global : (const void*)&tm;
    asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(b), "r"(bw * bh) : "memory");
    if (rank == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"(dst), "l"(t), "r"(x + 16 * w), "r"(y), "r"(z), "r"(b) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(dst), "l"(t), "r"(x + 16 * w), "r"(y), "r"(b) : "memory");
  }
  uint32_t ok = 0;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar[w]);
  while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(b) : "memory");
  for (int i = lane; i < bw * bh; i += 32) out[w * 64 * 64 + i] = win[w][i];
}

int main(int argc, char** argv) {
  const int rank = argc > 1 ? atoi(argv[1]) : 3, bw = argc > 2 ? atoi(argv[2]) : 48, bh = argc > 3 ? atoi(argv[3]) : 48;
  const int only_variant = argc > 4 ? atoi(argv[4]) : -1;
  const int x_arg = argc > 5 ? atoi(argv[5]) : 37;
  CKR(cudaSetDevice(0));
  CKR(cudaFree(0));
  const int W = 384, H = 256, N = 6;
  const size_t plane = (size_t)W * H + 256;          // multiple of 16
  uint8_t* h = new uint8_t[plane * N];
  for (size_t i = 0; i < plane * N; i++) h[i] = (uint8_t)((i * 131 + (i >> 9) * 7) & 0xff);
  uint8_t *d, *dout; void* dtm;
  CKR(cudaMalloc(&d, plane * N)); CKR(cudaMemcpy(d, h, plane * N, cudaMemcpyHostToDevice));
  CKR(cudaMalloc(&dout, 4 * 64 * 64)); CKR(cudaMalloc(&dtm, 128));
  for (int l2 = 0; l2 < 2; l2++) {
    CUtensorMap tm;
    const cuuint64_t dims[3] = {W, H, N}, strides[2] = {W, plane};
    const cuuint32_t bx[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
    CKD(cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, strides, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, l2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
    CKR(cudaMemcpy(dtm, &tm, 128, cudaMemcpyHostToDevice));
    for (int variant = 0; variant < 4; variant++) {
      if (only_variant >= 0 && variant != only_variant) continue;
      const int x = x_arg, y = 11, z = rank == 3 ? 4 : 0;
      k_probe<<<1, 128>>>(tm, dtm, variant, x, y, z, dout, rank, bw, bh);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("x=%d rank=%d box=%dx%d l2promo=%d variant=%d (desc %s, issued from %s): FAILED %s\n", x_arg, rank, bw, bh, l2, variant, (variant & 1) ? "global" : "param",
               (variant & 2) ? "function" : "kernel body", cudaGetErrorString(e));
        return 2;                                       // sticky error: the context is gone
      }
      static uint8_t got[4 * 64 * 64];
      CKR(cudaMemcpy(got, dout, sizeof(got), cudaMemcpyDeviceToHost));
      int bad = 0;
      for (int w = 0; w < 4; w++) for (int r = 0; r < bh; r++) for (int c = 0; c < bw; c++)
        if (got[w * 4096 + r * bw + c] != h[z * plane + (size_t)(y + r) * W + x + 16 * w + c]) bad++;
      printf("x=%d rank=%d box=%dx%d l2promo=%d variant=%d (desc %s, issued from %s): %s (%d wrong bytes)\n", x_arg, rank, bw, bh, l2, variant, (variant & 1) ? "global" : "param",
             (variant & 2) ? "function" : "kernel body", bad ? "WRONG DATA" : "ok", bad);
    }
  }
  return 0;
}
