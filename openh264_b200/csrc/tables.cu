// tables.cu — quantiser / cost tables in constant memory + library runtime (init, memory helpers).
// The tables are closed forms of the H.264 quantiser design; tests check them entry-by-entry against
// the reference's literal arrays (g_kiQuantInterFF / g_kiQuantMF encode_mb_aux.cpp:38,103;
// g_kuiDequantCoeff common_tables.cpp:208; g_kiQpCostTable encoder_data_tables.cpp:59).
#include <stdlib.h>
#include <atomic>
#include <math.h>

#include "b2h264_internal.h"
#include "cavlc_tables.h"
#include "enc_cavlc_bits.cuh"
#include "h264_bitstream.h"

namespace mbk {
__constant__ int16_t c_quant_ff[58][8];
__constant__ int16_t c_quant_mf[52][8];
__constant__ uint16_t c_dequant[52][8];
__constant__ uint8_t c_lambda[52];
__constant__ uint8_t c_chroma_qp[52];
}  // namespace mbk

namespace mbk {
__device__ CavlcLen d_cavlc_len;
CavlcLen h_cavlc_len;
}  // namespace mbk

static std::atomic<unsigned long long> g_launches{0};
int b2h264_launched() {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return (int)cudaPeekAtLastError();
}

// host copies (also used by the host-side encoder)
namespace mbk {
int16_t h_quant_ff[58][8];
int16_t h_quant_mf[52][8];
uint16_t h_dequant[52][8];
uint8_t h_lambda[52];
uint8_t h_chroma_qp[52];
}  // namespace mbk
using namespace mbk;

static void build_host_tables() {
  // 2x the standard's multiplication factors for qp%6 at position classes (0,0) / (0,1) / (1,1)
  static const int mf_base[6][3] = {{26214, 16132, 10486}, {23832, 14980, 9320}, {20164, 13108, 8388},
                                    {18724, 11650, 7294},  {16384, 10486, 6710}, {14564, 9118, 5786}};
  static const int dq_base[6][3] = {{10, 13, 16}, {11, 14, 18}, {13, 16, 20}, {14, 18, 23}, {16, 20, 25}, {18, 23, 29}};
  static const int pos_class[8] = {0, 1, 0, 1, 1, 2, 1, 2};
  for (int qp = 0; qp < 58; qp++) {
    const int s = qp / 6;
    for (int j = 0; j < 8; j++) {
      const long long base = mf_base[qp % 6][pos_class[j]];
      const long long num = 65536LL << s, den = 6 * base;
      h_quant_ff[qp][j] = (int16_t)((2 * num + den) / (2 * den));        // round(2^16 / MF / 6)
      if (qp < 52) {
        h_quant_mf[qp][j] = (int16_t)((base + (s ? (1 << (s - 1)) : 0)) >> s);
        h_dequant[qp][j] = (uint16_t)(dq_base[qp % 6][pos_class[j]] << s);
      }
    }
  }
  for (int qp = 0; qp < 52; qp++) {
    const double l = pow(2.0, (qp - 12) / 6.0);                          // lambda = max(1, round(2^((qp-12)/6)))
    h_lambda[qp] = (uint8_t)(l < 1.0 ? 1 : (int)floor(l + 0.5));
    static const uint8_t hi[22] = {29, 30, 31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39};
    h_chroma_qp[qp] = (uint8_t)(qp < 30 ? qp : hi[qp - 30]);            // H.264 Table 8-15
  }
  // code lengths of the CAVLC tables the host writer uses (cavlc_tables.h: entry = bits << 8 | codeword)
  CavlcLen& T = h_cavlc_len;
  for (int c = 0; c < 5; c++) for (int t = 0; t < 17; t++) for (int o = 0; o < 4; o++) T.coeff_token[c][t][o] = (uint8_t)(kCoeffToken[c][t][o] >> 8);
  for (int t = 0; t < 16; t++) for (int z = 0; z < 16; z++) T.total_zeros[t][z] = (uint8_t)(kTotalZeros[t][z] >> 8);
  for (int t = 0; t < 4; t++) for (int z = 0; z < 4; z++) T.total_zeros_cdc[t][z] = (uint8_t)(kTotalZerosChromaDc[t][z] >> 8);
  for (int z = 0; z < 8; z++) for (int r = 0; r < 15; r++) T.run_before[z][r] = (uint8_t)(kRunBefore[z][r] >> 8);
  for (int i = 0; i < 18; i++) T.nc_class[i] = kNcClass[i];
  for (int i = 0; i < 48; i++) { T.cbp_intra[i] = b2h264::cbp_me_table(true)[i]; T.cbp_inter[i] = b2h264::cbp_me_table(false)[i]; }
}

extern "C" {

int b2h264_abi_version(void) { return B2H264_ABI_VERSION; }

int b2h264_init(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess) return (int)e;
  if (count == 0) return (int)cudaErrorNoDevice;
  if ((e = cudaSetDevice(device)) != cudaSuccess) return (int)e;
  if (const char* ss = getenv("B2H264_STACK")) cudaDeviceSetLimit(cudaLimitStackSize, (size_t)atoi(ss));   // debugging knob
  build_host_tables();
  if ((e = cudaMemcpyToSymbol(mbk::c_quant_ff, h_quant_ff, sizeof(h_quant_ff))) != cudaSuccess) return (int)e;
  if ((e = cudaMemcpyToSymbol(mbk::c_quant_mf, h_quant_mf, sizeof(h_quant_mf))) != cudaSuccess) return (int)e;
  if ((e = cudaMemcpyToSymbol(mbk::c_dequant, h_dequant, sizeof(h_dequant))) != cudaSuccess) return (int)e;
  if ((e = cudaMemcpyToSymbol(mbk::c_lambda, h_lambda, sizeof(h_lambda))) != cudaSuccess) return (int)e;
  if ((e = cudaMemcpyToSymbol(mbk::c_chroma_qp, h_chroma_qp, sizeof(h_chroma_qp))) != cudaSuccess) return (int)e;
  if ((e = cudaMemcpyToSymbol(mbk::d_cavlc_len, &h_cavlc_len, sizeof(h_cavlc_len))) != cudaSuccess) return (int)e;
  return (int)cudaDeviceSynchronize();
}

int b2h264_dev_malloc(void** dptr, size_t bytes) { return (int)cudaMalloc(dptr, bytes); }
int b2h264_dev_free(void* dptr) { return (int)cudaFree(dptr); }
int b2h264_h2d(void* dptr, const void* h_src, size_t bytes, void* stream) {
  return (int)cudaMemcpyAsync(dptr, h_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream);
}
int b2h264_d2h(void* h_dst, const void* dptr, size_t bytes, void* stream) {
  return (int)cudaMemcpyAsync(h_dst, dptr, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
}
int b2h264_sync(void* stream) { return (int)cudaStreamSynchronize((cudaStream_t)stream); }
const char* b2h264_error_string(int code) { return cudaGetErrorString((cudaError_t)code); }
unsigned long long b2h264_launch_count(void) { return g_launches.load(); }

/* host-visible copies of the tables, for tests that pin them against the reference's arrays */
const int16_t* b2h264_table_quant_ff(int q) { build_host_tables(); return h_quant_ff[q]; }
const int16_t* b2h264_table_quant_mf(int q) { build_host_tables(); return h_quant_mf[q]; }
const uint16_t* b2h264_table_dequant(int q) { build_host_tables(); return h_dequant[q]; }
int b2h264_table_lambda(int q) { build_host_tables(); return h_lambda[q]; }
int b2h264_table_chroma_qp(int q) { build_host_tables(); return h_chroma_qp[q]; }

}  // extern "C"
