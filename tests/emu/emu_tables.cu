// host copies of the quantiser tables for the emulation build (same closed forms as tables.cu)
#include <math.h>
#include <stdint.h>

#include "../../openh264_b200/csrc/cavlc_tables.h"
#include "../../openh264_b200/csrc/enc_cavlc_bits.cuh"
#include "../../openh264_b200/csrc/h264_bitstream.h"
namespace mbk {
int16_t h_quant_ff[58][8];
int16_t h_quant_mf[52][8];
uint16_t h_dequant[52][8];
uint8_t h_lambda[52];
uint8_t h_chroma_qp[52];
CavlcLen h_cavlc_len;
// H.264 Tables 8-16 / 8-17 (alpha, beta, tc0 for bS 1..3), indexA/indexB 0..51
uint8_t h_alpha[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 4, 4, 5, 6, 7, 8, 9, 10, 12, 13, 15, 17, 20, 22, 25, 28,
                       32, 36, 40, 45, 50, 56, 63, 71, 80, 90, 101, 113, 127, 144, 162, 182, 203, 226, 255, 255};
uint8_t h_beta[52] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 6, 6, 7, 7, 8, 8,
                      9, 9, 10, 10, 11, 11, 12, 12, 13, 13, 14, 14, 15, 15, 16, 16, 17, 17, 18, 18};
uint8_t h_tc0[52][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0},
                        {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 1},
                        {0, 0, 1}, {0, 0, 1}, {0, 0, 1}, {0, 1, 1}, {0, 1, 1}, {1, 1, 1}, {1, 1, 1}, {1, 1, 1}, {1, 1, 1},
                        {1, 1, 2}, {1, 1, 2}, {1, 1, 2}, {1, 1, 2}, {1, 2, 3}, {1, 2, 3}, {2, 2, 3}, {2, 2, 4}, {2, 3, 4},
                        {2, 3, 4}, {3, 3, 5}, {3, 4, 6}, {3, 4, 6}, {4, 5, 7}, {4, 5, 8}, {4, 6, 9}, {5, 7, 10}, {6, 8, 11},
                        {6, 8, 13}, {7, 10, 14}, {8, 11, 16}, {9, 12, 18}, {10, 13, 20}, {11, 15, 23}, {13, 17, 25}};
}  // namespace mbk
using namespace mbk;
extern "C" void b2h264_build_host_tables() {
  static const int mf_base[6][3] = {{26214, 16132, 10486}, {23832, 14980, 9320}, {20164, 13108, 8388},
                                    {18724, 11650, 7294},  {16384, 10486, 6710}, {14564, 9118, 5786}};
  static const int dq_base[6][3] = {{10, 13, 16}, {11, 14, 18}, {13, 16, 20}, {14, 18, 23}, {16, 20, 25}, {18, 23, 29}};
  static const int pos_class[8] = {0, 1, 0, 1, 1, 2, 1, 2};
  for (int qp = 0; qp < 58; qp++) {
    const int s = qp / 6;
    for (int j = 0; j < 8; j++) {
      const long long base = mf_base[qp % 6][pos_class[j]], num = 65536LL << s, den = 6 * base;
      h_quant_ff[qp][j] = (int16_t)((2 * num + den) / (2 * den));
      if (qp < 52) {
        h_quant_mf[qp][j] = (int16_t)((base + (s ? (1 << (s - 1)) : 0)) >> s);
        h_dequant[qp][j] = (uint16_t)(dq_base[qp % 6][pos_class[j]] << s);
      }
    }
  }
  static const uint8_t hi[22] = {29, 30, 31, 32, 32, 33, 34, 34, 35, 35, 36, 36, 37, 37, 37, 38, 38, 38, 39, 39, 39, 39};
  for (int qp = 0; qp < 52; qp++) {
    const double l = pow(2.0, (qp - 12) / 6.0);
    h_lambda[qp] = (uint8_t)(l < 1.0 ? 1 : (int)floor(l + 0.5));
    h_chroma_qp[qp] = (uint8_t)(qp < 30 ? qp : hi[qp - 30]);
  }
  CavlcLen& T = h_cavlc_len;             // as tables.cu: code lengths of the writer's own tables
  for (int c = 0; c < 5; c++) for (int t = 0; t < 17; t++) for (int o = 0; o < 4; o++) T.coeff_token[c][t][o] = (uint8_t)(kCoeffToken[c][t][o] >> 8);
  for (int t = 0; t < 16; t++) for (int z = 0; z < 16; z++) T.total_zeros[t][z] = (uint8_t)(kTotalZeros[t][z] >> 8);
  for (int t = 0; t < 4; t++) for (int z = 0; z < 4; z++) T.total_zeros_cdc[t][z] = (uint8_t)(kTotalZerosChromaDc[t][z] >> 8);
  for (int z = 0; z < 8; z++) for (int r = 0; r < 15; r++) T.run_before[z][r] = (uint8_t)(kRunBefore[z][r] >> 8);
  for (int i = 0; i < 18; i++) T.nc_class[i] = kNcClass[i];
  for (int i = 0; i < 48; i++) { T.cbp_intra[i] = b2h264::cbp_me_table(true)[i]; T.cbp_inter[i] = b2h264::cbp_me_table(false)[i]; }
}
