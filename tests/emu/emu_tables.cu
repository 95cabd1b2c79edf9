// host copies of the quantiser tables for the emulation build (same closed forms as tables.cu)
#include <math.h>
#include <stdint.h>
namespace mbk {
int16_t h_quant_ff[58][8];
int16_t h_quant_mf[52][8];
uint16_t h_dequant[52][8];
uint8_t h_lambda[52];
uint8_t h_chroma_qp[52];
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
}
