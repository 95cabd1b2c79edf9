/*
 * oracle/ref_shim_dec.cpp — plain-C entry points onto the decoder's weighted / bi-directional prediction (rec_mb.cpp:298-460).
 *
 * TEST INFRASTRUCTURE ONLY.  WeightPrediction, BiWeightPrediction and BiPrediction are file-local (static) in the reference, so this
 * translation unit is compiled WITH the reference's rec_mb.cpp where it lies (#include of the unmodified source, never copied) and
 * forwards to those functions with the reference's own structs; output oracle/_ref/librefshim_dec.so.  Nothing here re-implements
 * any arithmetic.  Each call works on the three planes of a block: Y (w x h), Cb and Cr (w/2 x h/2), as the reference functions do.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "codec/decoder/core/src/rec_mb.cpp"

using namespace WelsDec;

namespace {
struct Ctx {
  SDqLayer* layer;
  SPredWeightTabSyn* tab;
  Ctx() {
    layer = (SDqLayer*) calloc (1, sizeof (SDqLayer));
    tab = (SPredWeightTabSyn*) calloc (1, sizeof (SPredWeightTabSyn));
    layer->pPredWeightTable = tab;
  }
};
Ctx* ctx() { static Ctx* c = new Ctx(); return c; }
void dst_planes (sMCRefMember* m, uint8_t* y, uint8_t* u, uint8_t* v, int sy, int sc) {
  memset (m, 0, sizeof (*m));
  m->pDstY = y; m->pDstU = u; m->pDstV = v;
  m->iDstLineLuma = sy; m->iDstLineChroma = sc;
}
}

extern "C" {

/* explicit weights of ONE list: luma (weight, offset), Cb, Cr */
void ref_weight_pred (uint8_t* y, uint8_t* u, uint8_t* v, int sy, int sc, int w, int h, int log2_denom_luma, int log2_denom_chroma,
                      const int32_t weight[3], const int32_t offset[3]) {
  Ctx* c = ctx();
  c->tab->uiLumaLog2WeightDenom = log2_denom_luma; c->tab->uiChromaLog2WeightDenom = log2_denom_chroma;
  c->tab->sPredList[LIST_0].iLumaWeight[0] = weight[0]; c->tab->sPredList[LIST_0].iLumaOffset[0] = offset[0];
  for (int k = 0; k < 2; k++) { c->tab->sPredList[LIST_0].iChromaWeight[0][k] = weight[1 + k]; c->tab->sPredList[LIST_0].iChromaOffset[0][k] = offset[1 + k]; }
  sMCRefMember m;
  dst_planes (&m, y, u, v, sy, sc);
  WeightPrediction (c->layer, &m, LIST_0, 0, w, h);
}

/* explicit (idc 1): weights / offsets of both lists; implicit (idc 2): w1[0] is iImplicitWeight, log2 denominators are 5 */
void ref_biweight_pred (uint8_t* y, uint8_t* u, uint8_t* v, const uint8_t* ty, const uint8_t* tu, const uint8_t* tv, int sy, int sc, int w, int h,
                        int explicit_weights, int log2_denom_luma, int log2_denom_chroma, const int32_t w1[3], const int32_t o1[3],
                        const int32_t w2[3], const int32_t o2[3]) {
  Ctx* c = ctx();
  c->tab->uiLumaLog2WeightDenom = log2_denom_luma; c->tab->uiChromaLog2WeightDenom = log2_denom_chroma;
  c->tab->sPredList[LIST_0].iLumaWeight[0] = w1[0]; c->tab->sPredList[LIST_0].iLumaOffset[0] = o1[0];
  c->tab->sPredList[LIST_1].iLumaWeight[0] = w2[0]; c->tab->sPredList[LIST_1].iLumaOffset[0] = o2[0];
  for (int k = 0; k < 2; k++) {
    c->tab->sPredList[LIST_0].iChromaWeight[0][k] = w1[1 + k]; c->tab->sPredList[LIST_0].iChromaOffset[0][k] = o1[1 + k];
    c->tab->sPredList[LIST_1].iChromaWeight[0][k] = w2[1 + k]; c->tab->sPredList[LIST_1].iChromaOffset[0][k] = o2[1 + k];
  }
  c->tab->iImplicitWeight[0][0] = w1[0];
  sMCRefMember m, t;
  dst_planes (&m, y, u, v, sy, sc);
  dst_planes (&t, (uint8_t*) ty, (uint8_t*) tu, (uint8_t*) tv, sy, sc);
  BiWeightPrediction (c->layer, &m, &t, 0, 0, explicit_weights != 0, w, h);
}

void ref_bi_pred (uint8_t* y, uint8_t* u, uint8_t* v, const uint8_t* ty, const uint8_t* tu, const uint8_t* tv, int sy, int sc, int w, int h) {
  Ctx* c = ctx();
  sMCRefMember m, t;
  dst_planes (&m, y, u, v, sy, sc);
  dst_planes (&t, (uint8_t*) ty, (uint8_t*) tu, (uint8_t*) tv, sy, sc);
  BiPrediction (c->layer, &m, &t, w, h);
}

} // extern "C"
