// Host side of the tcgen05 weight-gradient kernel (see wgrad_umma.cuh, include/fabric_b200.h).
#include "host_common.cuh"
#include "wgrad_umma.cuh"

using namespace fbh;

namespace {

struct WgPlan {
  fb::WgradParams p;
  int qck, grid, smem, wide, v2, hp;
};

// halo-P form (Ca == 64, 16-row tiles): default on, FABRIC_B200_WGRAD_HP=0 is the A/B switch, =16 / =64 restrict it to the
// 13-band stem / to the 64-channel Q chunks
static int wgrad_hp_mode() {
  static const int v = [] {
    const char* e = getenv("FABRIC_B200_WGRAD_HP");
    return e ? atoi(e) : 1;
  }();
  return v;
}

int plan_wgrad_on(const fb_wgrad_desc* d, WgPlan* pl, const DeviceInfo& di);

int plan_wgrad(const fb_wgrad_desc* d, WgPlan* pl) {
  if (!d) return fail(FB_ERR_ARG, "null descriptor");
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  return plan_wgrad_on(d, pl, di);
}

// the launch plan as a pure function of the descriptor and the device limits (no CUDA calls: testable without a GPU)
int plan_wgrad_on(const fb_wgrad_desc* d, WgPlan* pl, const DeviceInfo& di) {
  if (!d) return fail(FB_ERR_ARG, "null descriptor");
  if (d->G < 1 || d->G > 2 || d->B < 1 || d->H < 1 || d->W < 1) return fail(FB_ERR_SHAPE, "bad G/B/H/W");
  if (d->Ca <= 0 || d->Ca % 64) return fail(FB_ERR_SHAPE, "Ca %d must be a multiple of 64", d->Ca);
  if (!(d->Cb == 16 || (d->Cb > 0 && d->Cb % 64 == 0))) return fail(FB_ERR_SHAPE, "Cb %d must be 16 or k*64", d->Cb);
  fb::WgradParams& p = pl->p;
  memset(&p, 0, sizeof(p));
  p.G = d->G, p.B = d->B, p.H = d->H, p.W = d->W, p.Ca = d->Ca, p.Cb = d->Cb;
  if (d->H > 8) p.bh = 16;
  else if (d->H > 4) p.bh = 8;
  else if (d->H > 2) p.bh = 4;
  else p.bh = 2;
  p.bn = 16 / p.bh;
  p.tiles_x = (d->W + 7) / 8;
  p.tiles_y = (d->H + p.bh - 1) / p.bh;
  p.tiles_b = (d->B + p.bn - 1) / p.bn;
  p.tiles_total = p.tiles_x * p.tiles_y * p.tiles_b * d->G;
  pl->qck = d->Cb == 16 ? 16 : 64;
  // second form (filter row through the Q halo tile, 32-channel Q chunks): every layer but the 13-band stem and the
  // tiny maps of the tests; wide = 0 / 1 select the first form explicitly
  // Measured (64 pairs, both forms on every layer, tools/prof_wgrad.py): the second form trades L2 reads for A-operand
  // shared-memory reads (each P slice is read once per filter row) and wins where the first one is L2-bound hardest --
  // 64->128 0.385 -> 0.324 ms, 128->128 0.627 -> 0.520 ms, 128->256 0.303 -> 0.260 ms -- and loses on Ca = 64 (half of
  // M is padding) and on wide Q (Cb >= 256).  wide = 3 picks per shape.
  const bool v2_ok = d->Cb % 64 == 0 && p.bh == 16;
  pl->v2 = (v2_ok && (d->wide == 2 || (d->wide == 3 && d->Ca >= 128 && d->Cb <= 128))) ? 1 : 0;
  const int hpm = wgrad_hp_mode();
  // (measured, 64 pairs: 13->64@256 0.355 -> 0.280 ms, 64->64@256 G2 0.624 -> 0.579, G1 0.316 -> 0.296; 64->64@128 0.089 -> 0.098:
  //  with few tiles per CTA the doubled epilogue shows, so small maps keep the two-item form unless asked: hpm == 2)
  pl->hp = (!pl->v2 && d->Ca == 64 && p.bh == 16 && d->wide != 0 &&
            (hpm == 2 || d->wide == 4 || ((hpm == 1 || hpm == pl->qck) && p.tiles_total >= 16384)))
               ? 1 : 0;
  p.m_tiles = (d->Ca + 127) / 128;
  p.n_chunks = pl->v2 ? d->Cb / 32 : d->Cb / pl->qck;
  p.row_items = (pl->v2 || pl->hp) ? 1 : (d->Ca == 64 ? 2 : 3);
  const int items = p.m_tiles * p.n_chunks * p.row_items;
  int splits = d->splits;
  if (splits <= 0) {
    // The grid is not persistent: items * splits CTAs run in waves of #SMs, and a partly filled last wave costs a whole
    // CTA duration (ncu: 300 CTAs on 148 SMs kept the SMs active 66 % of the time).  Pick the split count that minimises
    // waves / splits (= time for a fixed amount of work), smallest count on ties (less reduction work).
    const int max_splits = p.tiles_total / 4 > 0 ? p.tiles_total / 4 : 1;
    int hi = 4 * di.sms / items + 1;
    if (hi > max_splits) hi = max_splits;
    if (hi < 1) hi = 1;
    double best = 1e30;
    splits = 1;
    for (int sp = 1; sp <= hi; ++sp) {
      const int waves = (items * sp + di.sms - 1) / di.sms;
      const double cost = (double)waves / sp;
      if (cost < best * 0.999) best = cost, splits = sp;
    }
  }
  p.splits = splits;
  const int stage = pl->v2 ? fb::kWg2Stage : fb::wg_stage_bytes(pl->qck, pl->hp != 0);
  int stages = (di.smem_optin - 2048) / stage;
  if (stages > 6) stages = 6;
  if (stages < 2) return fail(FB_ERR_SHAPE, "not enough shared memory");
  p.stages = stages;
  pl->smem = stages * stage + 2048;
  pl->grid = items * splits;
  pl->wide = d->wide != 0;
  p.ws = d->ws;
  return FB_OK;
}

template <int QCK, bool WIDE, bool HP = false>
int launch_wgrad(const WgPlan& pl, const CUtensorMap& tP, const CUtensorMap& tQ, cudaStream_t st) {
  auto k = fb::wgrad_umma_kernel<QCK, WIDE, HP>;
  FB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem));
  k<<<pl.grid, fb::kWgThreads, pl.smem, st>>>(tP, tQ, pl.p);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

}  // namespace

extern "C" {

int fabric_b200_conv3x3_wgrad_plan(const fb_wgrad_desc* d, int sms, int smem_optin, fb_wgrad_plan* out) {
  if (!d || !out) return fail(FB_ERR_ARG, "null pointer");
  DeviceInfo di;
  di.ok = 1, di.sms = sms, di.smem_optin = smem_optin;
  WgPlan pl;
  int rc = plan_wgrad_on(d, &pl, di);
  if (rc) return rc;
  out->form = pl.v2 ? 2 : (pl.hp ? 3 : 1), out->grid = pl.grid, out->splits = pl.p.splits, out->stages = pl.p.stages;
  out->smem_bytes = pl.smem, out->items = pl.grid / pl.p.splits, out->tiles_total = pl.p.tiles_total;
  return FB_OK;
}

int64_t fabric_b200_conv3x3_wgrad_ws_floats(const fb_wgrad_desc* d) {
  WgPlan pl;
  int rc = plan_wgrad(d, &pl);
  return rc ? rc : (int64_t)pl.p.splits * d->Ca * 9 * d->Cb;
}

int fabric_b200_conv3x3_wgrad_splits(const fb_wgrad_desc* d) {
  WgPlan pl;
  int rc = plan_wgrad(d, &pl);
  return rc ? rc : pl.p.splits;
}

int fabric_b200_conv3x3_wgrad(const fb_wgrad_desc* d, void* stream) {
  WgPlan pl;
  int rc = plan_wgrad(d, &pl);
  if (rc) return rc;
  if (!d->p || !d->q || !d->ws) return fail(FB_ERR_ARG, "null tensor pointer");
  if (!aligned16(d->p) || !aligned16(d->q) || !aligned16(d->ws)) return fail(FB_ERR_ALIGN, "pointers must be 16-byte aligned");
  const fb::WgradParams& p = pl.p;
  CUtensorMap tP, tQ;
  // (halo-P form: one box of 18 rows -- the 16-row tile plus a row above and below)
  rc = make_tmap_act(&tP, d->p, p.Ca, p.W, p.H, p.B, p.G, 64, 8, pl.hp ? 18 : p.bh, p.bn, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  if (pl.v2) rc = make_tmap_act(&tQ, d->q, p.Cb, p.W, p.H, p.B, p.G, 32, 10, 18, 1, CU_TENSOR_MAP_SWIZZLE_64B);
  else rc = make_tmap_act(&tQ, d->q, p.Cb, p.W, p.H, p.B, p.G, pl.qck, 10, p.bh, p.bn,
                          pl.qck == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (pl.v2) {
    auto k = fb::wgrad2_umma_kernel;
    FB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem));
    k<<<pl.grid, fb::kWgThreads, pl.smem, st>>>(tP, tQ, pl.p);
    FB_CUDA(cudaGetLastError());
    return FB_OK;
  }
  if (pl.hp) return pl.qck == 64 ? launch_wgrad<64, true, true>(pl, tP, tQ, st) : launch_wgrad<16, true, true>(pl, tP, tQ, st);
  if (pl.qck == 64) return pl.wide ? launch_wgrad<64, true>(pl, tP, tQ, st) : launch_wgrad<64, false>(pl, tP, tQ, st);
  return pl.wide ? launch_wgrad<16, true>(pl, tP, tQ, st) : launch_wgrad<16, false>(pl, tP, tQ, st);
}

}  // extern "C"
