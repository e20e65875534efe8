// C ABI of fabric_b200 (see include/fabric_b200.h): host-side launch logic + the small bandwidth-bound kernels.
#include "host_common.cuh"
#include "conv3x3_umma.cuh"

using namespace fbh;

namespace {

// ------------------------------------------------------------------------------------------------ conv planning
struct ConvPlan {
  fb::Conv3x3Params p;
  int n_tile, ck, halo, grid, smem, ctas, ew, minb;
  int rs;   // register-statistics instantiation (training launches of tiles with <= 2 chunks per epilogue warp)
};

int plan_conv_on(const fb_conv3x3_desc* d, ConvPlan* pl, const DeviceInfo& di);

int plan_conv(const fb_conv3x3_desc* d, ConvPlan* pl) {
  if (!d) return fail(FB_ERR_ARG, "null descriptor");
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  return plan_conv_on(d, pl, di);
}

// the launch plan as a pure function of the descriptor and the device limits (no CUDA calls: testable without a GPU)
int plan_conv_on(const fb_conv3x3_desc* d, ConvPlan* pl, const DeviceInfo& di) {
  if (!d) return fail(FB_ERR_ARG, "null descriptor");
  if (d->G < 1 || d->G > 2 || d->B < 1 || d->H < 1 || d->W < 1) return fail(FB_ERR_SHAPE, "bad G/B/H/W");
  if (!(d->Cin == 16 || (d->Cin > 0 && d->Cin % 64 == 0))) return fail(FB_ERR_SHAPE, "Cin %d must be 16 or k*64", d->Cin);
  if (d->Cout <= 0 || d->Cout % 64) return fail(FB_ERR_SHAPE, "Cout %d must be a multiple of 64", d->Cout);
  const int ck = d->Cin == 16 ? 16 : 64;
  int n_tile = d->tune.n_tile;
  if (n_tile == 0) n_tile = d->Cout == 64 ? 64 : (d->Cout % 256 == 0 ? 256 : 128);
  // product pair without a main output keeps two staging buffers: at 128 wide they leave no room for resident weights
  // (Cin >= 128), and 64-wide tiles with resident weights measured 10 % faster than 128-wide tiles with streamed ones
  if (d->tune.n_tile == 0 && d->prod_out && !d->store_main && n_tile == 128 && d->Cin >= 128) n_tile = 64;
  if (!(n_tile == 64 || n_tile == 128 || n_tile == 256) || d->Cout % n_tile)
    return fail(FB_ERR_SHAPE, "n_tile %d does not divide Cout %d", n_tile, d->Cout);
  if (ck == 16 && n_tile != 64) return fail(FB_ERR_SHAPE, "Cin=16 path is built for n_tile 64 only");
  if (d->head_out && (d->Cout != 64 || n_tile != 64 || !d->head_w || !d->head_b))
    return fail(FB_ERR_SHAPE, "fused head needs Cout == 64 and head weights");
  if (!d->store_main && !d->head_out && !d->prod_out) return fail(FB_ERR_ARG, "nothing to write");
  if (d->prod_out && (d->G != 2 || d->prod_channels < d->Cout || d->prod_channels % 8))
    return fail(FB_ERR_SHAPE, "product fusion needs both date groups and prod_channels >= Cout");
  if (d->prod_out && !d->store_main && n_tile > 128)
    return fail(FB_ERR_SHAPE, "product fusion without the main output needs n_tile <= 128 (date-0 tile kept in smem)");

  fb::Conv3x3Params& p = pl->p;
  memset(&p, 0, sizeof(p));
  p.G = d->G, p.B = d->B, p.H = d->H, p.W = d->W, p.Cin = d->Cin, p.Cout = d->Cout;
  if (d->H > 8) p.bh = 16;
  else if (d->H > 4) p.bh = 8;
  else if (d->H > 2) p.bh = 4;
  else p.bh = 2;
  p.bn = 16 / p.bh;
  p.tiles_x = (d->W + 7) / 8;
  p.tiles_y = (d->H + p.bh - 1) / p.bh;
  p.tiles_b = (d->B + p.bn - 1) / p.bn;
  p.num_m_tiles = p.tiles_x * p.tiles_y * p.tiles_b * d->G;
  p.num_n_tiles = d->Cout / n_tile;
  p.kchunks = d->Cin / ck;
  int halo = d->tune.halo;
  const bool halo_ok = (p.bh == 16);
  if (halo < 0) halo = halo_ok ? 1 : 0;
  if (halo && !halo_ok) return fail(FB_ERR_SHAPE, "halo mode needs H > 8");

  // epilogue warps: 8 (two per TMEM lane quarter) for the 64- and 128-wide tiles, whose epilogue is the bottleneck
  // The 13-band stem (9 MMAs per tile) is bound by the per-tile latency chain of its epilogue, not by any pipe: two CTAs
  // per SM with four epilogue warps each (two independent tile pipelines) measured 0.32 ms against 0.43 ms for one CTA
  // with eight; every other shape lost (their resident weights / stage counts do not fit twice).
  const bool stem_auto = ck == 16 && d->tune.occupancy == 0 && d->tune.epi_warps == 0;
  int ew = d->tune.epi_warps;
  if (ew == 0) ew = stem_auto ? 4 : (n_tile <= 128 ? 8 : 4);
  if (!(ew == 4 || (ew == 8 && n_tile <= 128))) return fail(FB_ERR_ARG, "tune.epi_warps must be 0, 4 or (n_tile <= 128) 8");
  // CTAs per SM: 2 only on request, for 64-wide tiles with four epilogue warps
  const int minb = (d->tune.occupancy == 2 || stem_auto) ? 2 : 1;
  if (d->tune.occupancy < 0 || d->tune.occupancy > 2) return fail(FB_ERR_ARG, "tune.occupancy must be 0, 1 or 2");
  if (minb == 2 && (n_tile != 64 || ew != 4)) return fail(FB_ERR_ARG, "two CTAs per SM need n_tile 64 and 4 epilogue warps");
  const int smem_cap = minb == 2 ? 112 * 1024 : di.smem_optin;
  const int sm_slots = di.sms * minb;
  // CTA pairs (cta_group::2): two adjacent M tiles share one M = 256 MMA and each CTA stages half of the weight rows.
  // Needs an even number of M tiles per date-pair unit and room for at least one pair per N tile.
  const int m_units = p.num_m_tiles / (d->prod_out ? 2 : 1);   // M tiles the persistent loop enumerates
  int ctas = d->tune.ctas;
  const bool pair_ok = (m_units % 2 == 0) && sm_slots >= 2 * p.num_n_tiles;
  if (ctas == 0) ctas = pair_ok ? 2 : 1;
  if (ctas != 1 && ctas != 2) return fail(FB_ERR_ARG, "tune.ctas must be 0 (auto), 1 or 2");
  if (ctas == 2 && !pair_ok) return fail(FB_ERR_SHAPE, "CTA pairs need an even number of M tiles (%d)", m_units);
  // in date-pair mode the persistent loop runs over units (spatial tile x N tile), each visited for both dates;
  // slots = CTAs or CTA pairs
  const long long total = (long long)(m_units / ctas) * p.num_n_tiles;
  p.total_units = (int)total;
  int slots = (d->tune.grid > 0 ? d->tune.grid : sm_slots) / ctas;
  if (slots > total) slots = (int)total;
  else slots = (slots / p.num_n_tiles) * p.num_n_tiles;  // each CTA keeps one N tile for its whole life
  if (slots < 1) slots = (int)(total < p.num_n_tiles ? total : p.num_n_tiles);
  const int grid = slots * ctas;
  const bool n_const = (slots % p.num_n_tiles == 0) || slots == total;
  if (d->stats_ws && slots % p.num_n_tiles != 0 && slots != total)
    return fail(FB_ERR_SHAPE, "stats need a grid that is a multiple of the N tiles");
  if (d->bnbwd_z && (!d->bnbwd_coef || !d->stats_ws)) return fail(FB_ERR_ARG, "bnbwd_z needs bnbwd_coef and stats_ws");
  // per-channel sums (BatchNorm moments / fused BatchNorm-backward reduce) of tiles with at most two chunks per epilogue warp
  // run in the RS instantiation, whose epilogue has nothing else
  const int rs = (d->stats_ws && (n_tile / 32) / (ew / 4) <= 2) ? 1 : 0;
  if (rs && (d->pool_out || d->prod_out || d->head_out || d->scale || d->shift || d->relu || d->shift_in_acc || !d->store_main))
    return fail(FB_ERR_ARG, "per-channel sums on %d-wide tiles come with the raw-accumulator epilogue only", n_tile);
  if (d->shift_in_acc && (d->scale || !d->shift)) return fail(FB_ERR_ARG, "shift_in_acc needs shift and no scale");
  if (d->shift_in_acc && !n_const) return fail(FB_ERR_SHAPE, "shift_in_acc needs one N tile per CTA");

  const int a_bytes = fb::conv_a_stage_bytes(ck, halo);
  const int b_bytes = fb::conv_b_stage_bytes(n_tile, ck) / ctas;
  // product fusion keeps the date-0 tile in a second staging buffer when the N tile is small enough
  // (measured: a second buffer does NOT help ordinary tiles -- the previous store has long drained -- and costs the
  //  128->64 layers their resident weights, so it is used for product pairs only)
  const int out_bufs = (d->prod_out && n_tile <= 128) ? 2 : 1;
  // pooled copy / product through smem staging + TMA store instead of scattered per-lane 16-byte stores
  const int pool_tma = (d->pool_out && p.bh == 16) ? 1 : 0;
  const int prod_tma = (d->prod_out && out_bufs == 2 && !d->store_main) ? 1 : 0;
  const int fixed = out_bufs * 128 * n_tile * 2 + (pool_tma ? out_bufs * 4 * (n_tile / 64) * 1024 : 0) +
                    fb::conv_misc_bytes(n_tile, d->stats_ws != nullptr && !rs, d->bnbwd_z != nullptr) +
                    ((rs && d->bnbwd_z) ? fb::conv_zbuf_bytes(n_tile, ew) : 0) + 1024;
  const int avail = smem_cap - fixed;
  const int kblocks = 9 * p.kchunks;
  int b_res = d->tune.b_resident;
  const bool res_fits = n_const && (slots % p.num_n_tiles == 0 || p.num_n_tiles == 1 || slots == total) &&
                        (long long)kblocks * b_bytes + 2 * a_bytes <= avail && kblocks <= 36;
  if (b_res < 0) b_res = (res_fits && slots < total) ? 1 : 0;
  if (ck == 16) b_res = 1;  // the 13-band stem keeps its 18 KB of weights resident and stages all nine taps at once
  if (b_res && !res_fits) return fail(FB_ERR_SHAPE, "resident weights do not fit (%d k-blocks of %d B)", kblocks, b_bytes);
  // a CTA whose tiles alternate N tiles cannot keep weights resident; with grid == total each CTA has one tile
  int a_st = d->tune.a_stages, b_st = d->tune.b_stages;
  if (b_res) {
    b_st = kblocks;
    if (a_st <= 0) a_st = (avail - b_st * b_bytes) / a_bytes;
    if (a_st > 8) a_st = 8;   // weights resident: all remaining smem prefetches input tiles (HBM latency)
  } else if (halo) {
    if (a_st <= 0) a_st = 2;
    if (b_st <= 0) b_st = (avail - a_st * a_bytes) / b_bytes;
    if (b_st > 8) b_st = 8;
  } else {
    int s = avail / (a_bytes + b_bytes);
    if (s > 8) s = 8;
    if (a_st <= 0) a_st = s;
    if (b_st <= 0) b_st = s;
  }
  if (a_st < 1 || b_st < 1 || a_st > 8 || b_st > 36) return fail(FB_ERR_SHAPE, "bad stage counts %d/%d", a_st, b_st);
  const long long smem = (long long)a_st * a_bytes + (long long)b_st * b_bytes + fixed;
  if (smem > smem_cap) return fail(FB_ERR_SHAPE, "shared memory %lld > %d", smem, smem_cap);
  p.a_stages = a_st, p.b_stages = b_st, p.b_resident = b_res;
  p.relu = d->relu, p.store_main = d->store_main, p.acc_init = d->shift_in_acc ? 1 : 0;
  p.scale = d->scale, p.shift = d->shift;
  p.pool_out = reinterpret_cast<__nv_bfloat16*>(d->pool_out);
  p.stats_out = d->stats_ws;
  p.bnb_z = reinterpret_cast<const __nv_bfloat16*>(d->bnbwd_z), p.bnb_coef = d->bnbwd_coef;
  p.head_w = d->head_w, p.head_b = d->head_b, p.head_out = d->head_out;
  p.prod_out = reinterpret_cast<__nv_bfloat16*>(d->prod_out), p.prod_ct = d->prod_channels, p.y0_ptr = d->y;
  p.pair_dates = d->prod_out ? 1 : 0;
  auto magic = [](int dv) { return (unsigned long long)(((1ULL << 40) + dv - 1) / dv); };
  p.mg_nt = magic(p.num_n_tiles), p.mg_tx = magic(p.tiles_x), p.mg_ty = magic(p.tiles_y), p.mg_tb = magic(p.tiles_b);
  if ((double)p.num_m_tiles * p.num_n_tiles * 65536.0 >= 1.0e12) return fail(FB_ERR_SHAPE, "too many tiles");
  p.out_bufs = out_bufs, p.pool_tma = pool_tma, p.prod_tma = prod_tma;
  pl->n_tile = n_tile, pl->ck = ck, pl->halo = halo, pl->grid = grid, pl->smem = (int)smem, pl->ctas = ctas, pl->ew = ew, pl->minb = minb;
  pl->rs = rs;
  return FB_OK;
}

template <int N_TILE, int CK, bool HALO, bool RES, int CTAS, int EW, int MINB = 1, bool RS = false, uint32_t FIX = 0>
int launch_conv(const ConvPlan& pl, const CUtensorMap& tA, const CUtensorMap& tB, const CUtensorMap& tY,
                const CUtensorMap& tP, const CUtensorMap& tQ, cudaStream_t st) {
  auto k = fb::conv3x3_umma_kernel<N_TILE, CK, HALO, RES, CTAS, EW, MINB, RS, FIX>;
  FB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pl.grid), cfg.blockDim = dim3(fb::conv_threads(EW)), cfg.dynamicSmemBytes = pl.smem, cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;   // a CTA pair = one cluster of 2 on one TPC
  at[0].val.clusterDim.x = CTAS, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
  cfg.attrs = at, cfg.numAttrs = CTAS == 2 ? 1 : 0;
  FB_CUDA(cudaLaunchKernelEx(&cfg, k, tA, tB, tY, tP, tQ, pl.p));
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

// ------------------------------------------------------------------------------------------------ small kernels

// NCHW fp32 -> NHWC bf16 (channel padded).  One block = PX pixels of one image row: coalesced per-channel row reads
// into smem, then 16-byte NHWC writes.  PX = 256 for the 13-band input, 32 for wide tensors (standalone block calls).
__global__ void __launch_bounds__(256) pack_nchw_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                        int C, int Cpad, int H, int W, int PX) {
  extern __shared__ float tile[];  // [C][PX + 1]
  const int b = blockIdx.z, y = blockIdx.y, x0 = blockIdx.x * PX;
  const int nx = min(PX, W - x0);
  const int P1 = PX + 1;
  const float* in = src + ((size_t)b * C * H + y) * W + x0;
  for (int i = threadIdx.x; i < C * PX; i += blockDim.x) {
    const int c = i / PX, x = i % PX;
    if (x < nx) tile[c * P1 + x] = __ldg(in + (size_t)c * H * W + x);
  }
  __syncthreads();
  uint4* out = reinterpret_cast<uint4*>(dst + (((size_t)b * H + y) * W + x0) * Cpad);
  const int C8 = Cpad / 8;
  for (int i = threadIdx.x; i < nx * C8; i += blockDim.x) {
    const int x = i / C8, c0 = (i % C8) * 8;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (c0 + j < C) ? tile[(c0 + j) * P1 + x] : 0.f;
    out[i] = fb::pack8(f);
  }
}

// NCHW fp32 -> NHWC bf16 for the 13-band input (C <= 16, Cpad == 16): one thread = one pixel.  A warp reads 128
// contiguous bytes per channel plane and writes 1 KB contiguous (one 32-byte st.global.v8 per lane): ~35 instructions
// per pixel instead of the smem transpose's ~1000 (that kernel was issue-bound at 23 % of HBM bandwidth).
// T = float (z-scored patches as the reference's loader hands them over) or unsigned short (raw Sentinel-2 digital numbers:
// the per-band z-score of utils/dataloaders.py:94-99 is applied here, so the host ships half the bytes).
// aug (nullable, int32 [B][3] = rot90 quarter turns, flip rows, flip columns; square patches of side S): the loader's
// augmentation (utils/dataloaders.py:152-163) as a gather -- output pixel (i, j) reads the source pixel that np.rot90 +
// np.flip would have moved there.
__device__ __forceinline__ uint32_t aug_source(uint32_t pos, int S, const int* a) {
  int i = pos / S, j = pos - i * S;
  const int s1 = S - 1;
  if (a[2]) j = s1 - j;
  if (a[1]) i = s1 - i;
  const int r = a[0] & 3;
  int si = i, sj = j;
  if (r == 1) si = j, sj = s1 - i;
  else if (r == 2) si = s1 - i, sj = s1 - j;
  else if (r == 3) si = s1 - j, sj = i;
  return (uint32_t)(si * S + sj);
}

template <typename T>
__global__ void __launch_bounds__(256) pack_nchw16_kernel(const T* __restrict__ src, uint32_t* __restrict__ dst, int C,
                                                          uint32_t hw, uint32_t total, const float* __restrict__ mean,
                                                          const float* __restrict__ inv_std, const int* __restrict__ aug = nullptr,
                                                          int S = 0) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i >= total) return;
  const uint32_t b = i / hw;
  uint32_t pos = i - b * hw;
  if (aug) pos = aug_source(pos, S, aug + 3 * b);
  const T* s = src + (size_t)b * C * hw + pos;
  float v[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    v[c] = c < C ? (float)__ldg(s + (size_t)c * hw) : 0.f;
    if (mean && c < C) v[c] = (v[c] - __ldg(mean + c)) * __ldg(inv_std + c);
  }
  uint32_t r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = fb::pack_bf16x2(v[2 * j], v[2 * j + 1]);
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + (size_t)i * 8), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// NHWC bf16 -> NCHW fp32.  One block = 32 pixels of one row x all channels (smem transpose).
__global__ void unpack_nhwc_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int C, int H, int W) {
  extern __shared__ float tile[];  // [32][C+1]
  const int b = blockIdx.z, y = blockIdx.y, x0 = blockIdx.x * 32;
  const int nx = min(32, W - x0);
  const __nv_bfloat16* in = src + (((size_t)b * H + y) * W + x0) * C;
  for (int i = threadIdx.x; i < nx * C; i += blockDim.x) tile[(i / C) * (C + 1) + (i % C)] = __bfloat162float(in[i]);
  __syncthreads();
  for (int i = threadIdx.x; i < C * 32; i += blockDim.x) {
    const int c = i / 32, x = i % 32;
    if (x < nx) dst[(((size_t)b * C + c) * H + y) * W + x0 + x] = tile[x * (C + 1) + c];
  }
}

__global__ void pack_weight_kernel(const float* __restrict__ w, const float* __restrict__ scale, __nv_bfloat16* __restrict__ dst,
                                   int Cout, int Cin, int CinPad, int mode) {
  const size_t n = mode == 0 ? (size_t)Cout * 9 * CinPad : (size_t)Cin * 9 * Cout;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v = 0.f;
    if (mode == 0) {
      const int ci = i % CinPad, tap = (i / CinPad) % 9, co = i / ((size_t)CinPad * 9);
      if (ci < Cin) v = w[((size_t)co * Cin + ci) * 9 + tap] * (scale ? scale[co] : 1.f);
    } else {
      const int co = i % Cout, tap = (i / Cout) % 9, ci = i / ((size_t)Cout * 9);
      v = w[((size_t)co * Cin + ci) * 9 + (8 - tap)];
    }
    dst[i] = __float2bfloat16_rn(v);
  }
}

__global__ void bn_fold_eval_kernel(const float* gamma, const float* beta, const float* rm, const float* rv,
                                    const float* bias, float eps, float* scale, float* shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float s = gamma[c] / sqrtf(rv[c] + eps);
  scale[c] = s;
  shift[c] = ((bias ? bias[c] : 0.f) - rm[c]) * s + beta[c];
}

// decoder input: [skip_d1 * skip_d2 | bilinear x2 (align_corners) of low, zero padded]; one thread = 8 channels.
// 32-bit index math (sizes are checked on the host).
__global__ void __launch_bounds__(256, 4)
build_up_input_kernel(const uint4* __restrict__ skip, const uint4* __restrict__ low, uint4* __restrict__ out, int B, int H,
                      int W, int Cs, int h, int w, int Cl, int low_groups) {
  const uint32_t Ct8 = (Cs + Cl) >> 3, Cs8 = Cs >> 3, Cl8 = Cl >> 3;
  const uint32_t npix = (uint32_t)B * H * W;
  // skip == nullptr: the skip half was already written by the encoder conv's fused product epilogue -> only the
  // upsampled channels [Cs, Cs+Cl) are produced here
  const uint32_t per_pix = skip ? Ct8 : Cl8;
  const uint32_t total = npix * per_pix;
  const uint32_t skip_g = npix * Cs8;  // one date group of skip, in uint4
  const uint32_t low_g = (uint32_t)B * h * w * Cl8;
  const int padT = (H - 2 * h) / 2, padL = (W - 2 * w) / 2;
  const float sy = (2 * h > 1) ? (float)(h - 1) / (float)(2 * h - 1) : 0.f;
  const float sx = (2 * w > 1) ? (float)(w - 1) / (float)(2 * w - 1) : 0.f;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t pix = i / per_pix;
    const uint32_t c8 = skip ? i % per_pix : Cs8 + i % per_pix;
    float r[8];
    if (c8 < Cs8) {
      const uint32_t o = pix * Cs8 + c8;
      float a[8], c[8];
      fb::unpack8(__ldg(skip + o), a);
      fb::unpack8(__ldg(skip + o + skip_g), c);
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = fmaxf(a[j] * c[j], 0.f);
    } else {
      const uint32_t cl = c8 - Cs8;
      const uint32_t x = pix % W, t = pix / W, y = t % H, b = t / H;
      const int uy = (int)y - padT, ux = (int)x - padL;
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = 0.f;
      if (uy >= 0 && uy < 2 * h && ux >= 0 && ux < 2 * w) {
        const float fy = sy * uy, fx = sx * ux;
        const int y0 = (int)fy, x0 = (int)fx;
        const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
        const float ly = fy - y0, lx = fx - x0;
        const float wgt[4] = {(1.f - ly) * (1.f - lx), (1.f - ly) * lx, ly * (1.f - lx), ly * lx};
        const int ys[4] = {y0, y0, y1, y1}, xs[4] = {x0, x1, x0, x1};
        uint4 v0[4], v1[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t o = ((b * h + ys[k]) * w + xs[k]) * Cl8 + cl;
          v0[k] = __ldg(low + o);
          if (low_groups == 2) v1[k] = __ldg(low + o + low_g);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float a[8];
          fb::unpack8(v0[k], a);
          if (low_groups == 2) {
            float c[8];
            fb::unpack8(v1[k], c);
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = fmaxf(a[j] * c[j], 0.f);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) r[j] = fmaf(wgt[k], a[j], r[j]);
        }
      }
    }
    out[pix * Ct8 + c8] = fb::pack8(r);
  }
}

// Upsample-only form of the kernel above (skip half already written by the encoder's fused product): grid = (items of
// one output row / 256, H, B), one thread = 8 channels of one output pixel, no integer division, row weights shared by
// the block.  ~100 instructions per 16 output bytes instead of ~240 (the general kernel is issue-bound).
template <int LG>
__global__ void __launch_bounds__(256) upsample_into_kernel(const uint4* __restrict__ low, uint4* __restrict__ out, int H, int W,
                                                            int Cs8, int h, int w, int Cl8, int cl8_shift, int padT, int padL,
                                                            float sy, float sx, uint32_t low_g) {
  const uint32_t item = blockIdx.x * 256u + threadIdx.x;
  const uint32_t x = cl8_shift >= 0 ? item >> cl8_shift : item / (uint32_t)Cl8;
  if (x >= (uint32_t)W) return;
  const uint32_t cl = item - x * Cl8;
  const uint32_t y = blockIdx.y, b = blockIdx.z;
  const int uy = (int)y - padT, ux = (int)x - padL;
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = 0.f;
  if (uy >= 0 && uy < 2 * h && ux >= 0 && ux < 2 * w) {
    const float fy = sy * uy, fx = sx * ux;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float ly = fy - y0, lx = fx - x0;
    const float wgt[4] = {(1.f - ly) * (1.f - lx), (1.f - ly) * lx, ly * (1.f - lx), ly * lx};
    const uint32_t row0 = (b * h + y0) * w, row1 = (b * h + y1) * w;
    const uint32_t o[4] = {(row0 + x0) * Cl8 + cl, (row0 + x1) * Cl8 + cl, (row1 + x0) * Cl8 + cl, (row1 + x1) * Cl8 + cl};
    uint4 v0[4], v1[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v0[k] = __ldg(low + o[k]);
      if (LG == 2) v1[k] = __ldg(low + o[k] + low_g);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float a[8];
      fb::unpack8(v0[k], a);
      if (LG == 2) {
        float c[8];
        fb::unpack8(v1[k], c);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = fmaxf(a[j] * c[j], 0.f);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = fmaf(wgt[k], a[j], r[j]);
    }
  }
  out[((b * H + y) * W + x) * (uint32_t)(Cs8 + Cl8) + Cs8 + cl] = fb::pack8(r);
}

// 1x1 head: one warp = 32 consecutive pixels, each lane one pixel; weights in smem
__global__ void outconv_kernel(const uint4* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                               float* __restrict__ logits, int B, int H, int W, int C) {
  extern __shared__ float sw[];  // [2][C] + [2]
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sw[i] = w[i];
  if (threadIdx.x < 2) sw[2 * C + threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const size_t plane = (size_t)H * W, total = (size_t)B * plane;
  const int C8 = C / 8;
  for (size_t pix = blockIdx.x * (size_t)blockDim.x + threadIdx.x; pix < total; pix += (size_t)gridDim.x * blockDim.x) {
    float a0 = sw[2 * C], a1 = sw[2 * C + 1];
    const uint4* row = x + pix * C8;
    for (int c8 = 0; c8 < C8; ++c8) {
      float f[8];
      fb::unpack8(row[c8], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a0 = fmaf(f[j], sw[c8 * 8 + j], a0);
        a1 = fmaf(f[j], sw[C + c8 * 8 + j], a1);
      }
    }
    const size_t b = pix / plane, o = pix % plane;
    logits[(b * 2 + 0) * plane + o] = a0;
    logits[(b * 2 + 1) * plane + o] = a1;
  }
}


}  // namespace

// ====================================================================================================== exports
extern "C" {

int fabric_b200_version(void) { return 100; }
const char* fabric_b200_last_error(void) { return g_err; }

int fabric_b200_sm_count(void) {
  DeviceInfo di;
  int rc = device_info(&di);
  return rc ? rc : di.sms;
}

int fabric_b200_pack_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int B, int C, int Cpad, int H, int W, void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!src || !dst) return fail(FB_ERR_ARG, "null pointer");
  if (C < 1 || Cpad < C || Cpad % 8 || B < 1 || H < 1 || W < 1 || H > 65535 || B > 65535) return fail(FB_ERR_SHAPE, "bad shape");
  if (!aligned16(dst)) return fail(FB_ERR_ALIGN, "dst must be 16-byte aligned");
  if (C <= 16 && Cpad == 16 && ((uintptr_t)dst & 31) == 0 && (double)B * H * W < 4.0e9) {
    const uint32_t total = (uint32_t)B * H * W;
    pack_nchw16_kernel<float><<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(src, reinterpret_cast<uint32_t*>(dst), C,
                                                                                    (uint32_t)H * W, total, nullptr, nullptr);
    FB_CUDA(cudaGetLastError());
    return FB_OK;
  }
  const int PX = C <= 32 ? 256 : 32;
  const size_t smem = (size_t)C * (PX + 1) * sizeof(float);
  if (smem > (size_t)di.smem_optin) return fail(FB_ERR_SHAPE, "too many channels for the pack kernel");
  if (smem > 48 * 1024) FB_CUDA(cudaFuncSetAttribute(pack_nchw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((W + PX - 1) / PX, H, B);
  pack_nchw_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), C, Cpad, H, W, PX);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int fabric_b200_pack_nchw_u16_to_nhwc_bf16(const uint16_t* src, void* dst, const float* mean, const float* inv_std, int B,
                                           int C, int H, int W, void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!src || !dst || !mean || !inv_std) return fail(FB_ERR_ARG, "null pointer");
  if (C < 1 || C > 16 || B < 1 || H < 1 || W < 1 || (double)B * H * W >= 4.0e9) return fail(FB_ERR_SHAPE, "bad shape");
  if ((uintptr_t)dst & 31) return fail(FB_ERR_ALIGN, "dst must be 32-byte aligned");
  const uint32_t total = (uint32_t)B * H * W;
  pack_nchw16_kernel<unsigned short><<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      src, reinterpret_cast<uint32_t*>(dst), C, (uint32_t)H * W, total, mean, inv_std);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

__global__ void augment_labels_kernel(const long long* __restrict__ src, long long* __restrict__ dst, const int* __restrict__ aug,
                                      int S, uint32_t total) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i >= total) return;
  const uint32_t hw = (uint32_t)S * S, b = i / hw;
  dst[i] = src[(size_t)b * hw + aug_source(i - b * hw, S, aug + 3 * b)];
}

int fabric_b200_pack_nchw_aug(const void* src, int src_dtype, void* dst, const int* aug, const float* mean,
                              const float* inv_std, int B, int C, int S, void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!src || !dst || !aug) return fail(FB_ERR_ARG, "null pointer");
  if ((mean == nullptr) != (inv_std == nullptr)) return fail(FB_ERR_ARG, "mean and inv_std go together");
  if (C < 1 || C > 16 || B < 1 || S < 1 || (double)B * S * S >= 4.0e9) return fail(FB_ERR_SHAPE, "bad shape");
  if ((uintptr_t)dst & 31) return fail(FB_ERR_ALIGN, "dst must be 32-byte aligned");
  const uint32_t total = (uint32_t)B * S * S;
  cudaStream_t st = (cudaStream_t)stream;
  if (src_dtype == 0)
    pack_nchw16_kernel<float><<<(total + 255) / 256, 256, 0, st>>>(reinterpret_cast<const float*>(src),
                                                                  reinterpret_cast<uint32_t*>(dst), C, (uint32_t)S * S, total,
                                                                  mean, inv_std, aug, S);
  else if (src_dtype == 1)
    pack_nchw16_kernel<unsigned short><<<(total + 255) / 256, 256, 0, st>>>(reinterpret_cast<const unsigned short*>(src),
                                                                           reinterpret_cast<uint32_t*>(dst), C, (uint32_t)S * S,
                                                                           total, mean, inv_std, aug, S);
  else
    return fail(FB_ERR_ARG, "src_dtype must be 0 (fp32) or 1 (uint16)");
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int fabric_b200_augment_labels(const int64_t* src, int64_t* dst, const int* aug, int B, int S, void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!src || !dst || !aug || src == dst) return fail(FB_ERR_ARG, "null pointer or in-place call");
  if (B < 1 || S < 1 || (double)B * S * S >= 4.0e9) return fail(FB_ERR_SHAPE, "bad shape");
  const uint32_t total = (uint32_t)B * S * S;
  augment_labels_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const long long*>(src), reinterpret_cast<long long*>(dst), aug, S, total);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int fabric_b200_unpack_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int B, int C, int H, int W, void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!src || !dst) return fail(FB_ERR_ARG, "null pointer");
  if (C < 1 || C > 1024 || B < 1 || H < 1 || W < 1 || H > 65535 || B > 65535) return fail(FB_ERR_SHAPE, "bad shape");
  const size_t smem = 32 * (size_t)(C + 1) * sizeof(float);
  FB_CUDA(cudaFuncSetAttribute(unpack_nhwc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((W + 31) / 32, H, B);
  unpack_nhwc_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(src), dst, C, H, W);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int fabric_b200_pack_conv3x3_weight(const float* w, void* dst, int Cout, int Cin, int CinPad, int mode, void* stream) {
  return fabric_b200_pack_conv3x3_weight_scaled(w, nullptr, dst, Cout, Cin, CinPad, mode, stream);
}

int fabric_b200_pack_conv3x3_weight_scaled(const float* w, const float* scale, void* dst, int Cout, int Cin, int CinPad, int mode,
                                           void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!w || !dst) return fail(FB_ERR_ARG, "null pointer");
  if (mode != 0 && mode != 1) return fail(FB_ERR_ARG, "mode must be 0 or 1");
  if (scale && mode != 0) return fail(FB_ERR_ARG, "per-output-channel scale is for forward weights (mode 0)");
  if (Cout < 1 || Cin < 1 || CinPad < Cin || (mode == 1 && CinPad != Cin)) return fail(FB_ERR_SHAPE, "bad shape");
  const size_t n = mode == 0 ? (size_t)Cout * 9 * CinPad : (size_t)Cin * 9 * Cout;
  pack_weight_kernel<<<ew_grid(n, 256, di.sms), 256, 0, (cudaStream_t)stream>>>(
      w, scale, reinterpret_cast<__nv_bfloat16*>(dst), Cout, Cin, CinPad, mode);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int fabric_b200_conv3x3_plan(const fb_conv3x3_desc* d, int sms, int smem_optin, fb_conv3x3_plan* out) {
  if (!d || !out) return fail(FB_ERR_ARG, "null pointer");
  DeviceInfo di;
  di.ok = 1, di.sms = sms, di.smem_optin = smem_optin;
  ConvPlan pl;
  int rc = plan_conv_on(d, &pl, di);
  if (rc) return rc;
  out->n_tile = pl.n_tile, out->ck = pl.ck, out->halo = pl.halo, out->grid = pl.grid, out->smem_bytes = pl.smem;
  out->ctas = pl.ctas, out->epi_warps = pl.ew, out->a_stages = pl.p.a_stages, out->b_stages = pl.p.b_stages;
  out->b_resident = pl.p.b_resident, out->out_bufs = pl.p.out_bufs, out->total_units = pl.p.total_units;
  out->pool_tma = pl.p.pool_tma, out->prod_tma = pl.p.prod_tma, out->ctas_per_sm = pl.minb;
  out->reg_stats = pl.rs;
  return FB_OK;
}

int fabric_b200_conv3x3_grid(const fb_conv3x3_desc* d) {
  ConvPlan pl;
  int rc = plan_conv(d, &pl);
  return rc ? rc : pl.grid;
}

int64_t fabric_b200_conv3x3_stats_ws_floats(const fb_conv3x3_desc* d) {
  ConvPlan pl;
  int rc = plan_conv(d, &pl);
  return rc ? rc : (int64_t)pl.grid * 2 * pl.n_tile * 2;
}

int fabric_b200_conv3x3(const fb_conv3x3_desc* d, void* stream) {
  ConvPlan pl;
  int rc = plan_conv(d, &pl);
  if (rc) return rc;
  if (!d->x || !d->w || (d->store_main && !d->y)) return fail(FB_ERR_ARG, "null tensor pointer");
  if (!aligned16(d->x) || !aligned16(d->w) || !aligned16(d->y) || !aligned16(d->pool_out))
    return fail(FB_ERR_ALIGN, "tensor pointers must be 16-byte aligned");
  const fb::Conv3x3Params& p = pl.p;
  const CUtensorMapSwizzle sw_a = pl.ck == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUtensorMap tA, tB, tY;
  if (pl.halo) rc = make_tmap_act(&tA, d->x, p.Cin, p.W, p.H, p.B, p.G, pl.ck, fb::kHaloW, fb::kHaloH, 1, sw_a);
  else rc = make_tmap_act(&tA, d->x, p.Cin, p.W, p.H, p.B, p.G, pl.ck, 8, p.bh, p.bn, sw_a);
  if (rc) return rc;
  rc = make_tmap_2d(&tB, d->w, p.Cout, 9LL * p.Cin, pl.n_tile / pl.ctas, pl.ck, sw_a);   // a CTA of a pair loads half the rows
  if (rc) return rc;
  // y may be absent (head-only): the map is still needed as a kernel argument, point it at x's storage
  // store box = one epilogue warp's 32 pixel rows
  const int bhw = p.bh < 4 ? p.bh : 4;
  // store boxes: one per epilogue warp -- 64 channels (128B swizzle) with 4 warps, 32 channels (64B swizzle) with 8
  const int sbc = pl.ew == 8 ? 32 : 64;
  const CUtensorMapSwizzle ssw = pl.ew == 8 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  if (d->store_main) rc = make_tmap_act(&tY, d->y, p.Cout, p.W, p.H, p.B, p.G, sbc, 8, bhw, 4 / bhw, ssw);
  else tY = tA;
  if (rc) return rc;
  CUtensorMap tP = tA, tQ = tA;   // unused maps still have to be valid kernel arguments
  if (p.prod_tma) rc = make_tmap_act(&tP, d->prod_out, p.prod_ct, p.W, p.H, p.B, 1, sbc, 8, bhw, 4 / bhw, ssw);
  if (rc) return rc;
  if (p.pool_tma) rc = make_tmap_act(&tQ, d->pool_out, p.Cout, p.W / 2, p.H / 2, p.B, p.G, sbc, 4, 2, 1, ssw);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool res = p.b_resident != 0;
  // eval-mode instantiations with the epilogue's option set fixed at compile time (64-wide halo tiles, resident weights, CTA
  // pairs, eight epilogue warps: inc.c2 / down1.c2); FABRIC_B200_CONV_FIX=0 is the A/B switch
  {
    static const bool fix_on = [] {
      const char* e = getenv("FABRIC_B200_CONV_FIX");
      return !(e && e[0] == '0');
    }();
    const uint32_t fl = (p.store_main ? fb::F_MAIN : 0u) | (p.prod_out ? fb::F_PROD : 0u) | ((p.stats_out && !p.bnb_z) ? fb::F_STATS : 0u) |
                        (p.bnb_z ? fb::F_BNB : 0u) | (p.pool_out ? fb::F_POOL : 0u) | (p.pool_tma ? fb::F_POOL_TMA : 0u) |
                        (p.prod_tma ? fb::F_PROD_TMA : 0u) | (p.head_out ? fb::F_HEAD : 0u) | (p.out_bufs == 2 ? fb::F_TWO : 0u) |
                        (p.acc_init ? fb::F_INIT : 0u);
    // (the training instantiation's option sets fixed the same way -- 64-wide layers and the stem -- measured no gain: its
    //  epilogue already folds)
    static const bool fix_more = [] {       // (=1: only the lean set -- the A/B switch of the head and stem sets)
      const char* e = getenv("FABRIC_B200_CONV_FIX");
      return !(e && e[0] == '1');
    }();
    if (fix_on && fix_more && !pl.rs && p.relu && p.acc_init && pl.n_tile == 64 && pl.ck == 16 && pl.halo && res && pl.ctas == 2 &&
        pl.ew == 4 && pl.minb == 2 && fl == fb::kFixMain)
      return launch_conv<64, 16, true, true, 2, 4, 2, false, fb::kFixMain>(pl, tA, tB, tY, tP, tQ, st);
    if (fix_on && !pl.rs && p.relu && p.acc_init && pl.n_tile == 64 && pl.ck == 64 && pl.halo && res && pl.ctas == 2 && pl.ew == 8) {
      // measured (eval forward, 64 pairs): inc.c2 0.662 -> 0.570 ms, down1.c2 likewise; up4.c2 + head 0.314 -> 0.275 ms; the 13-band
      // stem (four epilogue warps, two CTAs per SM) 0.382 -> 0.303 ms.  (A fixed {main output} set for the 64-wide
      // decoder convs measured 2-3 % SLOWER than the generic epilogue -- 92 registers, other schedule --, the same set on the
      // 256-wide product tiles (down2.c2) no different: neither is instantiated.)
      if (fl == fb::kFixLean) return launch_conv<64, 64, true, true, 2, 8, 1, false, fb::kFixLean>(pl, tA, tB, tY, tP, tQ, st);
      if (fl == fb::kFixHead && fix_more) return launch_conv<64, 64, true, true, 2, 8, 1, false, fb::kFixHead>(pl, tA, tB, tY, tP, tQ, st);
    }
  }
#define FB_LAUNCH(NT, CK, HL, RS, EW)                                                                   \
  {                                                                                                     \
    if (pl.ctas == 2) return launch_conv<NT, CK, HL, RS, 2, EW>(pl, tA, tB, tY, tP, tQ, st);            \
    return launch_conv<NT, CK, HL, RS, 1, EW>(pl, tA, tB, tY, tP, tQ, st);                              \
  }
  // the training instantiation (register statistics): CTA pairs only -- every training shape pairs up -- else one CTA
#define FB_LAUNCH_RS(NT, CK, HL, RS, EW, MB)                                                            \
  {                                                                                                     \
    if (pl.ctas == 2) return launch_conv<NT, CK, HL, RS, 2, EW, MB, true>(pl, tA, tB, tY, tP, tQ, st);  \
    return launch_conv<NT, CK, HL, RS, 1, EW, MB, true>(pl, tA, tB, tY, tP, tQ, st);                    \
  }
#define FB_DISPATCH(NT, CK, HL, RS)                                                                     \
  if (pl.n_tile == NT && pl.ck == CK && (pl.halo != 0) == HL && res == RS) {                            \
    if (pl.rs && pl.ew == 8) FB_LAUNCH_RS(NT, CK, HL, RS, 8, 1)                                         \
    if (pl.rs && NT == 64 && pl.minb == 2 && HL && RS) FB_LAUNCH_RS(64, CK, true, true, 4, 2)           \
    if (pl.rs && NT == 64) FB_LAUNCH_RS(64, CK, HL, RS, 4, 1)                                           \
    if (pl.ew == 8) FB_LAUNCH(NT, CK, HL, RS, 8)                                                        \
    if (NT == 64 && HL && RS && pl.minb == 2) {                                                         \
      if (pl.ctas == 2) return launch_conv<64, CK, true, true, 2, 4, 2>(pl, tA, tB, tY, tP, tQ, st);    \
      return launch_conv<64, CK, true, true, 1, 4, 2>(pl, tA, tB, tY, tP, tQ, st);                      \
    }                                                                                                   \
    FB_LAUNCH(NT, CK, HL, RS, 4)                                                                        \
  }
#define FB_DISPATCH4(NT, CK, HL, RS) \
  if (pl.n_tile == NT && pl.ck == CK && (pl.halo != 0) == HL && res == RS) FB_LAUNCH(NT, CK, HL, RS, 4)
  FB_DISPATCH(64, 16, false, true)
  FB_DISPATCH(64, 16, true, true)
  FB_DISPATCH(64, 64, false, false)
  FB_DISPATCH(64, 64, false, true)
  FB_DISPATCH(64, 64, true, false)
  FB_DISPATCH(64, 64, true, true)
  FB_DISPATCH(128, 64, false, false)
  FB_DISPATCH(128, 64, false, true)
  FB_DISPATCH(128, 64, true, false)
  FB_DISPATCH(128, 64, true, true)
  FB_DISPATCH4(256, 64, false, false)
  FB_DISPATCH4(256, 64, true, false)
#undef FB_DISPATCH4
#undef FB_LAUNCH
#undef FB_LAUNCH_RS
#undef FB_DISPATCH
  return fail(FB_ERR_SHAPE, "no kernel for n_tile %d ck %d halo %d resident %d", pl.n_tile, pl.ck, pl.halo, (int)res);
}

int fabric_b200_bn_fold_eval(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                             const float* conv_bias, float eps, float* scale, float* shift, int C, void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!gamma || !beta || !running_mean || !running_var || !scale || !shift) return fail(FB_ERR_ARG, "null pointer");
  bn_fold_eval_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gamma, beta, running_mean, running_var, conv_bias,
                                                                         eps, scale, shift, C);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int fabric_b200_build_up_input(const void* skip, const void* low, void* out, int B, int H, int W, int Cs, int h, int w,
                               int Cl, int low_groups, void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!low || !out) return fail(FB_ERR_ARG, "null pointer");
  if (Cs % 8 || Cl % 8 || 2 * h > H || 2 * w > W || (low_groups != 1 && low_groups != 2)) return fail(FB_ERR_SHAPE, "bad shape");
  if ((double)B * H * W * (Cs + Cl) / 8 >= 4.0e9) return fail(FB_ERR_SHAPE, "tensor too large for 32-bit indexing");
  if (!aligned16(skip) || !aligned16(low) || !aligned16(out)) return fail(FB_ERR_ALIGN, "pointers must be 16-byte aligned");
  if (!skip && H <= 65535 && B <= 65535) {
    const int Cl8 = Cl / 8;
    int shift = -1;
    for (int sft = 0; sft < 12; ++sft)
      if ((1 << sft) == Cl8) shift = sft;
    const float sy = (2 * h > 1) ? (float)(h - 1) / (float)(2 * h - 1) : 0.f;
    const float sx = (2 * w > 1) ? (float)(w - 1) / (float)(2 * w - 1) : 0.f;
    dim3 grid(((size_t)W * Cl8 + 255) / 256, H, B);
    const uint32_t low_g = (uint32_t)B * h * w * Cl8;
    if (low_groups == 2)
      upsample_into_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(low), reinterpret_cast<uint4*>(out),
                                                                      H, W, Cs / 8, h, w, Cl8, shift, (H - 2 * h) / 2, (W - 2 * w) / 2, sy, sx, low_g);
    else
      upsample_into_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(low), reinterpret_cast<uint4*>(out),
                                                                      H, W, Cs / 8, h, w, Cl8, shift, (H - 2 * h) / 2, (W - 2 * w) / 2, sy, sx, low_g);
    FB_CUDA(cudaGetLastError());
    return FB_OK;
  }
  const size_t n = (size_t)B * H * W * (skip ? Cs + Cl : Cl) / 8;
  build_up_input_kernel<<<ew_grid(n, 256, di.sms), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(skip), reinterpret_cast<const uint4*>(low), reinterpret_cast<uint4*>(out), B, H, W, Cs,
      h, w, Cl, low_groups);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int fabric_b200_outconv(const void* x, const float* w, const float* b, float* logits, int B, int H, int W, int C, void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!x || !w || !b || !logits) return fail(FB_ERR_ARG, "null pointer");
  if (C % 8 || C > 512) return fail(FB_ERR_SHAPE, "bad C");
  const size_t n = (size_t)B * H * W;
  outconv_kernel<<<ew_grid(n, 128, di.sms), 128, (2 * C + 2) * sizeof(float), (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(x), w, b, logits, B, H, W, C);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

}  // extern "C"
