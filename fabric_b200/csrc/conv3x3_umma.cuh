// 3x3 / pad 1 / stride 1 convolution as an im2col-free implicit GEMM on tcgen05 tensor cores (sm_100a).
//
// Replaces nn.Conv2d(in, out, 3, padding=1) (+ the BatchNorm/ReLU/MaxPool/1x1-head that follow it) at
// reference models/unet_parts.py:13,16 (and :14-15,17-18,40,86 through the epilogue options).
//
//   D[m, n] = sum_{tap, c} X[pixel(m) + tap, c] * Wt[n, tap, c]         m: 128 output pixels, n: N_TILE out channels
//
// Data layout: activations NHWC bf16 viewed as a 5-D tensor (C, W, H, B, G) (G = date groups), weights
// [Cout][9][Cin] bf16 (K-major), output NHWC bf16.  One CTA = one 128-pixel x N_TILE output tile, persistent
// over tiles; 6 warps: TMA producer, MMA issuer (one elected thread, tcgen05.mma, fp32 accumulators in TMEM,
// double buffered so the epilogue of tile i overlaps the main loop of tile i+1), 4 epilogue warps.
//
// Two A-operand feeding modes:
//   TAP : one TMA box (CK ch x 8 x bh x bn) per filter tap and 64-channel chunk; zero padding = TMA OOB fill.
//   HALO: one TMA box (64 ch x 10 x 18) per chunk = the tile plus its 1-pixel halo; the nine taps are nine
//         UMMA shared-memory descriptors into that ONE buffer (start address shifted by (r*10+s) pixels,
//         stride-byte-offset 1280 = one halo row), so the input crosses L2->SMEM once instead of nine times.
#pragma once
#include "ptx.cuh"

namespace fb {

struct Conv3x3Params {
  int G, B, H, W;
  int Cin;   // padded input channels (multiple of CK)
  int Cout;  // multiple of N_TILE
  int bh, bn;  // tile = bn images x bh rows x 8 columns, bh * bn == 16
  int tiles_x, tiles_y, tiles_b;
  int num_m_tiles, num_n_tiles;
  int total_units;  // persistent work items: (M tile [pair with CTAS = 2]) x N tile (x both dates in pair_dates mode)
  int kchunks;  // Cin / CK
  int a_stages, b_stages;
  int b_resident;  // weights for this CTA's N tile stay in smem for the whole kernel
  int relu;
  int acc_init;  // the accumulator starts at `shift` (written to TMEM by the epilogue warps) instead of zero: with the
                 // BatchNorm scale folded into the weights the epilogue is convert + store, no per-element affine
  int store_main;
  const float* scale;  // per out channel, nullable (=1)
  const float* shift;  // per out channel, nullable (=0)
  __nv_bfloat16* pool_out;  // nullable: 2x2 max-pooled copy [G,B,H/2,W/2,Cout]
  float* stats_out;         // nullable: per-CTA BN moment partials [grid][2][N_TILE][2] (sum, sum of squares)
  // fused BatchNorm-backward reduce (data-gradient launches): the conv output is dL/da of the BatchNorm+ReLU whose
  // pre-activation z is `bnb_z`; the epilogue applies the ReLU mask (z*scale+shift > 0), stores dy = mask * acc, and writes
  // (sum dy, sum dy * z) partials into stats_out instead of the moments.  bnb_coef: fp32 [4][G][Cout] = scale, shift,
  // mean, invstd of that BatchNorm.
  const __nv_bfloat16* bnb_z;
  const float* bnb_coef;
  const float* head_w;      // nullable: fused 1x1 head [2][64]
  const float* head_b;      // [2]
  float* head_out;          // [G*B, 2, H, W] fp32 NCHW
  // product fusion relu(d2*d1) (bidate_model.py:35-38): tiles are scheduled as (date 0, date 1) pairs on the same CTA;
  // the date-1 epilogue multiplies by the date-0 tile (just written, L2-hot) and writes the product into the first
  // Cout channels of the decoder input [B][H][W][prod_ct]
  __nv_bfloat16* prod_out;
  const void* y0_ptr;  // base of the output tensor (date 0 lives at offset 0)
  int prod_ct;
  int pair_dates;  // 1: iterate tiles as date pairs (needs G == 2)
  int out_bufs;    // 1 or 2 output staging buffers; 2 keeps the date-0 tile in smem for the date-1 product
  int pool_tma;    // pooled copy leaves through per-warp smem staging + TMA store (16-row tiles) instead of per-lane stores
  int prod_tma;    // product leaves through the date-1 staging buffer + TMA store (two buffers, no main output)
  unsigned long long mg_nt, mg_tx, mg_ty, mg_tb;  // fast_div magics for num_n_tiles, tiles_x, tiles_y, tiles_b
};

// warps [0, EW): epilogue; then the TMA producer warp; then the MMA issuer warp LAST: the warp scheduler arbitrates in
// favour of the higher warp id, and the MMA warp's instruction stream is the kernel's critical path.
// EW = 4: one warp per TMEM lane quarter.  EW = 8: two warps per quarter, warp group eg = warp / 4 takes every other
// 32-column chunk -- the epilogue of a 64- or 128-wide tile (~420 instructions per chunk with pool / product / moments,
// one warp per scheduler, nothing to hide latency with) was measured to be the bottleneck of those layers.
__host__ __device__ constexpr int conv_threads(int EW) { return EW * 32 + 64; }
constexpr int kHaloW = 10, kHaloH = 18;
constexpr int kHaloBytes = kHaloW * kHaloH * 128;  // 23040
constexpr int kHaloStage = 23552;                  // rounded up to 1024

// one A stage: the halo tile (HALO: 180 pixels x CK channels, rounded up to 1 KB), all nine 128px x 16ch tap boxes
// (Cin = 16 without halo), or one 128px x 64ch tap box
__host__ __device__ constexpr int conv_a_stage_bytes(int CK, bool halo) {
  return halo ? ((kHaloW * kHaloH * CK * 2 + 1023) / 1024) * 1024 : (CK == 16 ? 9 * 128 * 16 * 2 : 128 * CK * 2);
}
__host__ __device__ constexpr int conv_b_stage_bytes(int N_TILE, int CK) { return N_TILE * CK * 2; }
__host__ __device__ constexpr int conv_stats_bytes(int N_TILE) { return 4 * 2 * N_TILE * 2 * 4; }
__host__ __device__ constexpr int conv_bnb_bytes(int N_TILE) { return 2 * 4 * N_TILE * 4; }   // [2 groups][4][N_TILE] fp32
// (RS + fused BatchNorm-backward reduce) z prefetch slots: one 64-byte row (+16 bytes of padding against bank conflicts) per
// epilogue thread and chunk, filled by cp.async one tile ahead
__host__ __device__ constexpr int conv_zbuf_bytes(int N_TILE, int EW) { return EW * 32 * ((N_TILE / 32) / (EW / 4)) * 80; }
__host__ __device__ constexpr int conv_misc_bytes(int N_TILE, bool stats, bool bnb = false) {
  // scale/shift + head weights (+ head exchange with two epilogue groups), stats slabs (only when BN moments are
  // requested), BatchNorm-backward coefficients, barriers + tmem pointer
  return (2 * N_TILE + 136 + 256) * 4 + (stats ? conv_stats_bytes(N_TILE) : 0) + (bnb ? conv_bnb_bytes(N_TILE) : 0) + 1024;
}

// column sums over the 32 lanes of a warp: returns sum_lanes v[lane_id]  (31 shuffles instead of 160)
__device__ __forceinline__ float warp_colsum32(const float (&v)[32], int lane) {
  float a16[16], a8[8], a4[4], a2[2];
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float keep = b4 ? v[i + 16] : v[i], send = b4 ? v[i] : v[i + 16];
    a16[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float keep = b3 ? a16[i + 8] : a16[i], send = b3 ? a16[i] : a16[i + 8];
    a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float keep = b2 ? a8[i + 4] : a8[i], send = b2 ? a8[i] : a8[i + 4];
    a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float keep = b1 ? a4[i + 2] : a4[i], send = b1 ? a4[i] : a4[i + 2];
    a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  float keep = b0 ? a2[1] : a2[0], send = b0 ? a2[0] : a2[1];
  return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

// one level of the same transposing tree: in[2W] -> out[W]; lane bit W decides which half of the columns this lane keeps
template <int W>
__device__ __forceinline__ void colsum_step(const float (&in)[2 * W], float (&out)[W], int lane) {
  const bool b = lane & W;
#pragma unroll
  for (int i = 0; i < W; ++i) {
    const float keep = b ? in[i + W] : in[i], send = b ? in[i] : in[i + W];
    out[i] = keep + __shfl_xor_sync(0xffffffffu, send, W);
  }
}
// the remaining levels: W values per lane -> the sum of column `lane`
template <int W>
__device__ __forceinline__ float colsum_finish(const float (&v)[W], int lane) {
  if constexpr (W == 1) {
    return v[0];
  } else {
    float o[W / 2];
    colsum_step<W / 2>(v, o, lane);
    return colsum_finish<W / 2>(o, lane);
  }
}
template <int V>
struct IntC {
  static constexpr int value = V;
};

struct TileCoord {
  int n0, x0, y0, b0, g;
};
// epilogue options (one bit each; see the kernel's `flags`)
enum : uint32_t { F_MAIN = 1, F_PROD = 2, F_STATS = 4, F_POOL = 8, F_POOL_TMA = 16, F_PROD_TMA = 32, F_HEAD = 64, F_TWO = 128,
                  F_INIT = 256, F_BNB = 512 };
// FIX instantiations: the option set is a template argument (ReLU on, no affine: BatchNorm folded, shift primed in TMEM), so
// every option test folds at compile time.  The generic eval epilogue spends ~150 of its ~490 warp instructions per
// 32-column chunk on those tests (ncu source view of inc.c2: BRA 46, ISETP 49, PLOP3 16, LDCU 16, BSYNC 10 per chunk) and is
// what bounds the 64-wide layers (tensor pipe 52 %, issue slots 40 %: `profiles/r02_ncu_source.md`).
constexpr uint32_t kFixLean = F_PROD | F_POOL | F_POOL_TMA | F_PROD_TMA | F_TWO | F_INIT;   // encoder c2: pooled copy + date product
constexpr uint32_t kFixHead = F_HEAD | F_INIT;                                              // up4.c2: only the 1x1 head's logits leave
constexpr uint32_t kFixMain = F_MAIN | F_INIT;                                              // the 13-band stem: main output only

// n / d for n * d < 2^40 as one 64-bit multiply: m = ceil(2^40 / d) (host).  Replaces MUFU.RCP division sequences in
// the per-tile paths of all three warp roles.
__device__ __forceinline__ int fast_div(int n, unsigned long long m) {
  return (int)(((unsigned long long)(unsigned)n * m) >> 40);
}
// Persistent tile iteration shared by the three warp roles.  A slot is a CTA (CTAS = 1) or a CTA pair (CTAS = 2); slot s
// owns units s, s + nslots, ...; a unit = (M tile or pair of adjacent M tiles) x N tile.  In pair_dates mode every unit
// is visited twice back to back: date 0, then date 1 (product fusion).  With CTAS = 2 the two CTAs take M tiles
// 2k + rank of the same unit.
template <int CTAS>
__device__ __forceinline__ bool tile_at(const Conv3x3Params& p, int it, int N_TILE, int rank, TileCoord& c) {
  const int nslots = CTAS == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int slot = CTAS == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  int unit, date = 0;
  if (p.pair_dates) {
    unit = slot + (it >> 1) * nslots;
    date = it & 1;
  } else {
    unit = slot + it * nslots;
  }
  if (unit >= p.total_units) return false;
  int m = fast_div(unit, p.mg_nt);
  const int nt = unit - m * p.num_n_tiles;
  if (CTAS == 2) m = 2 * m + rank;
  c.n0 = nt * N_TILE;
  int q = fast_div(m, p.mg_tx);
  const int tx = m - q * p.tiles_x;
  m = q;
  q = fast_div(m, p.mg_ty);
  const int ty = m - q * p.tiles_y;
  m = q;
  q = fast_div(m, p.mg_tb);
  const int tb = m - q * p.tiles_b;
  c.g = p.pair_dates ? date : q;
  c.x0 = tx * 8;
  c.y0 = ty * p.bh;
  c.b0 = tb * p.bn;
  return true;
}

// MINB = 2: two CTAs per SM (64-wide tiles with four epilogue warps only: 192 threads x 168 registers and <= 112 KB of
// shared memory each) -- two independent tile pipelines per SM hide each other's per-tile latency chain
// RS ("register statistics"): the training-step instantiation for tiles with at most two 32-column chunks per epilogue warp.
// Its epilogue is raw-accumulator store (+ BatchNorm moments, or + the fused BatchNorm-backward reduce) only -- pooling,
// product fusion, the 1x1 head, the affine / ReLU and accumulator-priming paths are compiled out -- and the per-channel sums
// accumulate in registers across the CTA's tiles (see REGSTATS below).
template <int N_TILE, int CK, bool HALO, bool RES, int CTAS, int EW, int MINB = 1, bool RS = false, uint32_t FIX = 0>
__global__ void __launch_bounds__(conv_threads(EW), MINB)
conv3x3_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmP,
                    const __grid_constant__ CUtensorMap tmQ, const Conv3x3Params p) {
  // tmP: decoder input [1,B,H,W,Ct] (fused product, same per-warp box as tmY); tmQ: pooled output, box 64 ch x 4 x 2
  // tmY: output map with the PER-WARP store box (64 ch x 8 x min(bh,4) x 4/min(bh,4)): each epilogue warp stores its
  // own 32 pixel rows, so the four warps never synchronise with each other in steady state
  static_assert(CK == 64 || CK == 16, "channel chunk is 64 (128B swizzle) or 16 (32B swizzle)");
  static_assert(N_TILE == 64 || N_TILE == 128 || N_TILE == 256, "N tile");
  static_assert(CTAS == 1 || CTAS == 2, "one CTA or a CTA pair (cta_group::2) per tile row");
  static_assert(EW == 4 || EW == 8, "four or eight epilogue warps");
  constexpr int kEpiThreads = EW * 32, kEpiGroups = EW / 4;
  constexpr int A_BYTES = conv_a_stage_bytes(CK, HALO);
  constexpr int A_TX = HALO ? kHaloW * kHaloH * CK * 2 : 128 * CK * 2;
  constexpr int B_BYTES = conv_b_stage_bytes(N_TILE, CK) / CTAS;   // a CTA of a pair stages half of the weight rows
  constexpr int OUT_BYTES = 128 * N_TILE * 2;
  constexpr uint32_t LAYOUT = (CK == 64) ? kLayoutSw128 : kLayoutSw32;
  constexpr uint32_t ROW_BYTES = CK * 2;
  constexpr uint32_t A_SBO = HALO ? kHaloW * ROW_BYTES : 8 * ROW_BYTES;   // 8-pixel group stride: one (halo) row
  constexpr uint32_t PIX16 = ROW_BYTES >> 4;                              // one pixel in descriptor units
  constexpr uint32_t B_SBO = 8 * ROW_BYTES;
  // TMEM accumulator ring: four buffers where they fit (64-wide: 256 columns, also twice per SM; 128-wide: all 512), two
  // for the 256-wide tile.  Tiles of a product pair cost the epilogue very differently (date 1 carries the product), and
  // with only two buffers the MMA warp stalled on every slow one.
  constexpr int NACC = N_TILE == 256 ? 2 : 4;
  constexpr int ACC_SHIFT = NACC == 4 ? 2 : 1;
  constexpr int TMEM_COLS = NACC * N_TILE;
  constexpr int NCHUNK = N_TILE / 32;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_u32);

  const uint32_t a_off = 0;
  const uint32_t b_off = a_off + p.a_stages * A_BYTES;
  const uint32_t out_off = b_off + p.b_stages * B_BYTES;
  constexpr int POOL_BYTES = 4 * (N_TILE / 64) * 1024;   // per buffer: 4 warps x 8 pooled pixels x N_TILE channels
  const uint32_t pool_off = out_off + p.out_bufs * OUT_BYTES;
  const uint32_t ss_off = pool_off + (p.pool_tma ? p.out_bufs * POOL_BYTES : 0);
  const uint32_t st_off = ss_off + (2 * N_TILE + 136 + 256) * 4;
  const uint32_t bnb_off = st_off + ((p.stats_out && !RS) ? conv_stats_bytes(N_TILE) : 0);   // (RS: the sums are handed over through the output staging buffer)
  const uint32_t zbuf_off = bnb_off + (p.bnb_z ? conv_bnb_bytes(N_TILE) : 0);
  const uint32_t bar_off = zbuf_off + ((RS && p.bnb_z) ? conv_zbuf_bytes(N_TILE, EW) : 0);

  float* ss = reinterpret_cast<float*>(sm + ss_off);       // [0,N) scale, [N,2N) shift, [2N,2N+128) head w, +128.. head b
  float* stats = reinterpret_cast<float*>(sm + st_off);    // [4 warps][2 groups][N_TILE][2]
  float* bnb = reinterpret_cast<float*>(sm + bnb_off);     // [2 groups][4: scale, shift, mean, invstd][N_TILE]
  const uint32_t bars = base + bar_off;
  auto full_a = [&](int s) { return bars + 8u * s; };
  auto empty_a = [&](int s) { return bars + 8u * (p.a_stages + s); };
  auto full_b = [&](int s) { return bars + 8u * (2 * p.a_stages + s); };
  auto empty_b = [&](int s) { return bars + 8u * (2 * p.a_stages + p.b_stages + s); };
  const uint32_t tf_bar = bars + 8u * (2 * p.a_stages + 2 * p.b_stages);  // tmem_full[2], tmem_empty[2]
  auto tmem_full = [&](int s) { return tf_bar + 8u * s; };
  auto tmem_empty = [&](int s) { return tf_bar + 8u * NACC + 8u * s; };
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(sm + bar_off + 1000);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = CTAS == 2 ? (int)cluster_ctarank() : 0;
  // barriers the MMA issuer waits on live in the leader CTA (rank 0); the peer arms / completes them remotely
  const uint32_t lead_delta = CTAS == 2 ? mapa_shared(bars, 0) - bars : 0u;
  auto on_leader = [&](uint32_t bar) { return bar + lead_delta; };

  constexpr int kProducerWarp = kEpiThreads / 32, kMmaWarp = kProducerWarp + 1;
  if (warp == kProducerWarp && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmY);
    if (p.prod_tma) prefetch_tmap(&tmP);
    if (p.pool_tma) prefetch_tmap(&tmQ);
  }
  if (warp == kMmaWarp) {
    if (CTAS == 2) {
      tmem2_alloc(base + bar_off + 1000, TMEM_COLS);
      tmem2_relinquish();
    } else {
      tmem_alloc(base + bar_off + 1000, TMEM_COLS);
      tmem_relinquish();
    }
    if (lane == 0) {
      for (int s = 0; s < p.a_stages; ++s) {
        mbar_init(full_a(s), 1);
        mbar_init(empty_a(s), 1);
      }
      for (int s = 0; s < p.b_stages; ++s) {
        mbar_init(full_b(s), 1);
        mbar_init(empty_b(s), 1);
      }
      for (int s = 0; s < NACC; ++s) {
        mbar_init(tmem_full(s), 1);
        mbar_init(tmem_empty(s), CTAS * (kEpiThreads / 32));
      }
      fence_mbar_init();
    }
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all();   // the peer's barriers must be initialised before anything arrives on them
  // (plus the CTA barrier for CTAS == 2 as well: barrier.cluster already orders the allocator's shared-memory write of the
  //  TMEM address before the reads below; the CTA barrier is redundant, free, and keeps the intra-CTA ordering explicit.
  //  compute-sanitizer racecheck still reports ONE pattern on the pair kernels only -- the peer CTA's half of the collective
  //  tcgen05.alloc.cta_group::2 writing this CTA's result slot "racing" with this CTA's own alloc instruction -- which is
  //  internal to that instruction: profiles/r02_sanitizer.md)
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // ring position = (stage, phase) advanced without integer division (these loops are the issue-rate critical path)
  struct Ring {
    int stage = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance(int n) {
      if (++stage == n) {
        stage = 0;
        phase ^= 1;
      }
    }
  };
  constexpr int TAPS_PER_A = (HALO || CK == 16) ? 9 : 1;  // filter taps served by one A stage

  if (warp == kProducerWarp) {
    // ================================================================ TMA producer (whole warp loops, one lane issues)
    Ring ra, rb;
    bool first_tile = true;
    // CTA pair: the leader arms its own barrier for the bytes of BOTH CTAs; the peer only issues its loads, whose
    // complete_tx lands on the leader's barrier (a transiently negative tx-count is fine: the phase cannot complete
    // before the leader's arrival).  No remote arrive on the load path.
    auto arm = [&](uint32_t bar, uint32_t bytes) {
      if (rank == 0) mbar_arrive_expect_tx(bar, CTAS * bytes);
    };
    auto load_a = [&](uint32_t dst, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
      if (CTAS == 2) tma2_load_5d(dst, &tmA, on_leader(bar), c0, c1, c2, c3, c4);
      else tma_load_5d(dst, &tmA, bar, c0, c1, c2, c3, c4);
    };
    auto load_b = [&](uint32_t dst, uint32_t bar, int c0, int c1) {
      if (CTAS == 2) tma2_load_2d(dst, &tmB, on_leader(bar), c0, c1 + rank * (N_TILE / 2));
      else tma_load_2d(dst, &tmB, bar, c0, c1);
    };
    TileCoord tc;
    for (int it = 0; tile_at<CTAS>(p, it, N_TILE, rank, tc); ++it) {
      for (int c = 0; c < p.kchunks; ++c) {
        for (int tap = 0; tap < 9; ++tap) {
          if (tap % TAPS_PER_A == 0) {
            mbar_wait(empty_a(ra.stage), ra.phase ^ 1);
            if (elect_one()) {
              const uint32_t dst = base + a_off + ra.stage * A_BYTES;
              if (HALO) {
                arm(full_a(ra.stage), A_TX);
                load_a(dst, full_a(ra.stage), c * CK, tc.x0 - 1, tc.y0 - 1, tc.b0, tc.g);
              } else if (CK == 16) {
                arm(full_a(ra.stage), 9 * A_TX);
#pragma unroll
                for (int tt = 0; tt < 9; ++tt)
                  load_a(dst + tt * A_TX, full_a(ra.stage), c * CK, tc.x0 + (tt % 3) - 1, tc.y0 + (tt / 3) - 1, tc.b0, tc.g);
              } else {
                arm(full_a(ra.stage), A_TX);
                load_a(dst, full_a(ra.stage), c * CK, tc.x0 + (tap % 3) - 1, tc.y0 + (tap / 3) - 1, tc.b0, tc.g);
              }
            }
            __syncwarp();
            ra.advance(p.a_stages);
          }
          if constexpr (!RES) {
            mbar_wait(empty_b(rb.stage), rb.phase ^ 1);
            if (elect_one()) {
              arm(full_b(rb.stage), B_BYTES);
              load_b(base + b_off + rb.stage * B_BYTES, full_b(rb.stage), tap * p.Cin + c * CK, tc.n0);
            }
            __syncwarp();
            rb.advance(p.b_stages);
          } else if (first_tile) {  // RES: the CTA's weight slab is loaded once
            if (elect_one()) {
              const int s = c * 9 + tap;
              arm(full_b(s), B_BYTES);
              load_b(base + b_off + s * B_BYTES, full_b(s), tap * p.Cin + c * CK, tc.n0);
            }
            __syncwarp();
          }
        }
      }
      first_tile = false;
    }
  } else if (warp == kMmaWarp) {
    if (rank == 0) {   // a CTA pair is fed by the leader's MMA warp alone
    // ================================================================ MMA issuer (whole warp loops, one lane issues)
    // This warp's instruction stream is the critical path of the kernel (one thread feeds the tensor core), so the
    // loop is kept lean: taps fully unrolled with constant descriptor offsets, descriptors = base + small adds.
    constexpr uint32_t idesc = umma_idesc_bf16(128 * CTAS, N_TILE, 0, 0);
    auto mma = [&](uint32_t d, uint64_t ad, uint64_t bd, uint32_t acc_flag) {
      if (CTAS == 2) umma2_bf16(d, ad, bd, idesc, acc_flag);
      else umma_bf16(d, ad, bd, idesc, acc_flag);
    };
    auto commit = [&](uint32_t bar) {
      if (CTAS == 2) umma2_commit(bar);
      else umma_commit(bar);
    };
    constexpr uint64_t a_hi = umma_desc(0, 16, A_SBO, LAYOUT) & 0xFFFFFFFF00000000ull;
    constexpr uint64_t b_hi = umma_desc(0, 16, B_SBO, LAYOUT) & 0xFFFFFFFF00000000ull;
    constexpr uint32_t lo_fixed = static_cast<uint32_t>(umma_desc(0, 16, 0, 0) & 0xFFFFFFFFu);
    const uint32_t a_lo_base = lo_fixed + ((base + a_off) >> 4);   // smem < 256 KB: the 14-bit field cannot overflow
    const uint32_t b_lo_base = lo_fixed + ((base + b_off) >> 4);
    constexpr uint32_t A_STEP = A_BYTES >> 4, B_STEP = B_BYTES >> 4;
    Ring ra, rb;
    uint32_t tile_it = 0;
    TileCoord tc_unused;
    for (; tile_at<CTAS>(p, (int)tile_it, N_TILE, 0, tc_unused); ++tile_it) {
      const int acc = tile_it & (NACC - 1);
      // (acc_init: even the first use of a buffer waits for the epilogue warps, which preload it with the shift)
      mbar_wait(tmem_empty(acc), ((tile_it >> ACC_SHIFT) & 1) ^ (p.acc_init ? 0u : 1u));
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * N_TILE;
      uint32_t accumulate = p.acc_init ? 1u : 0u;
      for (int c = 0; c < p.kchunks; ++c) {
        const bool last_chunk = (c == p.kchunks - 1);
        if constexpr (TAPS_PER_A == 9) {
          const int sa = ra.stage;
          mbar_wait(full_a(sa), ra.phase);
          ra.advance(p.a_stages);
          const uint32_t a_lo = a_lo_base + sa * A_STEP;
          if constexpr (RES) {
            // weights resident: nothing to wait for per tap -> one straight-line run of 9 * CK/16 MMAs per chunk
            if (tile_it == 0) {
#pragma unroll 1
              for (int tap = 0; tap < 9; ++tap) mbar_wait(full_b(c * 9 + tap), 0);
            }
            tc_fence_after();
            if (elect_one()) {
              // rolled over the filter rows (running descriptor offsets): fully unrolling all 36 MMAs made ptxas hoist
              // every descriptor into uniform registers and spill them (8.5 instructions per MMA)
              uint32_t a_row = a_lo, b_row = b_lo_base + c * 9 * B_STEP;
#pragma unroll 1
              for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int s_ = 0; s_ < 3; ++s_) {
                  const uint32_t a_tap = a_row + (HALO ? s_ * PIX16 : s_ * (A_TX >> 4));
#pragma unroll
                  for (int k = 0; k < CK / 16; ++k) {
                    mma(d_tmem, a_hi | (a_tap + 2 * k), b_hi | (b_row + s_ * B_STEP + 2 * k), accumulate);
                    accumulate = 1;
                  }
                }
                a_row += HALO ? kHaloW * PIX16 : 3 * (A_TX >> 4);
                b_row += 3 * B_STEP;
              }
              commit(empty_a(sa));
              if (last_chunk) commit(tmem_full(acc));
            }
            __syncwarp();
            accumulate = 1;
          } else {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int sb = rb.stage;
              mbar_wait(full_b(sb), rb.phase);
              rb.advance(p.b_stages);
              tc_fence_after();
              if (elect_one()) {
                const uint32_t a_tap = a_lo + (HALO ? ((tap / 3) * kHaloW + (tap % 3)) * PIX16 : tap * (A_TX >> 4));
                const uint32_t b_lo = b_lo_base + sb * B_STEP;
#pragma unroll
                for (int k = 0; k < CK / 16; ++k) {
                  mma(d_tmem, a_hi | (a_tap + 2 * k), b_hi | (b_lo + 2 * k), accumulate);
                  accumulate = 1;
                }
                commit(empty_b(sb));
                if (tap == 8) {
                  commit(empty_a(sa));
                  if (last_chunk) commit(tmem_full(acc));
                }
              }
              __syncwarp();
              accumulate = 1;
            }
          }
        } else {
          // per-tap A boxes (TAP mode, 64-channel chunks): both operands arrive per tap
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            const int sa = ra.stage;
            mbar_wait(full_a(sa), ra.phase);
            ra.advance(p.a_stages);
            int sb;
            if constexpr (RES) {
              sb = c * 9 + tap;
              if (tile_it == 0) mbar_wait(full_b(sb), 0);
            } else {
              sb = rb.stage;
              mbar_wait(full_b(sb), rb.phase);
              rb.advance(p.b_stages);
            }
            tc_fence_after();
            if (elect_one()) {
              const uint32_t a_lo = a_lo_base + sa * A_STEP;
              const uint32_t b_lo = b_lo_base + sb * B_STEP;
#pragma unroll
              for (int k = 0; k < CK / 16; ++k) {
                mma(d_tmem, a_hi | (a_lo + 2 * k), b_hi | (b_lo + 2 * k), accumulate);
                accumulate = 1;
              }
              if constexpr (!RES) commit(empty_b(sb));
              commit(empty_a(sa));
              if (last_chunk && tap == 8) commit(tmem_full(acc));
            }
            __syncwarp();
            accumulate = 1;
          }
        }
      }
    }
    }  // leader
  } else {
    // ================================================================ epilogue (kEpiThreads threads)
    // With kEpiGroups == 2 there are two warps per TMEM lane quarter: group eg takes every other 32-column chunk.
    const int q = warp & 3;             // TMEM lane quarter this warp may touch
    const int eg = warp >> 2;           // chunk group (only with more than four epilogue warps)
    const int m = q * 32 + lane;        // pixel row of the tile
    const int etid = threadIdx.x;       // epilogue warps come first
    float* my_stats = stats + q * (2 * N_TILE * 2);
    if (p.stats_out) {
      if constexpr (RS) {
        // this CTA's slot of the global partial array starts at zero: each date group adds its sums once (flush_acc)
        const int slot_ = CTAS == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
        const int pidx_ = CTAS == 2 ? (slot_ / p.num_n_tiles) * (2 * p.num_n_tiles) + rank * p.num_n_tiles + slot_ % p.num_n_tiles
                                    : slot_;
        float* dst_ = p.stats_out + (size_t)pidx_ * (2 * N_TILE * 2);
        for (int i = etid; i < 2 * N_TILE * 2; i += kEpiThreads) dst_[i] = 0.f;
      } else {
        for (int i = etid; i < 4 * 2 * N_TILE * 2; i += kEpiThreads) stats[i] = 0.f;   // (the slabs exist only then)
      }
    }
    if (p.head_out) {
      for (int i = etid; i < 128; i += kEpiThreads) ss[2 * N_TILE + i] = p.head_w[i];
      if (etid < 2) ss[2 * N_TILE + 128 + etid] = p.head_b[etid];
    }
    if (p.bnb_z) {   // (the planner guarantees one N tile per CTA whenever partials are written)
      TileCoord t0;
      if (tile_at<CTAS>(p, 0, N_TILE, rank, t0)) {
        for (int i = etid; i < 2 * 4 * N_TILE; i += kEpiThreads) {
          const int g_ = i / (4 * N_TILE), k_ = (i / N_TILE) & 3, c_ = i % N_TILE;
          bnb[i] = g_ < p.G ? p.bnb_coef[((size_t)k_ * p.G + g_) * p.Cout + t0.n0 + c_] : 0.f;
        }
      }
    }
    const int px = m & 7;
    const int py = (m >> 3) % p.bh;
    const int pn = (m >> 3) / p.bh;
    // this warp's 4 tile rows (32 pixels) as a TMA store box: rows [wy, wy + min(bh,4)) of images [wn, ...)
    const int wn = (q * 4) / p.bh, wy = (q * 4) % p.bh;
    const bool affine = FIX ? false : (!RS && !p.acc_init && (p.scale != nullptr || p.shift != nullptr));
    const bool relu = RS ? false : (FIX ? true : (p.relu != 0));
    const bool extras = FIX ? (FIX & (F_STATS | F_BNB | F_POOL | F_PROD | F_HEAD)) != 0u
                            : (p.stats_out || p.pool_out || p.prod_out || p.head_out);   // one branch for the common plain tile
    // moments in registers: with at most two 32-column chunks per warp the per-tile transposing shuffle tree (62 shuffles and
    // ~250 selects / adds per chunk: measured +0.14 .. +0.28 ms on the 64- and 128-wide training convolutions) is replaced by
    // plain per-thread accumulation over the CTA's tiles; the tree runs once per date group
    constexpr int CPW = NCHUNK / kEpiGroups;          // chunks per warp
    static_assert(!RS || CPW <= 2, "register statistics need at most two chunks per epilogue warp");
    constexpr bool REGSTATS = RS;
    constexpr int AW = REGSTATS ? 32 / CPW : 1;       // accumulators per chunk and moment
    float acc1[REGSTATS ? CPW : 1][AW], acc2[REGSTATS ? CPW : 1][AW];
#pragma unroll
    for (int a_ = 0; a_ < (REGSTATS ? CPW : 1); ++a_)
#pragma unroll
      for (int j = 0; j < AW; ++j) acc1[a_][j] = acc2[a_][j] = 0.f;
    int acc_g = -1;
    // Hand the register sums of date group g_ over to the CTA's global partial slot.  Every epilogue warp walks the same tile
    // sequence, so all of them arrive here at the same tile index (at most three times per kernel): the four lane quarters are
    // summed through the output staging buffer (free once the warps' TMA stores have read it) -- no dedicated shared memory,
    // which is what lets the 128 -> 128 training convolutions keep their 147 KB weight slab resident like the plain launch.
    auto flush_acc = [&](int g_) {
      if constexpr (REGSTATS) {
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
        bar_sync(1, kEpiThreads);
        float* slab = reinterpret_cast<float*>(sm + out_off);   // [4 quarters][N_TILE][2]
#pragma unroll
        for (int sl = 0; sl < CPW; ++sl) {
          const int cc = eg + sl * kEpiGroups;
          const float c1 = colsum_finish<AW>(acc1[sl], lane), c2 = colsum_finish<AW>(acc2[sl], lane);
          *reinterpret_cast<float2*>(slab + ((q * N_TILE) + cc * 32 + lane) * 2) = make_float2(c1, c2);
#pragma unroll
          for (int j = 0; j < AW; ++j) acc1[sl][j] = acc2[sl][j] = 0.f;
        }
        bar_sync(1, kEpiThreads);
        const int slot_ = CTAS == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
        const int pidx_ = CTAS == 2 ? (slot_ / p.num_n_tiles) * (2 * p.num_n_tiles) + rank * p.num_n_tiles + slot_ % p.num_n_tiles
                                    : slot_;
        float* dst_ = p.stats_out + (size_t)pidx_ * (2 * N_TILE * 2) + (size_t)g_ * N_TILE * 2;
        for (int i = etid; i < N_TILE * 2; i += kEpiThreads) {
          float s_ = 0.f;
#pragma unroll
          for (int w_ = 0; w_ < 4; ++w_) s_ += slab[w_ * (N_TILE * 2) + i];
          dst_[i] += s_;
        }
        bar_sync(1, kEpiThreads);   // the slab is read; the staging buffer may be written again
      }
    };
    // epilogue options as ONE opaque register: the chunk loop tests bits instead of re-reading kernel parameters through
    // the constant bank (LDCU + ISETP + BRA per test, each a latency bubble with one or two warps per scheduler)
    uint32_t flags = (p.store_main ? F_MAIN : 0u) | (p.prod_out ? F_PROD : 0u) | ((p.stats_out && !p.bnb_z) ? F_STATS : 0u) |
                     (p.bnb_z ? F_BNB : 0u) |
                     (p.pool_out ? F_POOL : 0u) | (p.pool_tma ? F_POOL_TMA : 0u) | (p.prod_tma ? F_PROD_TMA : 0u) |
                     (p.head_out ? F_HEAD : 0u) | (p.out_bufs == 2 ? F_TWO : 0u) | (p.acc_init ? F_INIT : 0u);
    if constexpr (RS) flags &= (F_MAIN | F_STATS | F_BNB);   // the only epilogue options of the training instantiation
    if constexpr (FIX != 0u) flags = FIX;                    // (the host launches this instantiation for exactly this set)
    else asm volatile("mov.b32 %0, %0;" : "+r"(flags));
    // flag test; in the RS instantiation every other option folds to false at compile time (its code and registers vanish)
    // (per-channel sums of narrow tiles live in the RS instantiation only; the 13-band stem has no data gradient)
    constexpr uint32_t kAllowed = (RS ? (F_MAIN | F_STATS | F_BNB) : (CPW <= 2 ? ~(F_STATS | F_BNB) : ~0u)) &
                                  (CK == 16 ? ~static_cast<uint32_t>(F_BNB) : ~0u);
    auto has = [&](const uint32_t f) -> bool {
      if constexpr (FIX != 0u) return (FIX & f) != 0u;
      if ((f & kAllowed) == 0u) return false;
      return (flags & f & kAllowed) != 0u;
    };
    const int pH = p.H, pW = p.W, pB = p.B, pCout = p.Cout, pCt = p.prod_ct;
    int cur_n0 = -1;
    uint32_t tile_it = 0;
    // Output staging (what the TMA stores read).  EW = 4: per 64-channel group a [128 px][128 B] slab, 128B swizzle, one
    // store box of 64 ch x 32 px per warp.  EW = 8: per 32-channel chunk a [128 px][64 B] slab, 64B swizzle, one store
    // box of 32 ch x 32 px per warp -- the two warps of a lane quarter never have to meet.
    auto stage_ptr = [&](uint8_t* buf, int cc, int i) -> uint4* {   // 16-byte piece i (0..3) of chunk cc, this thread's row
      if (kEpiGroups == 1) return reinterpret_cast<uint4*>(buf + (cc >> 1) * 16384 + m * 128 + ((((cc & 1) * 4 + i) ^ (m & 7)) * 16));
      return reinterpret_cast<uint4*>(buf + cc * 8192 + m * 64 + ((i ^ ((m >> 1) & 3)) * 16));
    };
    float* hx = ss + 2 * N_TILE + 136;   // [128][2] head partial sums handed from warp group 1 to group 0 (EW = 8)
    bar_sync(1, kEpiThreads);   // stats / head constants visible; the ONLY CTA-wide epilogue barrier in steady state
    TileCoord tc;
    // this warp's 32 accumulator columns of chunk cc <- shift[n0 + cc*32 ..] in every lane (pixel row)
    auto preload_shift = [&](int acc_, int cc) {
      uint32_t sv[32];
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 sh = *reinterpret_cast<const float4*>(ss + N_TILE + cc * 32 + j);
        sv[j] = __float_as_uint(sh.x), sv[j + 1] = __float_as_uint(sh.y), sv[j + 2] = __float_as_uint(sh.z),
        sv[j + 3] = __float_as_uint(sh.w);
      }
      tmem_st_32x32(tmem_base + acc_ * N_TILE + cc * 32 + (static_cast<uint32_t>(q * 32) << 16), sv);
    };
    if (has(F_INIT) && tile_at<CTAS>(p, 0, N_TILE, rank, tc)) {
      // (the planner guarantees one N tile per CTA in this mode, so the shift vector is loaded once)
      for (int i = etid; i < N_TILE; i += kEpiThreads) {
        ss[i] = 1.f;
        ss[N_TILE + i] = p.shift[tc.n0 + i];
      }
      cur_n0 = tc.n0;
      bar_sync(1, kEpiThreads);
      for (int cc = eg; cc < NCHUNK; cc += kEpiGroups)
#pragma unroll
        for (int a_ = 0; a_ < NACC; ++a_) preload_shift(a_, cc);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int a_ = 0; a_ < NACC; ++a_) {
          if (CTAS == 2) mbar_arrive_cluster(on_leader(tmem_empty(a_)));
          else mbar_arrive(tmem_empty(a_));
        }
      }
    }
    // (RS, F_BNB) this thread's z row of chunk slot `sl` of tile `it_` -> its smem slot, asynchronously (cp.async): issued one
    // tile ahead, right after the slot's previous contents were consumed, so the global-load latency (the first version read
    // z with plain loads inside the chunk loop: data-gradient launches ran 1.3 - 2x slower) is off the epilogue's critical path
    const uint32_t zslot = base + zbuf_off + static_cast<uint32_t>(((warp * CPW) * 32 + lane) * 80);
    auto prefetch_z = [&](int it_, int sl) {
      if constexpr (RS) {
        TileCoord t_;
        if (!tile_at<CTAS>(p, it_, N_TILE, rank, t_)) return;
        const int gx_ = t_.x0 + px, gy_ = t_.y0 + py, gb_ = t_.b0 + pn;
        const bool ok = gx_ < pW && gy_ < pH && gb_ < pB;
        const int cc_ = eg + sl * kEpiGroups;
        const __nv_bfloat16* src = p.bnb_z + (ok ? ((((size_t)t_.g * pB + gb_) * pH + gy_) * pW + gx_) * pCout + t_.n0 + cc_ * 32 : 0);
#pragma unroll
        for (int i = 0; i < 4; ++i) cp_async16(zslot + sl * (32 * 80) + i * 16, src + i * 8, ok ? 16u : 0u);
        cp_async_commit();
      }
    };
    if constexpr (RS) {
      if (has(F_BNB)) {
#pragma unroll
        for (int sl = 0; sl < CPW; ++sl) prefetch_z(0, sl);
      }
    }
    for (; tile_at<CTAS>(p, (int)tile_it, N_TILE, rank, tc); ++tile_it) {
      const int acc = tile_it & (NACC - 1);
      const int gx = tc.x0 + px, gy = tc.y0 + py, gb = tc.b0 + pn;
      const bool valid = gx < pW && gy < pH && gb < pB;
      if (tc.n0 != cur_n0) {
        // (once per CTA when the grid is a multiple of the N tiles) the warps are not in lockstep: fence both sides
        bar_sync(1, kEpiThreads);
        for (int i = etid; i < N_TILE; i += kEpiThreads) {
          ss[i] = p.scale ? p.scale[tc.n0 + i] : 1.f;
          ss[N_TILE + i] = p.shift ? p.shift[tc.n0 + i] : 0.f;
        }
        cur_n0 = tc.n0;
        bar_sync(1, kEpiThreads);
      }
      // output staging: with two buffers the previous tile stays readable in smem (date-0 tile of a product pair)
      const int ob = has(F_TWO) ? (int)(tile_it & 1) : 0;
      uint8_t* out_sm = sm + out_off + ob * OUT_BYTES;
      const uint8_t* prev_sm = sm + out_off + (ob ^ 1) * OUT_BYTES;
      const bool prod_tile = has(F_PROD) && tc.g == 1;
      if (lane == 0) {   // this warp's own earlier stores
        if (has(F_TWO)) tma_store_wait_read<1>();           // the store issued two tiles ago has drained this buffer
        else if (prod_tile) tma_store_wait_all<0>();             // single buffer: date-0 rows are re-read from L2
        else tma_store_wait_read<0>();
      }
      mbar_wait(tmem_full(acc), (tile_it >> ACC_SHIFT) & 1);
      tc_fence_after();
      __syncwarp();

      float head0 = 0.f, head1 = 0.f;
      if constexpr (REGSTATS) {   // moments accumulate in registers per date group: hand them over when the group changes
        if (has(F_STATS | F_BNB) && tc.g != acc_g) {
          if (acc_g >= 0) flush_acc(acc_g);
          acc_g = tc.g;
        }
      }
      // one 32-column chunk of the tile; `slot_c` = which of this warp's chunks (compile time: indexes the accumulators)
      auto do_chunk = [&](const int cc, auto slot_c) {
        constexpr int SLOT = decltype(slot_c)::value;
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + acc * N_TILE + cc * 32 + (static_cast<uint32_t>(q * 32) << 16), r);
        tmem_ld_wait();
        if (has(F_INIT)) preload_shift(acc, cc);   // the buffer's next tile starts from the shift again
        uint32_t pk[16];
        uint4 zq[4];   // (F_BNB) this pixel's 32 pre-activations z of the BatchNorm being differentiated
        if (has(F_BNB)) {
          if constexpr (RS) {
            cp_async_wait_all();   // this thread's own copies (issued a tile ago)
            const uint4* zs = reinterpret_cast<const uint4*>(sm + zbuf_off + ((warp * CPW + SLOT) * 32 + lane) * 80);
#pragma unroll
            for (int i = 0; i < 4; ++i) zq[i] = zs[i];
          } else if (valid) {
            const uint4* zp = reinterpret_cast<const uint4*>(
                p.bnb_z + ((((size_t)tc.g * pB + gb) * pH + gy) * pW + gx) * pCout + tc.n0 + cc * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) zq[i] = __ldg(zp + i);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) zq[i] = make_uint4(0u, 0u, 0u, 0u);
          }
          // ReLU mask of the forward pass: relu'(z * scale + shift)
          const float* cf = bnb + (tc.g * 4) * N_TILE + cc * 32;
          const uint32_t zw[16] = {zq[0].x, zq[0].y, zq[0].z, zq[0].w, zq[1].x, zq[1].y, zq[1].z, zq[1].w,
                                   zq[2].x, zq[2].y, zq[2].z, zq[2].w, zq[3].x, zq[3].y, zq[3].z, zq[3].w};
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 sc = *reinterpret_cast<const float2*>(cf + 2 * j);
            const float2 sh = *reinterpret_cast<const float2*>(cf + N_TILE + 2 * j);
            if (!(fmaf(bf16_lo(zw[j]), sc.x, sh.x) > 0.f)) r[2 * j] = 0u;
            if (!(fmaf(bf16_hi(zw[j]), sc.y, sh.y) > 0.f)) r[2 * j + 1] = 0u;
          }
          // the slot's contents are in registers (and were used above): refill it for the next tile
          if constexpr (RS) prefetch_z((int)tile_it + 1, SLOT);
        }
        if (affine) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 sc = *reinterpret_cast<const float4*>(ss + cc * 32 + j);
            const float4 sh = *reinterpret_cast<const float4*>(ss + N_TILE + cc * 32 + j);
            v[j + 0] = fmaf(__uint_as_float(r[j + 0]), sc.x, sh.x), v[j + 1] = fmaf(__uint_as_float(r[j + 1]), sc.y, sh.y);
            v[j + 2] = fmaf(__uint_as_float(r[j + 2]), sc.z, sh.z), v[j + 3] = fmaf(__uint_as_float(r[j + 3]), sc.w, sh.w);
          }
          // ReLU rides in the conversion (cvt.rn.relu.bf16x2.f32): relu(round(x)) == round(relu(x))
          if (relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2_relu(v[2 * j], v[2 * j + 1]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
          }
        } else if (relu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2_relu(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
        } else {  // raw accumulator (training forward before BatchNorm, data gradient)
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
        }

        // staging for the TMA store: sub-tile (cc/2) of 64 channels, row m, 16-byte chunk index XOR (m & 7)
        // (without a main output only the date-0 tile of a product pair is staged: its date-1 partner reads it back)
        if (has(F_MAIN) || (has(F_PROD) && tc.g == 0)) {
#pragma unroll
          for (int i = 0; i < 4; ++i) *stage_ptr(out_sm, cc, i) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
        }
        if (extras) {
        if (has(F_STATS | F_BNB)) {
          // F_STATS: moments (sum x, sum x^2) of the values as stored (bf16-rounded); F_BNB: (sum dy, sum dy * xhat) of the
          // masked gradient as stored.  Invalid (out-of-image) pixels contribute 0.
          // value / weight pairs: (x, x) for the moments, (dy, xhat) for the backward sums: s1 = v, s2 = v * wv
          auto load_v = [&](float (&v)[32]) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              v[2 * j] = valid ? bf16_lo(pk[j]) : 0.f;
              v[2 * j + 1] = valid ? bf16_hi(pk[j]) : 0.f;
            }
          };
          auto mul_w = [&](float (&v)[32]) {   // v <- v * (x or xhat)
            if (has(F_BNB)) {
              // raw second sum: sum dy * z; the finalize kernel turns it into sum dy * xhat = invstd * (sum dy*z - mean * sum dy)
              // in double precision (no per-channel mean / invstd loads in this loop)
              const uint32_t zw[16] = {zq[0].x, zq[0].y, zq[0].z, zq[0].w, zq[1].x, zq[1].y, zq[1].z, zq[1].w,
                                       zq[2].x, zq[2].y, zq[2].z, zq[2].w, zq[3].x, zq[3].y, zq[3].z, zq[3].w};
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                v[2 * j] *= bf16_lo(zw[j]);
                v[2 * j + 1] *= bf16_hi(zw[j]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= v[j];
            }
          };
          if constexpr (REGSTATS) {
            // no (CPW = 1) or one (CPW = 2) level of the transposing tree per tile; the rest once per date group.  The two
            // sums are processed one after the other to keep the live register set small.
            float v[32];
            load_v(v);
            if constexpr (AW == 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j) acc1[SLOT][j] += v[j];
            } else {
              float t[16];
              colsum_step<16>(v, t, lane);
#pragma unroll
              for (int j = 0; j < 16; ++j) acc1[SLOT][j] += t[j];
            }
            load_v(v);
            mul_w(v);
            if constexpr (AW == 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j) acc2[SLOT][j] += v[j];
            } else {
              float t[16];
              colsum_step<16>(v, t, lane);
#pragma unroll
              for (int j = 0; j < 16; ++j) acc2[SLOT][j] += t[j];
            }
          } else {
            float s1[32], s2[32];
            load_v(s1);
            load_v(s2);
            mul_w(s2);
            const float cs1 = warp_colsum32(s1, lane);
            const float cs2 = warp_colsum32(s2, lane);
            float* dst = my_stats + (tc.g * N_TILE + cc * 32 + lane) * 2;
            dst[0] += cs1;
            dst[1] += cs2;
          }
        }
        if constexpr (!RS) {
        if (has(F_POOL)) {
          // 2x2 max over (x^1, y^1) neighbours = lanes ^1 and ^8 of the same warp
          uint32_t pm[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            uint32_t a = bf16x2_max(pk[j], __shfl_xor_sync(0xffffffffu, pk[j], 1));
            pm[j] = bf16x2_max(a, __shfl_xor_sync(0xffffffffu, a, 8));
          }
          const int Hp = pH >> 1, Wp = pW >> 1;
          if (has(F_POOL_TMA)) {
            // this warp's 4 tile rows x 8 columns pool to 2 x 4 pixels: row r of a 128B-swizzled [8][64 ch] slab per
            // 64-channel group; the TMA store clips whatever lies outside the pooled image
            if (!(lane & 9)) {   // even column, even row
              const int r = ((lane >> 4) << 2) | ((lane & 7) >> 1);
              uint8_t* pbuf = sm + pool_off + ob * POOL_BYTES + q * ((N_TILE / 64) * 1024);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint8_t* dstp = kEpiGroups == 1 ? pbuf + (cc >> 1) * 1024 + r * 128 + ((((cc & 1) * 4 + i) ^ r) * 16)
                                                : pbuf + cc * 512 + r * 64 + ((i ^ ((r >> 1) & 3)) * 16);
                *reinterpret_cast<uint4*>(dstp) = make_uint4(pm[4 * i], pm[4 * i + 1], pm[4 * i + 2], pm[4 * i + 3]);
              }
            }
          } else if (!(px & 1) && !(py & 1) && (gx >> 1) < Wp && (gy >> 1) < Hp && gb < pB) {
            size_t off = ((((size_t)tc.g * pB + gb) * Hp + (gy >> 1)) * Wp + (gx >> 1)) * pCout + tc.n0 + cc * 32;
            uint4* dst = reinterpret_cast<uint4*>(p.pool_out + off);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = make_uint4(pm[4 * i], pm[4 * i + 1], pm[4 * i + 2], pm[4 * i + 3]);
          }
        }
        if (prod_tile && (valid || has(F_PROD_TMA))) {
          // relu(d2 * d1): this tile is date 1; the same CTA produced the date-0 tile one iteration ago.  With two
          // staging buffers this thread re-reads ITS OWN row of that tile from smem, otherwise from L2.
          const size_t pix = ((size_t)gb * pH + gy) * pW + gx;
          const uint4* a0 = reinterpret_cast<const uint4*>(
              reinterpret_cast<const __nv_bfloat16*>(p.y0_ptr) + pix * pCout + tc.n0 + cc * 32);
          uint4* dst = reinterpret_cast<uint4*>(p.prod_out + pix * pCt + tc.n0 + cc * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint4 o = has(F_TWO) ? *stage_ptr(const_cast<uint8_t*>(prev_sm), cc, i) : __ldcg(a0 + i);
            // packed bf16 multiply: the fp32 product of two bf16 values is exact, so one rounding either way
            const uint32_t ov[4] = {o.x, o.y, o.z, o.w};
            uint32_t rv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) rv[k] = bf16x2_max(bf16x2_mul(pk[4 * i + k], ov[k]), 0u);
            if (has(F_PROD_TMA)) *stage_ptr(out_sm, cc, i) = make_uint4(rv[0], rv[1], rv[2], rv[3]);   // date-1 tile's own buffer
            else dst[i] = make_uint4(rv[0], rv[1], rv[2], rv[3]);
          }
        }
        if (has(F_HEAD)) {
          const float* hw = ss + 2 * N_TILE;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float lo = bf16_lo(pk[j]), hi = bf16_hi(pk[j]);
            head0 = fmaf(lo, hw[cc * 32 + 2 * j], head0);
            head0 = fmaf(hi, hw[cc * 32 + 2 * j + 1], head0);
            head1 = fmaf(lo, hw[64 + cc * 32 + 2 * j], head1);
            head1 = fmaf(hi, hw[64 + cc * 32 + 2 * j + 1], head1);
          }
        }
        }  // !RS
        }  // extras
      };
      if constexpr (REGSTATS) {
        do_chunk(eg, IntC<0>{});
        if constexpr (CPW == 2) do_chunk(eg + kEpiGroups, IntC<1>{});
      } else {
#pragma unroll 1
        for (int cc = eg; cc < NCHUNK; cc += kEpiGroups) do_chunk(cc, IntC<0>{});
      }
      // accumulator drained (and re-primed) -> MMA may overwrite it
      if (has(F_INIT)) tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTAS == 2) mbar_arrive_cluster(on_leader(tmem_empty(acc)));
        else mbar_arrive(tmem_empty(acc));
      }

      if (kEpiGroups == 2 && has(F_HEAD)) {   // each group summed its own 32 channels: group 1 hands its part to group 0
        if (eg == 1) {
          hx[2 * m] = head0;
          hx[2 * m + 1] = head1;
        }
        bar_sync(2 + q, 64);
        if (eg == 0) {
          head0 += hx[2 * m];
          head1 += hx[2 * m + 1];
        }
        bar_sync(2 + q, 64);   // hx is free again before group 1 reaches the next tile
      }
      if (has(F_HEAD) && valid && eg == 0) {
        const float* hb = ss + 2 * N_TILE + 128;
        const size_t img = (size_t)tc.g * pB + gb;
        const size_t plane = (size_t)pH * pW;
        p.head_out[(img * 2 + 0) * plane + (size_t)gy * pW + gx] = head0 + hb[0];
        p.head_out[(img * 2 + 1) * plane + (size_t)gy * pW + gx] = head1 + hb[1];
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        // this warp's boxes: EW = 4 -> 64 channels x 32 px per 64-channel group; EW = 8 -> 32 channels x 32 px per chunk
        constexpr int NBOX = kEpiGroups == 1 ? N_TILE / 64 : NCHUNK / 2;
        const uint32_t obuf = base + out_off + ob * OUT_BYTES, pbuf = base + pool_off + ob * POOL_BYTES + q * ((N_TILE / 64) * 1024);
#pragma unroll
        for (int j = 0; j < NBOX; ++j) {
          const int cc = 2 * j + eg;   // (EW = 8) chunk of box j
          const uint32_t src = kEpiGroups == 1 ? obuf + j * 16384 + q * 4096 : obuf + cc * 8192 + q * 2048;
          const int ch = tc.n0 + (kEpiGroups == 1 ? j * 64 : cc * 32);
          if (has(F_MAIN)) tma_store_5d(&tmY, src, ch, tc.x0, tc.y0 + wy, tc.b0 + wn, tc.g);
          if (has(F_PROD_TMA) && tc.g == 1) tma_store_5d(&tmP, src, ch, tc.x0, tc.y0 + wy, tc.b0 + wn, 0);
          if (has(F_POOL_TMA))
            tma_store_5d(&tmQ, kEpiGroups == 1 ? pbuf + j * 1024 : pbuf + cc * 512, ch, tc.x0 >> 1, (tc.y0 >> 1) + 2 * q, tc.b0,
                         tc.g);
        }
        tma_store_commit();
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
    if constexpr (REGSTATS) {
      if (acc_g >= 0) flush_acc(acc_g);
    }
    if (p.stats_out && !RS) {
      bar_sync(1, kEpiThreads);
      // partial index i with i % num_n_tiles == this CTA's N tile (what bn_finalize assumes): a pair's two CTAs sit
      // num_n_tiles apart
      const int slot = CTAS == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
      const int pidx = CTAS == 2 ? (slot / p.num_n_tiles) * (2 * p.num_n_tiles) + rank * p.num_n_tiles + slot % p.num_n_tiles
                                 : slot;
      float* dst = p.stats_out + (size_t)pidx * (2 * N_TILE * 2);
      for (int i = etid; i < 2 * N_TILE * 2; i += kEpiThreads) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) s += stats[w * (2 * N_TILE * 2) + i];
        dst[i] = s;
      }
    }
  }

  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all();   // the peer's smem / TMEM are in use until the leader's last MMA has retired
  else __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    if (CTAS == 2) tmem2_dealloc(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace fb
