// Weight gradient of the 3x3 convolution on tcgen05 tensor cores (sm_100a).
//
//   dW[ca][tap][cb] = sum over pixels  P[pix][ca] * Q[pix + tap][cb]        P = dL/d(conv output), Q = conv input
//
// (autograd of nn.Conv2d, reference train.py:94 -> models/unet_parts.py:13,16).  The reduction runs over PIXELS, so
// both GEMM operands are "MN-major": the NHWC tiles land in shared memory by TMA exactly as they lie in HBM
// ([pixel][64 channels], 128B swizzle) and the UMMA descriptors read them with the channel index as the fast M/N
// dimension and 8-pixel groups as the K dimension -- no transpose pass anywhere.
//
// Work item = (two 64-channel P atoms, 64 Q-channels, K split): three accumulators (filter columns s=0..2) of
// 128 x 64 fp32 in TMEM.  The Q tile is loaded once per pixel tile WITH its horizontal halo (10 pixels per row); the
// three filter columns are three descriptors into that buffer (start address + s pixels, stride-byte-offset = 10
// pixels), as in the forward HALO mode.  The filter ROW is applied to the P tile instead: its TMA row coordinate is
// y0 + 1 - r (out-of-image rows are zero-filled).  With Ca >= 128 the two P atoms are channels [c0, c0+64) and
// [c0+64, c0+128) of one filter row; with Ca == 64 they are the SAME 64 channels for two different filter rows
// (rows 0 and 1; row 2 runs with an empty second atom), so the M = 128 MMA is 75 % instead of 50 % useful.
// WIDE mode issues ONE N=192 MMA instead of three N=64 ones: the three 64-channel N atoms are the same buffer at
// leading-byte-offset = 1 pixel (overlapping atoms), which halves the A-operand shared-memory reads.
#pragma once
#include "ptx.cuh"

namespace fb {

struct WgradParams {
  int G, B, H, W;
  int Ca, Cb;   // channels of P (rows of dW) and of Q (padded to a multiple of 64)
  int bh, bn;   // pixel tile = bn images x bh rows x 8 columns (bh * bn == 16)
  int tiles_x, tiles_y, tiles_b;
  int m_tiles, n_chunks, splits;
  int row_items;  // filter-row work items per (m_tile, n_chunk): 3 (Ca >= 128) or 2 (Ca == 64: rows {0,1} and {2})
  int tiles_total;  // pixel tiles = G * tiles_b * tiles_y * tiles_x
  int stages;
  float* ws;  // [splits][Ca][9][Cb] fp32 partial sums
};

constexpr int kWgThreads = 192;
constexpr int kWgPBytes = 2 * 16384;             // two 64-channel atoms of 128 pixels
constexpr int kWgPHaloBytes = 18 * 1024;         // HP: ONE 64-channel tile of 18 rows x 8 pixels (the 16 rows + a row above / below)
// Q stage: 160 pixel slots (16 rows x 10 columns) x QCK channels; QCK = 16 is the 13-band stem (32B swizzle)
__host__ __device__ constexpr int wg_q_bytes(int QCK) { return 160 * QCK * 2; }
__host__ __device__ constexpr int wg_stage_bytes(int QCK, bool HP = false) { return (HP ? kWgPHaloBytes : kWgPBytes) + wg_q_bytes(QCK); }

// HP ("halo P", Ca == 64 on maps of 16 rows or more): ONE work item per (Q chunk, K split) covers all three filter rows.  The P
// tile is loaded once WITH a row above and below (18 rows x 8 pixels x 64 channels); a filter row is a start-address offset of
// one tile row (1024 B) into that buffer, and the two M atoms of an MMA are the same buffer one row apart (leading-byte-offset
// 1024: atom 0 = filter row 1, atom 1 = filter row 0; a second MMA takes filter row 2, its upper half is not stored).  The plain
// form loads the P tile once per filter row: 3 x 16 KB + 2 Q tiles per pixel tile across its two items -- the 13-band stem
// (N = 48, almost no math per byte) was bound by exactly that L2 -> shared-memory traffic (ncu: 3.75 GB in 0.38 ms).
template <int QCK, bool WIDE, bool HP = false>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_umma_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmQ, const WgradParams p) {
  static_assert(QCK == 64 || QCK == 16, "Q channel chunk");
  static_assert(!HP || WIDE, "the halo-P form issues the wide (three filter columns) MMA");
  constexpr int kWgStage = wg_stage_bytes(QCK, HP);
  constexpr int kPBytes = HP ? kWgPHaloBytes : kWgPBytes;
  constexpr int kTmemCols = HP ? (QCK == 64 ? 512 : 128) : 256;
  constexpr uint32_t PIX = QCK * 2;                 // bytes per pixel of the Q tile
  constexpr uint32_t Q_LAYOUT = QCK == 64 ? kLayoutSw128 : kLayoutSw32;
  constexpr int NQ = QCK;                           // N per filter column
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_u32);
  const uint32_t bar_off = p.stages * kWgStage;
  const uint32_t bars = base + bar_off;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (p.stages + s); };
  const uint32_t done_bar = bars + 8u * (2 * p.stages);
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(sm + bar_off + 512);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // work item
  int it = blockIdx.x;
  const int split = it % p.splits;
  it /= p.splits;
  const int ri = it % p.row_items;
  it /= p.row_items;
  const int nc = it % p.n_chunks;
  const int mt = it / p.n_chunks;
  // the two P atoms: (channel offset, filter row); row < 0 = empty atom
  const bool ca64 = !HP && p.row_items == 2;
  const int ra = ca64 ? (ri == 0 ? 0 : 2) : ri;
  const int rb = ca64 ? (ri == 0 ? 1 : -1) : ri;
  const int ca_a = mt * 128, ca_b = ca64 ? 0 : mt * 128 + 64;
  const int t_begin = (int)((long long)p.tiles_total * split / p.splits);
  const int t_end = (int)((long long)p.tiles_total * (split + 1) / p.splits);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmP);
    prefetch_tmap(&tmQ);
  }
  if (warp == 1) {
    tmem_alloc(base + bar_off + 512, kTmemCols);
    tmem_relinquish();
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(full(s), 1);
        mbar_init(empty(s), 1);
      }
      mbar_init(done_bar, 1);
      fence_mbar_init();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      int m = t;
      const int tx = m % p.tiles_x;
      m /= p.tiles_x;
      const int ty = m % p.tiles_y;
      m /= p.tiles_y;
      const int tb = m % p.tiles_b;
      const int g = m / p.tiles_b;
      const int x0 = tx * 8, y0 = ty * p.bh, b0 = tb * p.bn;
      mbar_wait(empty(stage), phase ^ 1);
      if (elect_one()) {
        const uint32_t dst = base + stage * kWgStage;
        mbar_arrive_expect_tx(full(stage), kWgStage);  // P: 2 x 16 KB (HP: 18 KB), Q: 160 pixel slots
        if constexpr (HP) {
          tma_load_5d(dst, &tmP, full(stage), 0, x0, y0 - 1, b0, g);   // rows y0-1 .. y0+16 (out of image = zero filled)
        } else {
          tma_load_5d(dst, &tmP, full(stage), ca_a, x0, y0 + 1 - ra, b0, g);
          // empty atom: a channel coordinate beyond Ca makes the whole box out of bounds = zero filled
          tma_load_5d(dst + 16384, &tmP, full(stage), rb < 0 ? p.Ca : ca_b, x0, y0 + 1 - (rb < 0 ? 0 : rb), b0, g);
        }
        tma_load_5d(dst + kPBytes, &tmQ, full(stage), nc * QCK, x0 - 1, y0, b0, g);
      }
      __syncwarp();
      if (++stage == p.stages) stage = 0, phase ^= 1;
    }
  } else if (warp == 1) {
    // A: M = P channels (MN-major), 2 atoms of 64 at LBO = 16384, K groups of 8 pixels at SBO = 1024
    // B: N = Q channels (MN-major), K groups of 8 pixels at SBO = 1280 (one halo row of 10 pixels)
    constexpr uint32_t idesc = umma_idesc_bf16(128, WIDE ? 3 * NQ : NQ, 1, 1);
    constexpr uint32_t B_LBO = WIDE ? PIX : 16384;   // WIDE: the next N atom is the same buffer one pixel further
    constexpr uint32_t B_SBO = 10 * PIX;             // 8-pixel K group stride = one halo row
    constexpr uint32_t A_LBO = HP ? 1024 : 16384;    // HP: the second M atom is the same buffer one tile row further
    constexpr uint64_t a_hi = umma_desc(0, A_LBO, 1024, kLayoutSw128) & 0xFFFFFFFF00000000ull;
    constexpr uint64_t b_hi = umma_desc(0, B_LBO, B_SBO, Q_LAYOUT) & 0xFFFFFFFF00000000ull;
    constexpr uint32_t a_lo_fixed = static_cast<uint32_t>(umma_desc(0, A_LBO, 0, 0) & 0xFFFFFFFFu);
    constexpr uint32_t b_lo_fixed = static_cast<uint32_t>(umma_desc(0, B_LBO, 0, 0) & 0xFFFFFFFFu);
    constexpr uint32_t B_JSTEP = (2 * B_SBO) >> 4, B_SSTEP = PIX >> 4;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t accumulate = 0;
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(full(stage), phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_lo = a_lo_fixed + ((base + stage * kWgStage) >> 4);
        const uint32_t b_lo = b_lo_fixed + ((base + stage * kWgStage + kPBytes) >> 4);
#pragma unroll
        for (int j = 0; j < 8; ++j) {  // 16 pixels (two tile rows) per MMA
          if constexpr (HP) {
            // P halo row = Q row - r + 2: filter rows (1, 0) start one row down, filter row 2 at the top of the buffer
            umma_bf16(tmem_base, a_hi | (a_lo + j * 128 + 64), b_hi | (b_lo + j * B_JSTEP), idesc, accumulate);
            umma_bf16(tmem_base + 3 * NQ, a_hi | (a_lo + j * 128), b_hi | (b_lo + j * B_JSTEP), idesc, accumulate);
          } else if constexpr (WIDE) {
            umma_bf16(tmem_base, a_hi | (a_lo + j * 128), b_hi | (b_lo + j * B_JSTEP), idesc, accumulate);
          } else {
#pragma unroll
            for (int s = 0; s < 3; ++s)
              umma_bf16(tmem_base + s * NQ, a_hi | (a_lo + j * 128), b_hi | (b_lo + j * B_JSTEP + s * B_SSTEP), idesc, accumulate);
          }
          accumulate = 1;
        }
        umma_commit(empty(stage));
        if (t == t_end - 1) umma_commit(done_bar);
      }
      __syncwarp();
      accumulate = 1;
      if (++stage == p.stages) stage = 0, phase ^= 1;
    }
  } else {
    // epilogue: 128 rows (P channels) x 3 taps x 64 Q channels -> fp32 partial slab of this split
    const int q = warp & 3;
    const int row = q * 32 + lane;
    if (t_end > t_begin) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
    }
    // HP: two accumulator blocks; block 0 rows 0-63 = filter row 1, rows 64-127 = filter row 0; block 1 rows 0-63 = filter row 2
#pragma unroll 1
    for (int blk = 0; blk < (HP ? 2 : 1); ++blk) {
      // accumulator row -> (dW row, filter row)
      const int r = HP ? (blk == 0 ? (row < 64 ? 1 : 0) : (row < 64 ? 2 : -1)) : (row < 64 ? ra : rb);
      const int ca = HP ? (row & 63) : (row < 64 ? ca_a + row : ca_b + (row - 64));
#pragma unroll 1
      for (int s = 0; s < 3; ++s) {
#pragma unroll 1
        for (int cc = 0; cc < (NQ + 31) / 32; ++cc) {
          uint32_t v[32];
          if (t_end > t_begin) {
            // (for NQ = 16 this reads 16 columns past the tap's accumulator; they are simply not stored)
            tmem_ld_32x32(tmem_base + blk * 3 * NQ + s * NQ + cc * 32 + (static_cast<uint32_t>(q * 32) << 16), v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0u;
          }
          if (r >= 0 && ca < p.Ca) {
            float4* dst = reinterpret_cast<float4*>(p.ws + (((size_t)split * p.Ca + ca) * 9 + (r * 3 + s)) * p.Cb + nc * QCK + cc * 32);
#pragma unroll
            for (int i = 0; i < (NQ < 32 ? NQ : 32) / 4; ++i)
              dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                                   __uint_as_float(v[4 * i + 3]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Second form (used for Cb % 64 == 0 on maps of 16 rows or more): the filter ROW is applied through the Q tile as well.
// Work item = (128 P channels, 32 Q channels, K split); per pixel tile ONE P tile (rows y0..y0+15) and ONE Q tile with its
// full halo (18 rows x 10 columns x 32 channels, 64B swizzle) feed all nine taps: for each 16-pixel K slice j and filter row
// r one N = 96 MMA (three 32-channel N atoms = the same buffer at leading-byte-offset one pixel, start address at halo row
// 2j + r) into accumulator r.  The first form re-reads 52 KB from L2 per (filter row, 64 Q channels) item -- 121 flop per L2
// byte, L2-bound at ~10 TB/s -- this one moves 43.5 KB per 32 Q channels for all three rows: 217 flop per byte.
constexpr int kWg2PBytes = 2 * 16384;
constexpr int kWg2QBytes = 12288;                       // 18 x 10 pixels x 64 B = 11520, rounded up to 1 KB
constexpr int kWg2Stage = kWg2PBytes + kWg2QBytes;

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad2_umma_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmQ, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base = (raw_u32 + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw_u32);
  const uint32_t bar_off = p.stages * kWg2Stage;
  const uint32_t bars = base + bar_off;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (p.stages + s); };
  const uint32_t done_bar = bars + 8u * (2 * p.stages);
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(sm + bar_off + 512);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int it = blockIdx.x;
  const int split = it % p.splits;
  it /= p.splits;
  const int nc = it % p.n_chunks;       // 32-channel Q chunk
  const int mt = it / p.n_chunks;       // 128-channel P tile
  const int ca_a = mt * 128;
  const int ca_b = ca_a + 64 < p.Ca ? ca_a + 64 : p.Ca;   // beyond Ca: the box is out of bounds = zero filled
  const int t_begin = (int)((long long)p.tiles_total * split / p.splits);
  const int t_end = (int)((long long)p.tiles_total * (split + 1) / p.splits);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmP);
    prefetch_tmap(&tmQ);
  }
  if (warp == 1) {
    tmem_alloc(base + bar_off + 512, 512);
    tmem_relinquish();
    if (lane == 0) {
      for (int s = 0; s < p.stages; ++s) {
        mbar_init(full(s), 1);
        mbar_init(empty(s), 1);
      }
      mbar_init(done_bar, 1);
      fence_mbar_init();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      int m = t;
      const int tx = m % p.tiles_x;
      m /= p.tiles_x;
      const int ty = m % p.tiles_y;
      m /= p.tiles_y;
      const int tb = m % p.tiles_b;
      const int g = m / p.tiles_b;
      const int x0 = tx * 8, y0 = ty * 16, b0 = tb;
      mbar_wait(empty(stage), phase ^ 1);
      if (elect_one()) {
        const uint32_t dst = base + stage * kWg2Stage;
        mbar_arrive_expect_tx(full(stage), kWg2PBytes + 18 * 10 * 64);
        tma_load_5d(dst, &tmP, full(stage), ca_a, x0, y0, b0, g);
        tma_load_5d(dst + 16384, &tmP, full(stage), ca_b, x0, y0, b0, g);
        tma_load_5d(dst + kWg2PBytes, &tmQ, full(stage), nc * 32, x0 - 1, y0 - 1, b0, g);
      }
      __syncwarp();
      if (++stage == p.stages) stage = 0, phase ^= 1;
    }
  } else if (warp == 1) {
    // A: M = P channels (MN-major, 128B swizzle), 2 atoms of 64 at LBO = 16384, K groups of 8 pixels at SBO = 1024
    // B: N = 3 x 32 Q channels (MN-major, 64B swizzle), N atoms at LBO = 64 B (one pixel), K groups at SBO = 640 B (halo row)
    constexpr uint32_t idesc = umma_idesc_bf16(128, 96, 1, 1);
    constexpr uint64_t a_hi = umma_desc(0, 16384, 1024, kLayoutSw128) & 0xFFFFFFFF00000000ull;
    constexpr uint64_t b_hi = umma_desc(0, 64, 640, kLayoutSw64) & 0xFFFFFFFF00000000ull;
    constexpr uint32_t a_lo_fixed = static_cast<uint32_t>(umma_desc(0, 16384, 0, 0) & 0xFFFFFFFFu);
    constexpr uint32_t b_lo_fixed = static_cast<uint32_t>(umma_desc(0, 64, 0, 0) & 0xFFFFFFFFu);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t accumulate = 0;
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(full(stage), phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_lo = a_lo_fixed + ((base + stage * kWg2Stage) >> 4);
        const uint32_t b_lo = b_lo_fixed + ((base + stage * kWg2Stage + kWg2PBytes) >> 4);
#pragma unroll
        for (int j = 0; j < 8; ++j) {       // 16 pixels (two tile rows) per K slice
#pragma unroll
          for (int r = 0; r < 3; ++r)       // filter row: the Q window starts r halo rows further down
            umma_bf16(tmem_base + r * 96, a_hi | (a_lo + j * 128), b_hi | (b_lo + (2 * j + r) * 40), idesc, accumulate);
          accumulate = 1;
        }
        umma_commit(empty(stage));
        if (t == t_end - 1) umma_commit(done_bar);
      }
      __syncwarp();
      accumulate = 1;
      if (++stage == p.stages) stage = 0, phase ^= 1;
    }
  } else {
    // epilogue: 128 rows (P channels) x 9 taps x 32 Q channels -> fp32 partial slab of this split
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int ca = ca_a + row;
    if (t_end > t_begin) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
      uint32_t v[32];
      if (t_end > t_begin) {
        tmem_ld_32x32(tmem_base + (tap / 3) * 96 + (tap % 3) * 32 + (static_cast<uint32_t>(q * 32) << 16), v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0u;
      }
      if (ca < p.Ca) {
        float4* dst = reinterpret_cast<float4*>(p.ws + (((size_t)split * p.Ca + ca) * 9 + tap) * p.Cb + nc * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]),
                               __uint_as_float(v[4 * i + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace fb
