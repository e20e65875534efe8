// Training-path kernels of fabric_b200 that are HBM-bound (everything except the convolutions):
// BatchNorm batch statistics / apply / backward, the segmentation losses and their gradient, the 1x1 head backward,
// and the adjoint of the decoder-input builder.  See include/fabric_b200.h for the reference constructs replaced.
#include "host_common.cuh"
#include "ptx.cuh"

using namespace fbh;
using fb::pack8;
using fb::unpack8;

namespace {

__device__ __forceinline__ void ld8f(const float* __restrict__ p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}

// ================================================================================================ BN forward
// reference models/unet_parts.py:14,17 (nn.BatchNorm2d in training mode), per date group.
__global__ void bn_finalize_kernel(const float* __restrict__ stats, int grid, int n_tile, int C, int G, double count,
                                   const float* __restrict__ conv_bias, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* running_mean, float* running_var,
                                   long long* nbt, float momentum, float eps, float* scale, float* shift, float* mean,
                                   float* invstd) {
  // one warp per channel: lanes stride over the CTAs' partials
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  const int ntiles = C / n_tile, nt = c / n_tile, lc = c % n_tile;
  float rm = running_mean[c], rv = running_var[c];
  const float b = conv_bias ? conv_bias[c] : 0.f;
  for (int g = 0; g < G; ++g) {  // date 1 first, then date 2: the order in which the reference calls the encoder
    double s1 = 0.0, s2 = 0.0;
    for (int cta = nt + lane * ntiles; cta < grid; cta += 32 * ntiles) {
      const float2 v = *reinterpret_cast<const float2*>(stats + (((size_t)cta * 2 + g) * n_tile + lc) * 2);
      s1 += v.x;
      s2 += v.y;
    }
    for (int o = 16; o; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const double m = s1 / count;
    double var = s2 / count - m * m;
    if (var < 0.0) var = 0.0;
    const float inv = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * inv;
    if (lane == 0) {
      scale[g * C + c] = sc;
      shift[g * C + c] = beta[c] - (float)m * sc;  // the conv bias cancels against the batch mean
      mean[g * C + c] = (float)m;
      invstd[g * C + c] = inv;
    }
    rm = (1.f - momentum) * rm + momentum * ((float)m + b);
    rv = (1.f - momentum) * rv + momentum * (float)(var * (count / (count > 1.0 ? count - 1.0 : 1.0)));
  }
  if (lane == 0) {
    running_mean[c] = rm;
    running_var[c] = rv;
    if (c == 0 && nbt) *nbt += G;
  }
}

// a = relu(z * scale_g + shift_g) (+ 2x2 max pool) (+ relu(a_d2 * a_d1) into the decoder input).  One thread = 8 channels
// of one 2x2 pixel quad, for BOTH date groups when the product is fused; 32-bit indexing.
// EVEN: H and W even (every real level): no bounds tests, the quad is one basic block.
template <bool EVEN>
__global__ void __launch_bounds__(256, 3) bn_apply_kernel(const uint4* __restrict__ z, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, uint4* __restrict__ a,
                                                          uint4* __restrict__ pool, uint4* __restrict__ prod, int prod_c8,
                                                          int G, int B, int H, int W, int C) {
  const uint32_t C8 = C >> 3, Hq = (H + 1) >> 1, Wq = (W + 1) >> 1, Hp = H >> 1, Wp = W >> 1;
  const uint32_t GB = prod ? (uint32_t)B : (uint32_t)G * B;   // with the product fused a thread walks both date groups
  const uint32_t total = GB * Hq * Wq * C8;
  const uint32_t ngroups = prod ? 2u : 1u;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t c8 = i % C8;
    uint32_t q = i / C8;
    const uint32_t qx = q % Wq;
    q /= Wq;
    const uint32_t qy = q % Hq;
    q /= Hq;  // = g * B + b (or b with the product fused)
    uint4 first[4];
    for (uint32_t gi = 0; gi < ngroups; ++gi) {
      const uint32_t img = q + gi * B;                       // image index in [0, G*B)
      const uint32_t g = img >= (uint32_t)B ? 1u : 0u;
      float sc[8], sh[8], m[8];
      ld8f(scale + g * C + c8 * 8, sc);
      ld8f(shift + g * C + c8 * 8, sh);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = 0.f;  // post-ReLU values are >= 0
      uint4 zin[4];
      bool ok[4];
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const uint32_t y = 2 * qy + (d >> 1), x = 2 * qx + (d & 1);
        ok[d] = EVEN || (y < (uint32_t)H && x < (uint32_t)W);
        if (ok[d]) zin[d] = __ldg(z + (size_t)((img * H + y) * W + x) * C8 + c8);
      }
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        if (!ok[d]) continue;
        const uint32_t y = 2 * qy + (d >> 1), x = 2 * qx + (d & 1);
        float f[8];
        unpack8(zin[d], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          f[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
          m[j] = fmaxf(m[j], f[j]);
        }
        const uint4 av = pack8(f);
        if (a) a[(size_t)((img * H + y) * W + x) * C8 + c8] = av;
        if (prod) {
          if (gi == 0) {
            first[d] = av;
          } else {   // relu(a_d2 * a_d1) on the stored (bf16) activations, bidate_model.py:35-38
            float f0[8], f1[8], r[8];
            unpack8(first[d], f0);
            unpack8(av, f1);
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = fmaxf(f0[j] * f1[j], 0.f);
            prod[(size_t)((q * H + y) * W + x) * prod_c8 + c8] = pack8(r);
          }
        }
      }
      if (pool && (EVEN || (qy < Hp && qx < Wp))) pool[(size_t)((img * Hp + qy) * Wp + qx) * C8 + c8] = pack8(m);
    }
  }
}


// ================================================================================================ fused 1x1 head (training)
// The last decoder BatchNorm (up4, 64 channels) and `outconv` (unet_parts.py:83-90, bidate_model.py:39) as ONE pass each way.
// Forward: a = relu(z*scale+shift) is stored (the next step's wgrad needs nothing of it, but backward recomputes it) and the
// logits = W a + b leave in the same pass.  One thread = 8 channels of one pixel; the 8 threads of a pixel are adjacent lanes.
__global__ void __launch_bounds__(256) bn_apply_head_kernel(const uint4* __restrict__ z, const float* __restrict__ scale,
                                                            const float* __restrict__ shift, uint4* __restrict__ a,
                                                            const float* __restrict__ hw, const float* __restrict__ hb,
                                                            float* __restrict__ logits, uint32_t npix, uint32_t plane) {
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t c8 = gtid & 7u;
  float sc[8], sh[8], w0[8], w1[8];
  ld8f(scale + c8 * 8, sc);
  ld8f(shift + c8 * 8, sh);
  ld8f(hw + c8 * 8, w0);
  ld8f(hw + 64 + c8 * 8, w1);
  const float b0 = hb[0], b1 = hb[1];
  const uint32_t pstride = (gridDim.x * blockDim.x) >> 3;
  constexpr int NB = 4;   // pixels in flight per thread (one 16-byte load each: with a single one the kernel sat at 2.9 TB/s)
  for (uint32_t p0 = gtid >> 3; p0 < npix; p0 += NB * pstride) {
    const unsigned lanes = __activemask();   // the 8 lanes of a pixel enter and leave the loop together
    uint4 zr[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k)
      if (p0 + k * pstride < npix) zr[k] = __ldg(z + (size_t)(p0 + k * pstride) * 8 + c8);
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      const uint32_t pix = p0 + k * pstride;
      const bool live = pix < npix;      // uniform over the 8 lanes of a pixel; the shuffles below run for every k
      float f[8], q[8];
      unpack8(live ? zr[k] : make_uint4(0u, 0u, 0u, 0u), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
      const uint4 av = pack8(f);
      if (live && a) a[(size_t)pix * 8 + c8] = av;
      unpack8(av, q);   // the head sees the activation as stored
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        l0 = fmaf(q[j], w0[j], l0);
        l1 = fmaf(q[j], w1[j], l1);
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        l0 += __shfl_xor_sync(lanes, l0, o);
        l1 += __shfl_xor_sync(lanes, l1, o);
      }
      if (live && c8 == 0) {
        const uint32_t b = pix / plane, o = pix - b * plane;
        logits[(size_t)(b * 2) * plane + o] = l0 + b0;
        logits[(size_t)(b * 2 + 1) * plane + o] = l1 + b1;
      }
    }
  }
}

// Backward of the same pair: du = W^T dlogits is never materialised.  dy = relu'(.) * du;
// pass 1: partial[blk] = { (sum dy, sum dy*xhat)[64][2], dW[2][64], db[2] };  pass 2: dz = k0*dy + kz*z + kc.
struct HeadBwd {
  const uint4* z;
  const float* dlogits;
  const float* scale;
  const float* shift;
  const float* mean;
  const float* invstd;
  const float* hw;
  uint32_t npix, plane;
  unsigned long long plane_magic;   // ceil(2^40 / plane) when npix * plane < 2^40 (pix / plane as one multiply), else 0
};
__device__ __forceinline__ uint32_t head_image(const HeadBwd& p, uint32_t pix) {
  return p.plane_magic ? (uint32_t)(((unsigned long long)pix * p.plane_magic) >> 40) : pix / p.plane;
}
constexpr int kHeadPartial = 64 * 2 + 2 * 64 + 2;

__global__ void __launch_bounds__(256, 2) bn_head_bwd_reduce_kernel(HeadBwd p, float* __restrict__ partial) {
  extern __shared__ float sm[];  // [blockDim][34]
  const uint32_t c8 = threadIdx.x & 7u, lane_p = threadIdx.x >> 3, ppb = blockDim.x >> 3;
  float sc[8], sh[8], mu[8], is[8], w0[8], w1[8];
  ld8f(p.scale + c8 * 8, sc);
  ld8f(p.shift + c8 * 8, sh);
  ld8f(p.mean + c8 * 8, mu);
  ld8f(p.invstd + c8 * 8, is);
  ld8f(p.hw + c8 * 8, w0);
  ld8f(p.hw + 64 + c8 * 8, w1);
  float s1[8], s2[8], g0[8], g1[8], d0s = 0.f, d1s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = g0[j] = g1[j] = 0.f;
  const uint32_t stride = gridDim.x * ppb;
  constexpr int NB = 4;
  auto load = [&](uint32_t pix, uint4& zr, float& dl0, float& dl1) {
    const uint32_t b = head_image(p, pix), o = pix - b * p.plane;
    zr = __ldg(p.z + (size_t)pix * 8 + c8);
    dl0 = __ldg(p.dlogits + (size_t)(b * 2) * p.plane + o);
    dl1 = __ldg(p.dlogits + (size_t)(b * 2 + 1) * p.plane + o);
  };
  auto one = [&](const uint4& zr, float d0, float d1) {
    float zf[8], pre[8];
    unpack8(zr, zf);
#pragma unroll
    for (int j = 0; j < 8; ++j) pre[j] = fmaf(zf[j], sc[j], sh[j]);
    float aq[8];
    // as stored by the forward pass: max(x, 0) rides in the bf16 conversion
    unpack8(make_uint4(fb::pack_bf16x2_relu(pre[0], pre[1]), fb::pack_bf16x2_relu(pre[2], pre[3]), fb::pack_bf16x2_relu(pre[4], pre[5]),
                       fb::pack_bf16x2_relu(pre[6], pre[7])), aq);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float dy = pre[j] > 0.f ? fmaf(d0, w0[j], d1 * w1[j]) : 0.f;
      s1[j] += dy;
      s2[j] = fmaf(dy, (zf[j] - mu[j]) * is[j], s2[j]);
      g0[j] = fmaf(d0, aq[j], g0[j]);
      g1[j] = fmaf(d1, aq[j], g1[j]);
    }
    if (c8 == 0) d0s += d0, d1s += d1;
  };
  uint32_t p0 = blockIdx.x * ppb + lane_p;
  for (; p0 + (NB - 1) * stride < p.npix; p0 += NB * stride) {   // main loop: four pixels in flight, no bounds tests
    uint4 zr[NB];
    float dl0[NB], dl1[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) load(p0 + k * stride, zr[k], dl0[k], dl1[k]);
#pragma unroll
    for (int k = 0; k < NB; ++k) one(zr[k], dl0[k], dl1[k]);
  }
  for (; p0 < p.npix; p0 += stride) {
    uint4 zr;
    float d0, d1;
    load(p0, zr, d0, d1);
    one(zr, d0, d1);
  }
  float* mine = sm + threadIdx.x * 34;
#pragma unroll
  for (int j = 0; j < 8; ++j) mine[j] = s1[j], mine[8 + j] = s2[j], mine[16 + j] = g0[j], mine[24 + j] = g1[j];
  mine[32] = d0s, mine[33] = d1s;
  __syncthreads();
  float* dst = partial + (size_t)blockIdx.x * kHeadPartial;
  for (int i = threadIdx.x; i < kHeadPartial; i += blockDim.x) {
    float s = 0.f;
    if (i < 128) {          // (sum dy, sum dy*xhat) interleaved per channel
      const int c = i >> 1, k = i & 1;
      for (uint32_t l = 0; l < ppb; ++l) s += sm[(l * 8 + (c >> 3)) * 34 + k * 8 + (c & 7)];
    } else if (i < 256) {   // dW[k][c]
      const int k = (i - 128) >> 6, c = (i - 128) & 63;
      for (uint32_t l = 0; l < ppb; ++l) s += sm[(l * 8 + (c >> 3)) * 34 + 16 + k * 8 + (c & 7)];
    } else {
      for (uint32_t l = 0; l < ppb; ++l) s += sm[(l * 8) * 34 + 32 + (i - 256)];
    }
    dst[i] = s;
  }
}

// one block: reduce the partials, finish BatchNorm's dgamma / dbeta / coefficients and the head's dW / db
__global__ void bn_head_bwd_finalize_kernel(const float* __restrict__ partial, int nblk, double count, const float* __restrict__ gamma,
                                            const float* __restrict__ invstd, const float* __restrict__ mean,
                                            float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ coef,
                                            float* __restrict__ dw, float* __restrict__ db, float grad_scale) {
  for (int i = threadIdx.x; i < kHeadPartial; i += blockDim.x) {
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += partial[(size_t)b * kHeadPartial + i];
    if (i >= 128 && i < 256) dw[i - 128] = (float)s * grad_scale;
    else if (i >= 256) db[i - 256] = (float)s * grad_scale;
    else coef[3 * 64 + i] = (float)s;   // stash the sums behind the coefficients
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 64; c += blockDim.x) {
    const double s1 = coef[3 * 64 + 2 * c], s2 = coef[3 * 64 + 2 * c + 1];
    const float is = invstd[c], mu = mean[c];
    const float k0 = gamma[c] * is, k1 = (float)(s1 / count), k2 = (float)(s2 / count);
    coef[c] = k0;
    coef[64 + c] = -k0 * is * k2;
    coef[128 + c] = -k0 * (k1 - mu * is * k2);
    dgamma[c] = (float)s2 * grad_scale;
    dbeta[c] = (float)s1 * grad_scale;
  }
}

__global__ void __launch_bounds__(256, 2) bn_head_bwd_apply_kernel(HeadBwd p, const float* __restrict__ coef, uint4* __restrict__ dz) {
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t c8 = gtid & 7u;
  float sc[8], sh[8], w0[8], w1[8], k0[8], kz[8], kc[8];
  ld8f(p.scale + c8 * 8, sc);
  ld8f(p.shift + c8 * 8, sh);
  ld8f(p.hw + c8 * 8, w0);
  ld8f(p.hw + 64 + c8 * 8, w1);
  ld8f(coef + c8 * 8, k0);
  ld8f(coef + 64 + c8 * 8, kz);
  ld8f(coef + 128 + c8 * 8, kc);
  const uint32_t pstride = (gridDim.x * blockDim.x) >> 3;
  constexpr int NB = 4;
  auto load = [&](uint32_t pix, uint4& zr, float& dl0, float& dl1) {
    const uint32_t b = head_image(p, pix), o = pix - b * p.plane;
    zr = __ldg(p.z + (size_t)pix * 8 + c8);
    dl0 = __ldg(p.dlogits + (size_t)(b * 2) * p.plane + o);
    dl1 = __ldg(p.dlogits + (size_t)(b * 2 + 1) * p.plane + o);
  };
  auto one = [&](uint32_t pix, const uint4& zr, float d0, float d1) {
    float zf[8], r[8];
    unpack8(zr, zf);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float dy = fmaf(zf[j], sc[j], sh[j]) > 0.f ? fmaf(d0, w0[j], d1 * w1[j]) : 0.f;
      r[j] = fmaf(k0[j], dy, fmaf(kz[j], zf[j], kc[j]));
    }
    dz[(size_t)pix * 8 + c8] = pack8(r);
  };
  uint32_t p0 = gtid >> 3;
  for (; p0 + (NB - 1) * pstride < p.npix; p0 += NB * pstride) {
    uint4 zr[NB];
    float dl0[NB], dl1[NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) load(p0 + k * pstride, zr[k], dl0[k], dl1[k]);
#pragma unroll
    for (int k = 0; k < NB; ++k) one(p0 + k * pstride, zr[k], dl0[k], dl1[k]);
  }
  for (; p0 < p.npix; p0 += pstride) {
    uint4 zr;
    float d0, d1;
    load(p0, zr, d0, d1);
    one(p0, zr, d0, d1);
  }
}

// ================================================================================================ losses
// utils/metrics.py:51-171 (dice / jaccard / tversky share one front end) and :19-48 (focal); C = 2.
// pass 1: per block, per image column w: I_c = sum p_c t_c, P_c = sum p_c, T_c = sum t_c over the block's rows
__global__ void seg_loss_partial_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, int B, int H,
                                        int W, int rows_per_block, float* __restrict__ partial) {
  const int rows = B * H;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  const size_t plane = (size_t)H * W;
  for (int w = threadIdx.x; w < W; w += blockDim.x) {
    float I0 = 0, I1 = 0, P0 = 0, P1 = 0, T0 = 0, T1 = 0;
    for (int r = r0; r < r1; ++r) {
      const int b = r / H, h = r % H;
      const size_t o = (size_t)b * 2 * plane + (size_t)h * W + w;
      const float l0 = logits[o], l1 = logits[o + plane];
      const float p1 = 1.f / (1.f + expf(l0 - l1)), p0 = 1.f - p1;
      const bool t = labels[(size_t)b * plane + (size_t)h * W + w] != 0;
      P0 += p0;
      P1 += p1;
      if (t) {
        I1 += p1;
        T1 += 1.f;
      } else {
        I0 += p0;
        T0 += 1.f;
      }
    }
    float* dst = partial + (size_t)blockIdx.x * 6 * W;
    dst[0 * W + w] = I0, dst[1 * W + w] = I1, dst[2 * W + w] = P0, dst[3 * W + w] = P1, dst[4 * W + w] = T0, dst[5 * W + w] = T1;
  }
}

__device__ float block_sum(float v, float* red) {
  __syncthreads();
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < (blockDim.x + 31) / 32; ++i) s += red[i];
  return s;
}

// pass 2 (one block): sums -> loss and the coefficients dL/dI_c[w], dL/dP_c[w]
// kind: 0 tversky, 1 dice, 2 jaccard.  label_ndim 3: terms are (class, column) (reference `dims` quirk), 4: (class).
__global__ void seg_loss_finalize_kernel(const float* __restrict__ partial, int nblk, int W, int label_ndim, int kind,
                                         float alpha, float beta, float eps, float* __restrict__ coef, float* loss_out) {
  __shared__ float red[32];
  __shared__ float tot[6];
  extern __shared__ float S[];  // [6][W]
  for (int w = threadIdx.x; w < W; w += blockDim.x)
    for (int k = 0; k < 6; ++k) {
      float s = 0.f;
      for (int b = 0; b < nblk; ++b) s += partial[((size_t)b * 6 + k) * W + w];
      S[k * W + w] = s;
    }
  __syncthreads();
  if (label_ndim == 4) {
    for (int k = 0; k < 6; ++k) {
      float v = 0.f;
      for (int w = threadIdx.x; w < W; w += blockDim.x) v += S[k * W + w];
      v = block_sum(v, red);
      if (threadIdx.x == 0) tot[k] = v;
    }
    __syncthreads();
  }
  const float nterms = label_ndim == 4 ? 2.f : 2.f * W;
  float acc = 0.f;
  for (int w = threadIdx.x; w < W; w += blockDim.x) {
    for (int c = 0; c < 2; ++c) {
      const float I = label_ndim == 4 ? tot[c] : S[c * W + w];
      const float P = label_ndim == 4 ? tot[2 + c] : S[(2 + c) * W + w];
      const float T = label_ndim == 4 ? tot[4 + c] : S[(4 + c) * W + w];
      float r, dI, dP;
      if (kind == 0) {
        const float den = I + alpha * (P - I) + beta * (T - I) + eps;
        r = I / den;
        dI = (den - I * (1.f - alpha - beta)) / (den * den);
        dP = -I * alpha / (den * den);
      } else if (kind == 1) {
        const float den = P + T + eps;
        r = 2.f * I / den;
        dI = 2.f / den;
        dP = -2.f * I / (den * den);
      } else {
        const float den = P + T - I + eps;
        r = I / den;
        dI = (den + I) / (den * den);
        dP = -I / (den * den);
      }
      if (label_ndim == 3 || w == 0) acc += r;
      coef[c * W + w] = -dI / nterms;        // dL/dI_c[w]
      coef[(2 + c) * W + w] = -dP / nterms;  // dL/dP_c[w]
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) *loss_out = 1.f - acc / nterms;
}

// pass 3: dL/dlogits through the 2-class softmax
__global__ void seg_loss_grad_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                                     const float* __restrict__ coef, int B, int H, int W, float* __restrict__ dlogits) {
  const size_t plane = (size_t)H * W, total = (size_t)B * plane;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / plane, o = i % plane;
    const int w = o % W;
    const size_t lo = b * 2 * plane + o;
    const float l0 = logits[lo], l1 = logits[lo + plane];
    const float p1 = 1.f / (1.f + expf(l0 - l1)), p0 = 1.f - p1;
    const bool t = labels[i] != 0;
    const float g0 = (t ? 0.f : coef[w]) + coef[2 * W + w];
    const float g1 = (t ? coef[W + w] : 0.f) + coef[3 * W + w];
    const float d = p0 * p1 * (g0 - g1);
    dlogits[lo] = d;
    dlogits[lo + plane] = -d;
  }
}

// focal (gamma) / cross entropy (gamma = 0): loss partial sums + gradient in one pass (pt detached, metrics.py:35)
__global__ void focal_loss_kernel(const float* __restrict__ logits, const long long* __restrict__ labels, int B, int H, int W,
                                  float gamma, float* __restrict__ dlogits, float* __restrict__ partial, float mean_scale) {
  __shared__ float red[32];
  const size_t plane = (size_t)H * W, total = (size_t)B * plane;
  const float invn = mean_scale / (float)total;
  float acc = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / plane, o = i % plane;
    const size_t lo = b * 2 * plane + o;
    const float l0 = logits[lo], l1 = logits[lo + plane];
    const bool t = labels[i] != 0;
    const float mx = fmaxf(l0, l1);
    const float lse = mx + logf(expf(l0 - mx) + expf(l1 - mx));
    const float logpt = (t ? l1 : l0) - lse;
    const float pt = expf(logpt);
    const float wgt = gamma == 0.f ? 1.f : powf(fmaxf(1.f - pt, 0.f), gamma);
    acc += -wgt * logpt;
    const float p0 = expf(l0 - lse), p1 = expf(l1 - lse);
    // d(-w logpt)/dl_k = -w (delta_kt - p_k)
    dlogits[lo] = -wgt * ((t ? 0.f : 1.f) - p0) * invn;
    dlogits[lo + plane] = -wgt * ((t ? 1.f : 0.f) - p1) * invn;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc * invn;
}

// out[j] = sum_i ws[i][j]
__global__ void reduce_partials_kernel(const float* __restrict__ ws, int n, int m, float* __restrict__ out) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) {
    // eight independent fp64 chains (fixed order: deterministic): one chain of ~300 dependent loads took 47 us
    double s[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    int i = 0;
    const float* col = ws + j;
    for (; i + 8 <= n; i += 8) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = __ldg(col + (size_t)(i + k) * m);   // all eight loads first
#pragma unroll
      for (int k = 0; k < 8; ++k) s[k] += (double)v[k];
    }
    for (; i < n; ++i) s[0] += ws[(size_t)i * m + j];
    out[j] = (float)(((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7])));
  }
}

// ================================================================================================ head backward
// outconv (unet_parts.py:86): du = dlogits^T W, dW = dlogits u^T, db = sum dlogits.  C = 64.
__global__ void outconv_bwd_kernel(const float* __restrict__ dlogits, const uint4* __restrict__ u, const float* __restrict__ w,
                                   uint4* __restrict__ du, float* __restrict__ partial, int B, int H, int W, int C) {
  extern __shared__ float sm[];  // [blockDim][18]
  const int C8 = C / 8;
  const int c8 = threadIdx.x % C8, lane_p = threadIdx.x / C8, ppb = blockDim.x / C8;
  const size_t plane = (size_t)H * W, total = (size_t)B * plane;
  float w0[8], w1[8], a0[8], a1[8];
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    w0[j] = w[c8 * 8 + j];
    w1[j] = w[C + c8 * 8 + j];
    a0[j] = a1[j] = 0.f;
  }
  for (size_t pix = (size_t)blockIdx.x * ppb + lane_p; pix < total; pix += (size_t)gridDim.x * ppb) {
    const size_t b = pix / plane, o = pix % plane;
    const float d0 = dlogits[(b * 2) * plane + o], d1 = dlogits[(b * 2 + 1) * plane + o];
    float f[8], r[8];
    unpack8(u[pix * C8 + c8], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      r[j] = d0 * w0[j] + d1 * w1[j];
      a0[j] = fmaf(d0, f[j], a0[j]);
      a1[j] = fmaf(d1, f[j], a1[j]);
    }
    du[pix * C8 + c8] = pack8(r);
    if (c8 == 0) s0 += d0, s1 += d1;
  }
  float* mine = sm + threadIdx.x * 18;
#pragma unroll
  for (int j = 0; j < 8; ++j) mine[j] = a0[j], mine[8 + j] = a1[j];
  mine[16] = s0, mine[17] = s1;
  __syncthreads();
  // partial layout: [2][C] dW then [2] db
  float* dst = partial + (size_t)blockIdx.x * (2 * C + 2);
  for (int i = threadIdx.x; i < 2 * C + 2; i += blockDim.x) {
    float s = 0.f;
    if (i < 2 * C) {
      const int k = i / C, c = i % C, cc8 = c / 8, j = c % 8;
      for (int pl = 0; pl < ppb; ++pl) s += sm[(pl * C8 + cc8) * 18 + k * 8 + j];
    } else {
      for (int pl = 0; pl < ppb; ++pl) s += sm[(pl * C8) * 18 + 16 + (i - 2 * C)];
    }
    dst[i] = s;
  }
}

// ================================================================================================ BN backward
// dy = relu'(z*scale+shift) * [ ga * (mul_other ? a[other date] : 1) + unpool(gp) ]   (see include/fabric_b200.h)
struct BnBwd {
  const uint4* z;
  const uint4* a;    // own activation (needed for mul_other / gp)
  const uint4* ga;   // nullable
  const uint4* gp;   // nullable
  const float* scale;
  const float* shift;
  const float* mean;
  const float* invstd;
  int ga_groups, ga_c8, mul_other;
  int G, B, H, W, C;
  int premasked;   // ga already is dy = relu'(.) * dL/da (written by the producing kernel's epilogue): no mask here
};

// dy for the 8 channels c8 of pixel `pix` (index inside one date group: (b*H + y)*W + x) of date group g.
// Split in two phases so that callers can issue the loads of several pixels before consuming any of them
// (memory-level parallelism: with one pixel per iteration these kernels sat at 2-2.7 TB/s).
// All index math is 32-bit (the host checks sizes): 64-bit div/mod per element made the first version ALU-bound.
struct BnBwdRaw {
  uint4 z, ga, ao;
};

__device__ __forceinline__ void bn_bwd_load(const BnBwd& p, uint32_t g, uint32_t pix, uint32_t c8, uint32_t npix, BnBwdRaw& r) {
  const uint32_t C8 = p.C >> 3;
  const uint32_t gpix = g * npix + pix;
  r.z = __ldg(p.z + (size_t)gpix * C8 + c8);
  if (p.ga) {
    r.ga = __ldg(p.ga + (size_t)(p.ga_groups == 1 ? pix : gpix) * p.ga_c8 + c8);
    if (p.mul_other) r.ao = __ldg(p.a + (size_t)((1 - g) * npix + pix) * C8 + c8);
  }
}

__device__ __forceinline__ void bn_bwd_finish(const BnBwd& p, uint32_t g, uint32_t pix, uint32_t c8, uint32_t npix,
                                              const BnBwdRaw& r, const float (&sc)[8], const float (&sh)[8], float (&dy)[8],
                                              float (&zf)[8]) {
  const uint32_t C8 = p.C >> 3;
  const uint32_t gpix = g * npix + pix;
  unpack8(r.z, zf);
#pragma unroll
  for (int j = 0; j < 8; ++j) dy[j] = 0.f;
  if (p.ga) {
    unpack8(r.ga, dy);
    if (p.mul_other) {
      float ao[8];
      unpack8(r.ao, ao);
#pragma unroll
      for (int j = 0; j < 8; ++j) dy[j] *= ao[j];
    }
  }
  if (p.gp) {
    const uint32_t W = p.W, H = p.H, Hp = H >> 1, Wp = W >> 1;
    const uint32_t x = pix % W, t = pix / W, y = t % H, b = t / H;
    if ((y >> 1) < Hp && (x >> 1) < Wp) {
      // nn.MaxPool2d backward routes the pooled gradient to the FIRST maximum of the 2x2 window in scan order;
      // the window is re-read by its four threads (L1 hits, no extra DRAM traffic)
      const uint32_t me = (y & 1) * 2 + (x & 1);
      const size_t w00 = (size_t)(gpix - (y & 1) * W - (x & 1)) * C8 + c8;
      const uint4 wv[4] = {__ldg(p.a + w00), __ldg(p.a + w00 + C8), __ldg(p.a + w00 + (size_t)W * C8),
                           __ldg(p.a + w00 + (size_t)(W + 1) * C8)};
      const uint4 gv = __ldg(p.gp + (size_t)(((g * p.B + b) * Hp + (y >> 1)) * Wp + (x >> 1)) * C8 + c8);
      float m[8];
      int best[8];
      unpack8(wv[0], m);
#pragma unroll
      for (int j = 0; j < 8; ++j) best[j] = 0;
#pragma unroll
      for (int d = 1; d < 4; ++d) {
        float v[8];
        unpack8(wv[d], v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (v[j] > m[j]) m[j] = v[j], best[j] = d;
      }
      float gpv[8];
      unpack8(gv, gpv);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (best[j] == (int)me) dy[j] += gpv[j];
    }
  }
  if (!p.premasked) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (!(fmaf(zf[j], sc[j], sh[j]) > 0.f)) dy[j] = 0.f;
  }
}

// pass 1: partial[blk][g][c][2] = (sum dy, sum dy * xhat)
__global__ void __launch_bounds__(256, 2) bn_bwd_reduce_kernel(BnBwd p, float* __restrict__ partial) {
  extern __shared__ float sm[];  // [blockDim][16]
  constexpr int NB = 3;          // pixels in flight per thread
  const uint32_t C8 = p.C >> 3;
  const uint32_t c8 = threadIdx.x % C8, lane_p = threadIdx.x / C8, ppb = blockDim.x / C8;
  const uint32_t npix = (uint32_t)p.B * p.H * p.W;
  const uint32_t stride = gridDim.x * ppb;
  for (uint32_t g = 0; g < (uint32_t)p.G; ++g) {
    float s1[8], s2[8], mu[8], is[8], sc[8], sh[8];
    ld8f(p.mean + g * p.C + c8 * 8, mu);
    ld8f(p.invstd + g * p.C + c8 * 8, is);
    ld8f(p.scale + g * p.C + c8 * 8, sc);
    ld8f(p.shift + g * p.C + c8 * 8, sh);
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    for (uint32_t q0 = blockIdx.x * ppb + lane_p; q0 < npix; q0 += NB * stride) {
      BnBwdRaw raw[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k)
        if (q0 + k * stride < npix) bn_bwd_load(p, g, q0 + k * stride, c8, npix, raw[k]);
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        if (q0 + k * stride >= npix) break;
        float dy[8], zf[8];
        bn_bwd_finish(p, g, q0 + k * stride, c8, npix, raw[k], sc, sh, dy, zf);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s1[j] += dy[j];
          s2[j] = fmaf(dy[j], (zf[j] - mu[j]) * is[j], s2[j]);
        }
      }
    }
    __syncthreads();
    float* mine = sm + threadIdx.x * 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) mine[j] = s1[j], mine[8 + j] = s2[j];
    __syncthreads();
    float* dst = partial + ((size_t)blockIdx.x * p.G + g) * p.C * 2;
    for (int i = threadIdx.x; i < p.C * 2; i += blockDim.x) {
      const int c = i >> 1, k = i & 1;
      float s = 0.f;
      for (uint32_t l = 0; l < ppb; ++l) s += sm[(l * C8 + (c >> 3)) * 16 + k * 8 + (c & 7)];
      dst[i] = s;
    }
  }
}

// pass 2: reduce partials (one warp per channel); dgamma, dbeta; per-group coefficients for the apply pass
// n_tile == 0: partial[blk][G][C][2] (the stand-alone reduce kernels); n_tile > 0: the conv epilogue's layout
// partial[cta][2][n_tile][2], where CTA i holds N tile i % (C / n_tile) (fabric_b200_conv3x3 with bnbwd_z).
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblk, int G, int C, double count,
                                       const float* __restrict__ gamma, const float* __restrict__ invstd,
                                       const float* __restrict__ mean, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, float* __restrict__ coef, float grad_scale, int n_tile = 0) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  double dg = 0.0, db = 0.0;
  for (int g = 0; g < G; ++g) {
    double s1 = 0.0, s2 = 0.0;
    if (n_tile > 0) {
      const int ntiles = C / n_tile, nt = c / n_tile, lc = c % n_tile;
      for (int cta = nt + lane * ntiles; cta < nblk; cta += 32 * ntiles) {
        const float2 v = *reinterpret_cast<const float2*>(partial + (((size_t)cta * 2 + g) * n_tile + lc) * 2);
        s1 += v.x;
        s2 += v.y;
      }
      // the conv epilogue accumulates the RAW second sum (sum dy * z); every lane applies the same linear map to its share
      s2 = (s2 - (double)mean[g * C + c] * s1) * (double)invstd[g * C + c];
    } else
    for (int b = lane; b < nblk; b += 32) {
      const float2 v = *reinterpret_cast<const float2*>(partial + (((size_t)b * G + g) * C + c) * 2);
      s1 += v.x;
      s2 += v.y;
    }
    for (int o = 16; o; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    dg += s2;
    db += s1;
    if (lane == 0) {
      // dz = k0 * (dy - k1 - (z - mean) * invstd * k2)  =  k0 * dy + kz * z + kc
      const float is = invstd[g * C + c], mu = mean[g * C + c];
      const float k0 = gamma[c] * is, k1 = (float)(s1 / count), k2 = (float)(s2 / count);
      coef[(g * 3 + 0) * C + c] = k0;
      coef[(g * 3 + 1) * C + c] = -k0 * is * k2;             // kz
      coef[(g * 3 + 2) * C + c] = -k0 * (k1 - mu * is * k2);  // kc
    }
  }
  if (lane == 0) {
    dgamma[c] = (float)dg * grad_scale;
    dbeta[c] = (float)db * grad_scale;
  }
}

// pass 3: dz = k0 * dy + kz * z + kc   (bf16).  A thread keeps ONE channel group (grid stride is a multiple of C/8), so
// the five per-channel coefficient vectors live in registers for the whole date group instead of being re-read from
// L1 for every 16 bytes of data (which made the first version LSU-bound at 3.9 TB/s).
__global__ void __launch_bounds__(256, 2) bn_bwd_apply_kernel(BnBwd p, const float* __restrict__ coef, uint4* __restrict__ dz) {
  constexpr int NB = 2;
  const uint32_t C8 = p.C >> 3;
  const uint32_t npix = (uint32_t)p.B * p.H * p.W;
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t c8 = gtid % C8;
  const uint32_t pstride = (gridDim.x * blockDim.x) / C8;   // host guarantees divisibility
  for (uint32_t g = 0; g < (uint32_t)p.G; ++g) {
    float sc[8], sh[8], k0[8], kz[8], kc[8];
    ld8f(p.scale + g * p.C + c8 * 8, sc);
    ld8f(p.shift + g * p.C + c8 * 8, sh);
    ld8f(coef + (g * 3 + 0) * p.C + c8 * 8, k0);
    ld8f(coef + (g * 3 + 1) * p.C + c8 * 8, kz);
    ld8f(coef + (g * 3 + 2) * p.C + c8 * 8, kc);
    for (uint32_t q0 = gtid / C8; q0 < npix; q0 += NB * pstride) {
      BnBwdRaw raw[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k)
        if (q0 + k * pstride < npix) bn_bwd_load(p, g, q0 + k * pstride, c8, npix, raw[k]);
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        const uint32_t pix = q0 + k * pstride;
        if (pix >= npix) break;
        float dy[8], zf[8], r[8];
        bn_bwd_finish(p, g, pix, c8, npix, raw[k], sc, sh, dy, zf);
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = fmaf(k0[j], dy[j], fmaf(kz[j], zf[j], kc[j]));
        dz[(size_t)(g * npix + pix) * C8 + c8] = pack8(r);
      }
    }
  }
}

// The same two passes for the PLAIN case (decoder levels and every first conv: one gradient source, no product, no pool):
// compile-time shape of the loop instead of the run-time flags of BnBwd -- four pixels in flight, no per-pixel bounds
// tests in the main loop, one basic block per iteration.  MASK: apply the ReLU mask from z (false = the producing data-gradient
// launch already masked dy).  (The generic kernels above measured 4.4-5.4 TB/s on these layers.)
template <bool MASK>
__global__ void __launch_bounds__(256, 3) bn_bwd_apply_plain_kernel(const uint4* __restrict__ z, const uint4* __restrict__ ga,
                                                                    uint32_t ga_c8, int ga_groups, const float* __restrict__ scale,
                                                                    const float* __restrict__ shift, const float* __restrict__ coef,
                                                                    uint4* __restrict__ dz, int G, uint32_t npix, int C) {
  constexpr int NB = 4;
  const uint32_t C8 = C >> 3;
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t c8 = gtid % C8;
  const uint32_t pstride = (gridDim.x * blockDim.x) / C8;   // host guarantees divisibility
  for (uint32_t g = 0; g < (uint32_t)G; ++g) {
    float sc[8], sh[8], k0[8], kz[8], kc[8];
    if (MASK) {
      ld8f(scale + g * C + c8 * 8, sc);
      ld8f(shift + g * C + c8 * 8, sh);
    }
    ld8f(coef + (g * 3 + 0) * C + c8 * 8, k0);
    ld8f(coef + (g * 3 + 1) * C + c8 * 8, kz);
    ld8f(coef + (g * 3 + 2) * C + c8 * 8, kc);
    const uint4* zg = z + (size_t)g * npix * C8 + c8;
    const uint4* gg = ga + (ga_groups == 1 ? (size_t)0 : (size_t)g * npix * ga_c8) + c8;
    uint4* og = dz + (size_t)g * npix * C8 + c8;
    auto one = [&](const uint4& zv, const uint4& gv, uint32_t pix) {
      float zf[8], dy[8], r[8];
      unpack8(zv, zf);
      unpack8(gv, dy);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (MASK && !(fmaf(zf[j], sc[j], sh[j]) > 0.f)) dy[j] = 0.f;
        r[j] = fmaf(k0[j], dy[j], fmaf(kz[j], zf[j], kc[j]));
      }
      og[(size_t)pix * C8] = pack8(r);
    };
    uint32_t q = gtid / C8;
    for (; q + (NB - 1) * pstride < npix; q += NB * pstride) {
      uint4 zv[NB], gv[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        zv[k] = __ldg(zg + (size_t)(q + k * pstride) * C8);
        gv[k] = __ldg(gg + (size_t)(q + k * pstride) * ga_c8);
      }
#pragma unroll
      for (int k = 0; k < NB; ++k) one(zv[k], gv[k], q + k * pstride);
    }
    for (; q < npix; q += pstride) one(__ldg(zg + (size_t)q * C8), __ldg(gg + (size_t)q * ga_c8), q);
  }
}

// pass 1 of the plain case: partial[blk][g][c][2] = (sum dy, sum dy * xhat), dy = relu'(z) * ga
__global__ void __launch_bounds__(256, 3) bn_bwd_reduce_plain_kernel(const uint4* __restrict__ z, const uint4* __restrict__ ga,
                                                                     uint32_t ga_c8, int ga_groups, const float* __restrict__ scale,
                                                                     const float* __restrict__ shift, const float* __restrict__ mean,
                                                                     const float* __restrict__ invstd, float* __restrict__ partial,
                                                                     int G, uint32_t npix, int C) {
  extern __shared__ float sm[];  // [blockDim][16]
  constexpr int NB = 4;
  const uint32_t C8 = C >> 3;
  const uint32_t c8 = threadIdx.x % C8, lane_p = threadIdx.x / C8, ppb = blockDim.x / C8;
  const uint32_t pstride = gridDim.x * ppb;
  for (uint32_t g = 0; g < (uint32_t)G; ++g) {
    float s1[8], s2[8], mu[8], sc[8], sh[8];
    ld8f(mean + g * C + c8 * 8, mu);
    ld8f(scale + g * C + c8 * 8, sc);
    ld8f(shift + g * C + c8 * 8, sh);
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    const uint4* zg = z + (size_t)g * npix * C8 + c8;
    const uint4* gg = ga + (ga_groups == 1 ? (size_t)0 : (size_t)g * npix * ga_c8) + c8;
    auto one = [&](const uint4& zv, const uint4& gv) {
      float zf[8], dy[8];
      unpack8(zv, zf);
      unpack8(gv, dy);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (!(fmaf(zf[j], sc[j], sh[j]) > 0.f)) dy[j] = 0.f;
        s1[j] += dy[j];
        s2[j] = fmaf(dy[j], zf[j] - mu[j], s2[j]);   // x invstd once, below
      }
    };
    uint32_t q = blockIdx.x * ppb + lane_p;
    for (; q + (NB - 1) * pstride < npix; q += NB * pstride) {
      uint4 zv[NB], gv[NB];
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        zv[k] = __ldg(zg + (size_t)(q + k * pstride) * C8);
        gv[k] = __ldg(gg + (size_t)(q + k * pstride) * ga_c8);
      }
#pragma unroll
      for (int k = 0; k < NB; ++k) one(zv[k], gv[k]);
    }
    for (; q < npix; q += pstride) one(__ldg(zg + (size_t)q * C8), __ldg(gg + (size_t)q * ga_c8));
    float is[8];
    ld8f(invstd + g * C + c8 * 8, is);
    __syncthreads();
    float* mine = sm + threadIdx.x * 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) mine[j] = s1[j], mine[8 + j] = s2[j] * is[j];
    __syncthreads();
    float* dst = partial + ((size_t)blockIdx.x * G + g) * C * 2;
    for (int i = threadIdx.x; i < C * 2; i += blockDim.x) {
      const int c = i >> 1, k = i & 1;
      float s = 0.f;
      for (uint32_t l = 0; l < ppb; ++l) s += sm[(l * C8 + (c >> 3)) * 16 + k * 8 + (c & 7)];
      dst[i] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------- both dates at once
// The encoder's second convs carry the product fusion (bidate_model.py:35-38): date g's gradient needs the OTHER date's
// activation and both dates share one upstream gradient,
//     dy_g = 1[a_g > 0] * ( ga * a_{1-g} + unpool(gp_g) ).
// Walking the two date groups one after the other (kernels above) reads a0, a1 and ga twice per pass (8.5 tensor units of
// HBM traffic); here one thread handles 8 channels of one pixel for BOTH dates, so z0, z1, a0, a1, ga, gp cross HBM once
// per pass (5.25 units).  a_g > 0 is the ReLU mask (a = relu(bn(z)) as stored), so scale / shift are not needed.
// RECOMP: the activation tensor is not read (nor stored by the forward pass at all): a = bf16(relu(z * scale + shift)) is
// recomputed from z exactly as bn_apply_kernel computed and rounded it, for the pixel, for the other date (product adjoint)
// and for the 2x2 pooling window (arg-max routing).  Per pass z0, z1, ga, gp cross HBM: 3.5 tensor units instead of 5.25.
template <bool RECOMP>
__device__ __forceinline__ void bn_act8(const uint4& zv, const float (&sc)[8], const float (&sh)[8], float (&af)[8]) {
  float zf[8];
  unpack8(zv, zf);
#pragma unroll
  for (int j = 0; j < 8; ++j) zf[j] = fmaxf(fmaf(zf[j], sc[j], sh[j]), 0.f);
  unpack8(pack8(zf), af);   // the value as the forward pass stored it (bf16, round to nearest even)
}

template <bool GP, bool RECOMP>
__device__ __forceinline__ void bn_bwd2_dy(const BnBwd& p, uint32_t pix, uint32_t c8, uint32_t npix, float (&dy)[2][8],
                                           float (&zf)[2][8], const float (&sc)[2][8], const float (&sh)[2][8]) {
  const uint32_t C8 = p.C >> 3;
  const uint4* asrc = RECOMP ? p.z : p.a;   // where the pooling window is read from
  uint4 zr[2], ar[2], gpr[2], wv[2][4];
  const uint4 gar = __ldg(p.ga + (size_t)pix * p.ga_c8 + c8);
  bool pool_ok = false;
  uint32_t me = 0;
  if (GP) {
    const uint32_t W = p.W, H = p.H, Hp = H >> 1, Wp = W >> 1;
    const uint32_t x = pix % W, t = pix / W, y = t % H, b = t / H;
    pool_ok = (y >> 1) < Hp && (x >> 1) < Wp;
    me = (y & 1) * 2 + (x & 1);
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      if (pool_ok) {
        const size_t w00 = (size_t)(g * npix + pix - (y & 1) * W - (x & 1)) * C8 + c8;
        wv[g][0] = __ldg(asrc + w00), wv[g][1] = __ldg(asrc + w00 + C8), wv[g][2] = __ldg(asrc + w00 + (size_t)W * C8),
        wv[g][3] = __ldg(asrc + w00 + (size_t)(W + 1) * C8);
        gpr[g] = __ldg(p.gp + (size_t)(((g * p.B + b) * Hp + (y >> 1)) * Wp + (x >> 1)) * C8 + c8);
      }
    }
  }
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    zr[g] = __ldg(p.z + (size_t)(g * npix + pix) * C8 + c8);
    if (!RECOMP) ar[g] = __ldg(p.a + (size_t)(g * npix + pix) * C8 + c8);
  }
  float gaf[8], af[2][8];
  unpack8(gar, gaf);
  if (RECOMP) {
    bn_act8<true>(zr[0], sc[0], sh[0], af[0]);
    bn_act8<true>(zr[1], sc[1], sh[1], af[1]);
  } else {
    unpack8(ar[0], af[0]);
    unpack8(ar[1], af[1]);
  }
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    unpack8(zr[g], zf[g]);
#pragma unroll
    for (int j = 0; j < 8; ++j) dy[g][j] = gaf[j] * af[1 - g][j];
    if (GP && pool_ok) {
      // nn.MaxPool2d backward: the pooled gradient goes to the FIRST maximum of the 2x2 window in scan order
      float m[8], gpv[8];
      int best[8];
      if (RECOMP) bn_act8<true>(wv[g][0], sc[g], sh[g], m);
      else unpack8(wv[g][0], m);
#pragma unroll
      for (int j = 0; j < 8; ++j) best[j] = 0;
#pragma unroll
      for (int d = 1; d < 4; ++d) {
        float v[8];
        if (RECOMP) bn_act8<true>(wv[g][d], sc[g], sh[g], v);
        else unpack8(wv[g][d], v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (v[j] > m[j]) m[j] = v[j], best[j] = d;
      }
      unpack8(gpr[g], gpv);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (best[j] == (int)me) dy[g][j] += gpv[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (!(af[g][j] > 0.f)) dy[g][j] = 0.f;
  }
}

// (RECOMP keeps 32 more per-channel coefficients in registers: 128-thread blocks, three per SM -> 170 registers per thread)
template <bool GP, bool RECOMP>
__global__ void __launch_bounds__(RECOMP ? 128 : 256, RECOMP ? 3 : 2) bn_bwd2_reduce_kernel(BnBwd p, float* __restrict__ partial) {
  extern __shared__ float sm[];  // [blockDim][16]
  const uint32_t C8 = p.C >> 3;
  const uint32_t c8 = threadIdx.x % C8, lane_p = threadIdx.x / C8, ppb = blockDim.x / C8;
  const uint32_t npix = (uint32_t)p.B * p.H * p.W;
  const uint32_t stride = gridDim.x * ppb;
  float s1[2][8], s2[2][8], mu[2][8], sc[2][8], sh[2][8];
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    ld8f(p.mean + g * p.C + c8 * 8, mu[g]);
    ld8f(p.scale + g * p.C + c8 * 8, sc[g]);
    ld8f(p.shift + g * p.C + c8 * 8, sh[g]);
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[g][j] = s2[g][j] = 0.f;
  }
  for (uint32_t q = blockIdx.x * ppb + lane_p; q < npix; q += stride) {
    float dy[2][8], zf[2][8];
    bn_bwd2_dy<GP, RECOMP>(p, q, c8, npix, dy, zf, sc, sh);
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[g][j] += dy[g][j];
        s2[g][j] = fmaf(dy[g][j], zf[g][j] - mu[g][j], s2[g][j]);   // x invstd once, below
      }
  }
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    float is[8];
    ld8f(p.invstd + g * p.C + c8 * 8, is);
    __syncthreads();
    float* mine = sm + threadIdx.x * 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) mine[j] = s1[g][j], mine[8 + j] = s2[g][j] * is[j];
    __syncthreads();
    float* dst = partial + ((size_t)blockIdx.x * p.G + g) * p.C * 2;
    for (int i = threadIdx.x; i < p.C * 2; i += blockDim.x) {
      const int c = i >> 1, k = i & 1;
      float s = 0.f;
      for (uint32_t l = 0; l < ppb; ++l) s += sm[(l * C8 + (c >> 3)) * 16 + k * 8 + (c & 7)];
      dst[i] = s;
    }
  }
}

template <bool GP, bool RECOMP>
__global__ void __launch_bounds__(RECOMP ? 128 : 256, RECOMP ? 3 : 2) bn_bwd2_apply_kernel(BnBwd p, const float* __restrict__ coef, uint4* __restrict__ dz) {
  const uint32_t C8 = p.C >> 3;
  const uint32_t npix = (uint32_t)p.B * p.H * p.W;
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t c8 = gtid % C8;
  const uint32_t pstride = (gridDim.x * blockDim.x) / C8;   // host guarantees divisibility
  float k0[2][8], kz[2][8], kc[2][8], sc[2][8], sh[2][8];
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    ld8f(coef + (g * 3 + 0) * p.C + c8 * 8, k0[g]);
    ld8f(coef + (g * 3 + 1) * p.C + c8 * 8, kz[g]);
    ld8f(coef + (g * 3 + 2) * p.C + c8 * 8, kc[g]);
    ld8f(p.scale + g * p.C + c8 * 8, sc[g]);
    ld8f(p.shift + g * p.C + c8 * 8, sh[g]);
  }
  for (uint32_t q = gtid / C8; q < npix; q += pstride) {
    float dy[2][8], zf[2][8];
    bn_bwd2_dy<GP, RECOMP>(p, q, c8, npix, dy, zf, sc, sh);
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      float r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = fmaf(k0[g][j], dy[g][j], fmaf(kz[g][j], zf[g][j], kc[g][j]));
      dz[(size_t)(g * npix + q) * C8 + c8] = pack8(r);
    }
  }
}


// ---------------------------------------------------------------------------------------------- quad form (recompute)
// Same arithmetic as bn_bwd2_* with the activation recomputed from z, but ONE THREAD = 8 channels of one 2x2 pixel QUAD, for
// BOTH dates: the pooling window is the thread's own four pixels, so z0, z1, ga and gp cross HBM exactly once per pass (3.5
// tensor units, against 5.25 for the kernels reading the stored activation) with 8 + 4 + 2 sixteen-byte loads in flight per
// thread, and the activation is recomputed 8 times per quad.  (A per-pixel recompute form -- every thread re-reading and
// re-computing its whole window -- ran 40 % SLOWER than reading the stored activation: issue-bound.  A per-date quad form
// read z twice.)
// V = channels per thread: 8 (16-byte loads) or 4 (8-byte loads).  The 8-channel form needs ~220 registers (2 blocks of 128
// threads per SM) and ncu showed it issue-latency-bound: 2 warps per scheduler, 3.6 cycles per issued instruction per warp,
// 54 % issue utilisation at 2.9 TB/s (reduce) / 4.2 TB/s (apply).  The 4-channel form halves every per-thread array
// (<= 128 registers, 4 blocks per SM): twice the warps to hide the same latencies, same bytes per warp-wide access pattern
// (8 B x 32 lanes = two 128-byte pixels of a 64-channel tensor per load instruction).
template <int V>
struct QVec;
template <>
struct QVec<8> {
  using type = uint4;
};
template <>
struct QVec<4> {
  using type = uint2;
};
__device__ __forceinline__ void unpackv(const uint4& v, float (&f)[8]) { unpack8(v, f); }
__device__ __forceinline__ void unpackv(const uint2& v, float (&f)[4]) {
  f[0] = fb::bf16_lo(v.x), f[1] = fb::bf16_hi(v.x), f[2] = fb::bf16_lo(v.y), f[3] = fb::bf16_hi(v.y);
}
__device__ __forceinline__ uint4 packv(const float (&f)[8]) { return pack8(f); }
__device__ __forceinline__ uint2 packv(const float (&f)[4]) {
  return make_uint2(fb::pack_bf16x2(f[0], f[1]), fb::pack_bf16x2(f[2], f[3]));
}
__device__ __forceinline__ void ldvf(const float* __restrict__ p, float (&v)[8]) { ld8f(p, v); }
__device__ __forceinline__ void ldvf(const float* __restrict__ p, float (&v)[4]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w;
}
// the activation as the forward pass stored it: bf16(relu(z * scale + shift)), as floats
template <int V, typename T>
__device__ __forceinline__ void bn_actv(const T& zv, const float (&sc)[V], const float (&sh)[V], float (&af)[V]) {
  float zf[V];
  unpackv(zv, zf);
#pragma unroll
  for (int j = 0; j < V; ++j) zf[j] = fmaxf(fmaf(zf[j], sc[j], sh[j]), 0.f);
  unpackv(packv(zf), af);
}

// EVEN: H and W are even (every real level), so every quad is whole and -- with GP -- has a pooled gradient: the per-pixel
// bounds branches disappear, which is worth more than their own instructions: the quad becomes ONE basic block in which the
// compiler shares the unpacked z / upstream-gradient values between the two dates and interleaves the four pixels.
// ... and packed: max(x, 0) rides in the conversion (cvt.rn.relu.bf16x2: negative values and -0 become +0, exactly what
// fmaxf(x, 0) followed by a round-to-nearest conversion stores)
template <int V>
__device__ __forceinline__ typename QVec<V>::type bn_act_packed(const typename QVec<V>::type& zv, const float (&sc)[V],
                                                               const float (&sh)[V]) {
  float zf[V];
  unpackv(zv, zf);
#pragma unroll
  for (int j = 0; j < V; ++j) zf[j] = fmaf(zf[j], sc[j], sh[j]);
  if constexpr (V == 8) {
    return make_uint4(fb::pack_bf16x2_relu(zf[0], zf[1]), fb::pack_bf16x2_relu(zf[2], zf[3]), fb::pack_bf16x2_relu(zf[4], zf[5]),
                      fb::pack_bf16x2_relu(zf[6], zf[7]));
  } else {
    return make_uint2(fb::pack_bf16x2_relu(zf[0], zf[1]), fb::pack_bf16x2_relu(zf[2], zf[3]));
  }
}

template <bool GP, bool APPLY, int V, bool EVEN = false, bool PIPE = false>
__global__ void __launch_bounds__(128, V == 8 ? 2 : (PIPE ? 2 : 4)) bn_bwd2q_kernel(BnBwd p, const float* __restrict__ coef, uint4* __restrict__ dz_,
                                                                     float* __restrict__ partial) {
  using T = typename QVec<V>::type;
  extern __shared__ float sm[];  // reduce: [blockDim][2 * V]
  const T* __restrict__ Z = reinterpret_cast<const T*>(p.z);
  const T* __restrict__ GA = reinterpret_cast<const T*>(p.ga);
  const T* __restrict__ GPP = reinterpret_cast<const T*>(p.gp);
  T* __restrict__ dz = reinterpret_cast<T*>(dz_);
  const uint32_t CV = p.C / V, ga_cv = p.ga_c8 * (8 / V);
  const uint32_t H = p.H, W = p.W, Hq = (H + 1) >> 1, Wq = (W + 1) >> 1, Hp = H >> 1, Wp = W >> 1;
  const uint32_t npix = (uint32_t)p.B * H * W, nquad = (uint32_t)p.B * Hq * Wq;
  const uint32_t cv = threadIdx.x % CV, lane_p = threadIdx.x / CV, ppb = blockDim.x / CV;
  const uint32_t stride = gridDim.x * ppb;
  float sc[2][V], sh[2][V];
  float ka[2][V], kb[2][V], kc[2][V];  // reduce: ka = mean, sums s1 / s2 in kb / kc;  apply: k0, kz, kc
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    ldvf(p.scale + g * p.C + cv * V, sc[g]);
    ldvf(p.shift + g * p.C + cv * V, sh[g]);
    if (APPLY) {
      ldvf(coef + (g * 3 + 0) * p.C + cv * V, ka[g]);
      ldvf(coef + (g * 3 + 1) * p.C + cv * V, kb[g]);
      ldvf(coef + (g * 3 + 2) * p.C + cv * V, kc[g]);
    } else {
      ldvf(p.mean + g * p.C + cv * V, ka[g]);
#pragma unroll
      for (int j = 0; j < V; ++j) kb[g][j] = kc[g][j] = 0.f;
    }
  }
  // one quad: its 14 loads ...
  auto load_quad = [&](uint32_t q, T (&zv)[2][4], T (&gq)[4], T (&gpv)[2], uint32_t (&pix)[4], bool (&ok)[4], bool& pool_ok) {
    const uint32_t qx = q % Wq, t = q / Wq, qy = t % Hq, b = t / Hq;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const uint32_t y = 2 * qy + (d >> 1), x = 2 * qx + (d & 1);
      ok[d] = EVEN || (y < H && x < W);
      pix[d] = (b * H + y) * W + x;
      if (ok[d]) {
        zv[0][d] = __ldg(Z + (size_t)pix[d] * CV + cv);
        zv[1][d] = __ldg(Z + (size_t)(npix + pix[d]) * CV + cv);
        gq[d] = __ldg(GA + (size_t)pix[d] * ga_cv + cv);
      }
    }
    pool_ok = GP && (EVEN || (qy < Hp && qx < Wp));   // (then the whole window exists)
    if (pool_ok) {
      gpv[0] = __ldg(GPP + (size_t)((b * Hp + qy) * Wp + qx) * CV + cv);
      gpv[1] = __ldg(GPP + (size_t)(((p.B + b) * Hp + qy) * Wp + qx) * CV + cv);
    }
  };
  // ... and its arithmetic
  auto do_quad = [&](const T (&zv)[2][4], const T (&gq)[4], const T (&gpv)[2], const uint32_t (&pix)[4], const bool (&ok)[4],
                     const bool pool_ok) {
    // activations of both dates as the forward pass stored them (bf16), kept packed
    T av[2][4];
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
      for (int d = 0; d < 4; ++d)
        if (ok[d]) av[g][d] = bn_act_packed<V>(zv[g][d], sc[g], sh[g]);
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      float A[4][V], v[4][V];   // own activation; upstream gradient through the product = ga * (other date's activation)
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        if (!ok[d]) continue;
        float at[V], gaf[V];
        unpackv(av[g][d], A[d]);
        unpackv(av[1 - g][d], at);
        unpackv(gq[d], gaf);
#pragma unroll
        for (int j = 0; j < V; ++j) v[d][j] = gaf[j] * at[j];
      }
      if (pool_ok) {
        // nn.MaxPool2d backward: the pooled gradient goes to the FIRST maximum of the window in scan order
        float gpf[V];
        unpackv(gpv[g], gpf);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const float m = fmaxf(fmaxf(A[0][j], A[1][j]), fmaxf(A[2][j], A[3][j]));
          const bool e0 = A[0][j] == m, e1 = A[1][j] == m, e2 = A[2][j] == m;   // (bitwise: no short-circuit branches)
          const bool p0 = e0, p1 = !e0 & e1, p2 = !e0 & !e1 & e2;
          v[0][j] += p0 ? gpf[j] : 0.f;           // (selects, not branches: the four lanes of a warp disagree)
          v[1][j] += p1 ? gpf[j] : 0.f;
          v[2][j] += p2 ? gpf[j] : 0.f;
          v[3][j] += (e0 | e1 | e2) ? 0.f : gpf[j];
        }
      }
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        if (!ok[d]) continue;
        float zf[V], dy[V];
        unpackv(zv[g][d], zf);
#pragma unroll
        for (int j = 0; j < V; ++j) dy[j] = A[d][j] > 0.f ? v[d][j] : 0.f;
        if (APPLY) {
          float r[V];
#pragma unroll
          for (int j = 0; j < V; ++j) r[j] = fmaf(ka[g][j], dy[j], fmaf(kb[g][j], zf[j], kc[g][j]));
          dz[(size_t)(g * npix + pix[d]) * CV + cv] = packv(r);
        } else {
#pragma unroll
          for (int j = 0; j < V; ++j) {
            kb[g][j] += dy[j];
            kc[g][j] = fmaf(dy[j], zf[j] - ka[g][j], kc[g][j]);   // x invstd once, below
          }
        }
      }
    }
  };
  if constexpr (PIPE) {
    // software pipeline over two register sets: the loads of quad i+1 are in flight while quad i is computed (the plain loop
    // issued them only after the arithmetic of quad i: every warp alternated between waiting and computing)
    T zvA[2][4], gqA[4], gpA[2], zvB[2][4], gqB[4], gpB[2];
    uint32_t pixA[4], pixB[4];
    bool okA[4], okB[4], poA = false, poB = false;
    uint32_t q = blockIdx.x * ppb + lane_p;
    if (q < nquad) load_quad(q, zvA, gqA, gpA, pixA, okA, poA);
    while (q < nquad) {
      if (q + stride < nquad) load_quad(q + stride, zvB, gqB, gpB, pixB, okB, poB);
      do_quad(zvA, gqA, gpA, pixA, okA, poA);
      q += stride;
      if (q >= nquad) break;
      if (q + stride < nquad) load_quad(q + stride, zvA, gqA, gpA, pixA, okA, poA);
      do_quad(zvB, gqB, gpB, pixB, okB, poB);
      q += stride;
    }
  } else {
    for (uint32_t q = blockIdx.x * ppb + lane_p; q < nquad; q += stride) {
      T zv[2][4], gq[4], gpv[2];
      bool ok[4], pool_ok;
      uint32_t pix[4];
      load_quad(q, zv, gq, gpv, pix, ok, pool_ok);
      do_quad(zv, gq, gpv, pix, ok, pool_ok);
    }
  }
  if (!APPLY) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      float is[V];
      ldvf(p.invstd + g * p.C + cv * V, is);
      __syncthreads();
      float* mine = sm + threadIdx.x * (2 * V);
#pragma unroll
      for (int j = 0; j < V; ++j) mine[j] = kb[g][j], mine[V + j] = kc[g][j] * is[j];
      __syncthreads();
      float* dst = partial + ((size_t)blockIdx.x * p.G + g) * p.C * 2;
      for (int i = threadIdx.x; i < p.C * 2; i += blockDim.x) {
        const int c = i >> 1, k = i & 1;
        float s_ = 0.f;
        for (uint32_t l = 0; l < ppb; ++l) s_ += sm[(l * CV + (c / V)) * (2 * V) + k * V + (c % V)];
        dst[i] = s_;
      }
    }
  }
}

// ================================================================================================ decoder-input adjoint
// dlow[b][i][j][c] = sum_{u,v} wy(u,i) wx(v,j) dcat[b][u+padT][v+padL][Cs+c]   (adjoint of bilinear x2 + pad)
__global__ void __launch_bounds__(256, 4)
up_input_bwd_kernel(const uint4* __restrict__ dcat, uint4* __restrict__ dlow, int B, int H, int W, int Cs, int h, int w, int Cl) {
  const uint32_t Ct8 = (Cs + Cl) >> 3, Cs8 = Cs >> 3, Cl8 = Cl >> 3;
  const int padT = (H - 2 * h) / 2, padL = (W - 2 * w) / 2;
  const float sy = (2 * h > 1) ? (float)(h - 1) / (float)(2 * h - 1) : 0.f;
  const float sx = (2 * w > 1) ? (float)(w - 1) / (float)(2 * w - 1) : 0.f;
  const uint32_t total = (uint32_t)B * h * w * Cl8;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const uint32_t c8 = idx % Cl8;
    uint32_t r = idx / Cl8;
    const int j = r % w;
    r /= w;
    const int i = r % h;
    const uint32_t b = r / h;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    // output rows u whose two taps (y0, y1) include i lie within [2i-3, 2i+3]
    for (int u = max(0, 2 * i - 3); u <= min(2 * h - 1, 2 * i + 3); ++u) {
      const float fy = sy * u;
      const int y0 = (int)fy, y1 = min(y0 + 1, h - 1);
      const float ly = fy - y0;
      const float wy = (y0 == i ? 1.f - ly : 0.f) + (y1 == i ? ly : 0.f);
      if (wy == 0.f) continue;
      for (int v = max(0, 2 * j - 3); v <= min(2 * w - 1, 2 * j + 3); ++v) {
        const float fx = sx * v;
        const int x0 = (int)fx, x1 = min(x0 + 1, w - 1);
        const float lx = fx - x0;
        const float wx = (x0 == j ? 1.f - lx : 0.f) + (x1 == j ? lx : 0.f);
        if (wx == 0.f) continue;
        float f[8];
        unpack8(__ldg(dcat + (size_t)((b * H + u + padT) * W + v + padL) * Ct8 + Cs8 + c8), f);
        const float ww = wy * wx;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(ww, f[k], acc[k]);
      }
    }
    dlow[idx] = pack8(acc);
  }
}

// Same adjoint, gather form with the tap search hoisted: grid = (items of one low-res row / 256, h, B), one thread = 8
// channels of one low-res pixel (i, j).  Output row u touches i only for u in [2i-2, 2i+3] (align_corners scale
// (h-1)/(2h-1) in [1/3, 1/2)), likewise for columns: the six row and six column weights are computed once and the
// 6 x 6 candidates are skipped where either weight is zero (~16 survive).  The scanning kernel above spent ~1000
// instructions per 16 output bytes (issue-bound at 15 % of HBM bandwidth).
__global__ void __launch_bounds__(256) up_input_bwd_rows_kernel(const uint4* __restrict__ dcat, uint4* __restrict__ dlow, int H,
                                                                int W, int Cs8, int h, int w, int Cl8, int cl8_shift,
                                                                int padT, int padL, float sy, float sx) {
  const uint32_t item = blockIdx.x * 256u + threadIdx.x;
  const uint32_t j = cl8_shift >= 0 ? item >> cl8_shift : item / (uint32_t)Cl8;
  if (j >= (uint32_t)w) return;
  const uint32_t c8 = item - j * Cl8;
  const int i = blockIdx.y;
  const uint32_t b = blockIdx.z;
  float wy[6], wx[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int u = 2 * i - 2 + k, v = 2 * (int)j - 2 + k;
    wy[k] = 0.f, wx[k] = 0.f;
    if (u >= 0 && u < 2 * h) {
      const float fy = sy * u;
      const int y0 = (int)fy, y1 = min(y0 + 1, h - 1);
      const float ly = fy - y0;
      wy[k] = (y0 == i ? 1.f - ly : 0.f) + (y1 == i ? ly : 0.f);
    }
    if (v >= 0 && v < 2 * w) {
      const float fx = sx * v;
      const int x0 = (int)fx, x1 = min(x0 + 1, w - 1);
      const float lx = fx - x0;
      wx[k] = (x0 == (int)j ? 1.f - lx : 0.f) + (x1 == (int)j ? lx : 0.f);
    }
  }
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  const uint32_t Ct8 = Cs8 + Cl8;
#pragma unroll
  for (int a = 0; a < 6; ++a) {
    if (wy[a] == 0.f) continue;
    const uint32_t rowbase = ((b * H + (uint32_t)(2 * i - 2 + a + padT)) * W + (uint32_t)(2 * (int)j - 2 + padL)) * Ct8 + Cs8 + c8;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      if (wx[c] == 0.f) continue;
      float f[8];
      unpack8(__ldg(dcat + (uint32_t)(rowbase + (uint32_t)c * Ct8)), f);   // 32-bit wrap-around: the sum is in range
      const float ww = wy[a] * wx[c];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fmaf(ww, f[k], acc[k]);
    }
  }
  dlow[((b * h + i) * w + j) * (uint32_t)Cl8 + c8] = pack8(acc);
}

// fp32 [S][Cout][9][CinPad] split-K partials -> nn.Conv2d weight gradient [Cout][Cin][3][3]
// One thread = four consecutive input channels of one (co, tap): the S partial slabs are read as float4 in their own order
// (coalesced; the first version walked the RESULT order and read 4 bytes per 128-byte line), the 36-byte-strided writes of the
// small result are absorbed by L2.
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, int S, int Cout, int Cin, int CinPad, float* __restrict__ dw) {
  const size_t slab = (size_t)Cout * 9 * CinPad;
  const size_t n4 = slab >> 2;   // CinPad is a multiple of 16
  const int cp4 = CinPad >> 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cp4) * 4;
    if (ci >= Cin) continue;      // channel padding (13 -> 16)
    const size_t row = i / cp4;   // co * 9 + tap
    const int tap = (int)(row % 9);
    const size_t co = row / 9;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < S; ++k) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(ws + k * slab) + i);
      s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
    }
    float* dst = dw + (co * Cin + ci) * 9 + tap;
    dst[0] = s.x;
    if (ci + 1 < Cin) dst[9] = s.y;
    if (ci + 2 < Cin) dst[18] = s.z;
    if (ci + 3 < Cin) dst[27] = s.w;
  }
}

// The same for the operand-swapped launch (64-channel dL/dz against a >= 128-channel conv input: the INPUT's channels are the
// M dimension, so all 128 MMA rows are useful instead of 64 x two filter rows = 75 %): ws holds
//   dW'[ci][tap'][co] = sum_px X[px][ci] * dZ[px + tap'][co]  =  dW[co][8 - tap'][ci]        [S][Cin][9][Cout] fp32
__global__ void wgrad_reduce_swapped_kernel(const float* __restrict__ ws, int S, int Cout, int Cin, float* __restrict__ dw) {
  const size_t n = (size_t)Cout * Cin * 9;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    // i walks ws' (coalesced reads); the 36-byte-strided writes of the small result are absorbed by L2
    const int co = i % Cout, tp = (i / Cout) % 9, ci = i / ((size_t)9 * Cout);
    float s = 0.f;
    for (int k = 0; k < S; ++k) s += ws[k * n + i];
    dw[((size_t)co * Cin + ci) * 9 + (8 - tp)] = s;
  }
}

}  // namespace

// ====================================================================================================== exports
extern "C" {

int fabric_b200_bn_finalize(const float* stats_ws, int grid, int n_tile, int C, int G, int64_t count_per_group,
                            const float* conv_bias, const float* gamma, const float* beta, float* running_mean,
                            float* running_var, int64_t* num_batches_tracked, float momentum, float eps, float* scale,
                            float* shift, float* mean, float* invstd, void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!stats_ws || !gamma || !beta || !running_mean || !running_var || !scale || !shift || !mean || !invstd)
    return fail(FB_ERR_ARG, "null pointer");
  if (C % n_tile || G < 1 || G > 2 || grid < 1 || count_per_group < 1) return fail(FB_ERR_SHAPE, "bad shape");
  bn_finalize_kernel<<<(C + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
      stats_ws, grid, n_tile, C, G, (double)count_per_group, conv_bias, gamma, beta, running_mean, running_var,
      reinterpret_cast<long long*>(num_batches_tracked), momentum, eps, scale, shift, mean, invstd);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int fabric_b200_bn_apply_relu(const void* z, const float* scale, const float* shift, void* a, void* pool_out, void* prod_out,
                              int prod_channels, int G, int B, int H, int W, int C, void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!z || !scale || !shift) return fail(FB_ERR_ARG, "null pointer");
  if (!a && !(pool_out && prod_out)) return fail(FB_ERR_ARG, "the activation may be skipped only when both its pooled copy and the date product are written");
  if (C % 8) return fail(FB_ERR_SHAPE, "C must be a multiple of 8");
  if ((double)G * B * H * W * C / 8 >= 4.0e9) return fail(FB_ERR_SHAPE, "tensor too large for 32-bit indexing");
  if (prod_out && (G != 2 || prod_channels < C || prod_channels % 8)) return fail(FB_ERR_SHAPE, "product fusion needs G == 2 and prod_channels >= C");
  const size_t n = (size_t)(prod_out ? 1 : G) * B * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
  auto kern = (H % 2 == 0 && W % 2 == 0) ? bn_apply_kernel<true> : bn_apply_kernel<false>;
  kern<<<ew_grid(n, 256, di.sms), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(z), scale, shift, reinterpret_cast<uint4*>(a), reinterpret_cast<uint4*>(pool_out),
      reinterpret_cast<uint4*>(prod_out), prod_channels / 8, G, B, H, W, C);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int64_t fabric_b200_seg_loss_ws_floats(int B, int H, int W) {
  const int rows = B * H;
  const int rpb = (rows + 295) / 296;
  const int nblk = (rows + rpb - 1) / rpb;
  return (int64_t)nblk * 6 * W + 10 * (int64_t)W + 1024;
}

// phase bits: 1 = per-column sums [6][W] (I_c, P_c, T_c) into the workspace tail; 2 = loss + dL/dlogits from those sums.
// Between the two a data-parallel caller may all-reduce the sums (exact-global loss, reference train.py:91-92 computes the
// loss on the gathered batch).  Focal / CE are plain means: `mean_scale` (1/world) scales their loss and gradient instead.
static int seg_loss_phases(int phase, int kind, float alpha, float beta, float gamma, float eps, const float* logits,
                           const int64_t* labels, int label_ndim, int B, int H, int W, float* loss_out, float* dlogits,
                           float* ws, float mean_scale, void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!logits || !labels || !loss_out || !dlogits || !ws) return fail(FB_ERR_ARG, "null pointer");
  if (label_ndim != 3 && label_ndim != 4) return fail(FB_ERR_SHAPE, "labels must be [B,H,W] or [B,1,H,W]");
  if (W > 8192) return fail(FB_ERR_SHAPE, "W too large");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)B * H * W;
  if (kind >= 0 && kind <= 2) {
    const int rows = B * H;
    const int rpb = (rows + 295) / 296;
    const int nblk = (rows + rpb - 1) / rpb;
    float* partial = ws;
    float* coef = ws + (size_t)nblk * 6 * W;
    float* sums = coef + 4 * (size_t)W;  // [6][W], inside the workspace tail
    if (phase & 1) {
      seg_loss_partial_kernel<<<nblk, 256, 0, st>>>(logits, reinterpret_cast<const long long*>(labels), B, H, W, rpb, partial);
      FB_CUDA(cudaGetLastError());
      reduce_partials_kernel<<<(6 * W + 255) / 256, 256, 0, st>>>(partial, nblk, 6 * W, sums);
      FB_CUDA(cudaGetLastError());
    }
    if (phase & 2) {
      const size_t smem = 6 * (size_t)W * sizeof(float);
      FB_CUDA(cudaFuncSetAttribute(seg_loss_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      seg_loss_finalize_kernel<<<1, 256, smem, st>>>(sums, 1, W, label_ndim, kind, alpha, beta, eps, coef, loss_out);
      FB_CUDA(cudaGetLastError());
      seg_loss_grad_kernel<<<ew_grid(n, 256, di.sms), 256, 0, st>>>(logits, reinterpret_cast<const long long*>(labels), coef, B,
                                                                  H, W, dlogits);
      FB_CUDA(cudaGetLastError());
    }
  } else if (kind == 3 || kind == 4) {  // focal / cross entropy: one pass, nothing to exchange (phase 1 is empty)
    if (phase & 2) {
      const int nblk = 296;
      focal_loss_kernel<<<nblk, 256, 0, st>>>(logits, reinterpret_cast<const long long*>(labels), B, H, W,
                                              kind == 4 ? 0.f : gamma, dlogits, ws, mean_scale);
      FB_CUDA(cudaGetLastError());
      reduce_partials_kernel<<<1, 32, 0, st>>>(ws, nblk, 1, loss_out);
      FB_CUDA(cudaGetLastError());
    }
  } else {
    return fail(FB_ERR_ARG, "unknown loss kind %d", kind);
  }
  return FB_OK;
}

int fabric_b200_seg_loss_fwd_bwd(int kind, float alpha, float beta, float gamma, float eps, const float* logits,
                                 const int64_t* labels, int label_ndim, int B, int H, int W, float* loss_out, float* dlogits,
                                 float* ws, void* stream) {
  return seg_loss_phases(3, kind, alpha, beta, gamma, eps, logits, labels, label_ndim, B, H, W, loss_out, dlogits, ws, 1.f, stream);
}

int fabric_b200_seg_loss_phase(int phase, int kind, float alpha, float beta, float gamma, float eps, const float* logits,
                               const int64_t* labels, int label_ndim, int B, int H, int W, float* loss_out, float* dlogits,
                               float* ws, float mean_scale, void* stream) {
  if (phase != 1 && phase != 2) return fail(FB_ERR_ARG, "phase must be 1 (sums) or 2 (loss + gradient)");
  return seg_loss_phases(phase, kind, alpha, beta, gamma, eps, logits, labels, label_ndim, B, H, W, loss_out, dlogits, ws,
                         mean_scale, stream);
}

int64_t fabric_b200_seg_loss_sums_offset(int B, int H, int W) {
  const int rows = B * H;
  const int rpb = (rows + 295) / 296;
  const int nblk = (rows + rpb - 1) / rpb;
  return (int64_t)nblk * 6 * W + 4 * (int64_t)W;
}

int fabric_b200_outconv_bwd(const float* dlogits, const void* u, const float* w, void* du, float* dw, float* db, float* ws,
                            int B, int H, int W, int C, void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!dlogits || !u || !w || !du || !dw || !db || !ws) return fail(FB_ERR_ARG, "null pointer");
  if (C % 8 || 256 % (C / 8)) return fail(FB_ERR_SHAPE, "bad C");
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = di.sms * 4;
  outconv_bwd_kernel<<<nblk, 256, 256 * 18 * sizeof(float), st>>>(dlogits, reinterpret_cast<const uint4*>(u), w,
                                                                   reinterpret_cast<uint4*>(du), ws, B, H, W, C);
  FB_CUDA(cudaGetLastError());
  // ws rows are [2C dW | 2 db]; reduce into a contiguous temp at the end of ws, then split
  float* red = ws + (size_t)nblk * (2 * C + 2);
  reduce_partials_kernel<<<1, 256, 0, st>>>(ws, nblk, 2 * C + 2, red);
  FB_CUDA(cudaGetLastError());
  FB_CUDA(cudaMemcpyAsync(dw, red, 2 * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
  FB_CUDA(cudaMemcpyAsync(db, red + 2 * C, 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return FB_OK;
}

int64_t fabric_b200_outconv_bwd_ws_floats(int C) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  return (int64_t)(di.sms * 4 + 1) * (2 * C + 2);
}

// reduce-pass blocks per SM at most (the partial array is [blocks][G][C][2])
constexpr int kBnBwdMaxBlk = 4;
// quad kernels, A/B switch FABRIC_B200_BWD2Q_V: default (7) = 4 channels per thread, even-size specialisation, loads of the next
// quad software-pipelined under the arithmetic of the current one (two register sets, 2 blocks per SM: measured 2.16 -> 2.01 ms
// for the eight launches of a step against 4 = the same without the pipeline at 4 blocks per SM); 5 = 4 with generic quads;
// 8 = the first form (8 channels per thread, 2.50 ms); 8 also switches the plain-case kernels back to the generic ones
static int bwd2q_vec() {
  static const int v = [] {
    const char* e = getenv("FABRIC_B200_BWD2Q_V");
    return (e && e[0] == '8') ? 8 : (e && e[0] == '5') ? 5 : (e && e[0] == '4') ? 4 : 7;
  }();
  return v;
}

int64_t fabric_b200_bn_bwd_ws_floats(int G, int C) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  return (int64_t)di.sms * kBnBwdMaxBlk * G * C * 2 + (int64_t)G * 3 * C;
}

// phase bits: 1 = reduce pass (partials into ws), 2 = finalize + apply pass.  `count_scale` multiplies the per-group element
// count (exact-global BatchNorm: world size, after the caller all-reduced the partials in ws); `grad_scale` multiplies
// dgamma / dbeta (1/world there, so that the SUM all-reduce of the gradient bucket restores the global value).
static int bn_relu_bwd_phases(int phase, const void* z, const void* a, const void* ga, int ga_groups, int ga_channels,
                              int mul_other, const void* gp, const float* scale, const float* shift, const float* mean,
                              const float* invstd, const float* gamma, void* dz, float* dgamma, float* dbeta, float* ws, int G,
                              int B, int H, int W, int C, float count_scale, float grad_scale, void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!z || !scale || !shift || !mean || !invstd || !gamma || !dz || !dgamma || !dbeta || !ws) return fail(FB_ERR_ARG, "null pointer");
  if (!ga && !gp) return fail(FB_ERR_ARG, "no gradient source");
  // product-fused encoder levels (both dates per thread): a == NULL means "recompute the activation from z"
  const bool dual_ = mul_other && G == 2 && ga && ga_groups == 1;
  if ((mul_other || gp) && !a && !dual_) return fail(FB_ERR_ARG, "activation tensor needed for product / pool routing");
  if (mul_other && G != 2) return fail(FB_ERR_SHAPE, "product fusion needs both date groups");
  if (C % 8 || 256 % (C / 8) || (ga && (ga_channels % 8 || ga_channels < C))) return fail(FB_ERR_SHAPE, "bad channels");
  if ((double)G * B * H * W * (ga ? ga_channels : C) / 8 >= 4.0e9) return fail(FB_ERR_SHAPE, "tensor too large for 32-bit indexing");
  BnBwd p;
  p.z = reinterpret_cast<const uint4*>(z), p.a = reinterpret_cast<const uint4*>(a);
  p.ga = reinterpret_cast<const uint4*>(ga), p.gp = reinterpret_cast<const uint4*>(gp);
  p.scale = scale, p.shift = shift, p.mean = mean, p.invstd = invstd;
  p.ga_groups = ga_groups, p.ga_c8 = ga_channels / 8, p.mul_other = mul_other;
  p.G = G, p.B = B, p.H = H, p.W = W, p.C = C;
  p.premasked = 0;
  cudaStream_t st = (cudaStream_t)stream;
  // product-fused encoder levels: both date groups per thread (see bn_bwd2_dy)
  const bool dual = mul_other && G == 2 && ga && ga_groups == 1;
  // quad kernels (activation recomputed): 4 channels per thread, 4 blocks of 128 threads per SM, where 128 % (C/4) == 0
  const bool quad = dual && !a;
  const int qv = (quad && bwd2q_vec() != 8 && C / 4 <= 128 && 128 % (C / 4) == 0) ? 4 : 8;
  const bool even = H % 2 == 0 && W % 2 == 0 && bwd2q_vec() != 5;   // (FABRIC_B200_BWD2Q_V=5: 4 channels, generic quads)
  const bool pipe = bwd2q_vec() == 7;                                // (=7: software-pipelined loads, 2 blocks per SM)
  // = resident blocks: one balanced wave (256 threads x 2 per SM; the quad kernels 128 threads x 2 or 4 per SM)
  // plain case (one gradient source, no product / pool): the specialised kernels, 3 blocks of 256 threads per SM
  const bool plain = ga && !gp && !mul_other && bwd2q_vec() != 8;
  const uint32_t npix_ = (uint32_t)B * H * W;
  const int nblk = di.sms * ((quad && qv == 4) ? (bwd2q_vec() == 7 ? 2 : 4) : plain ? 3 : 2);
  float* partial = ws;
  float* coef = ws + (size_t)nblk * G * C * 2;
  if (phase & 1) {
    const size_t sm2 = 256 * 16 * sizeof(float), sm1 = 128 * 2 * qv * sizeof(float);
    // (the recompute variants run nblk blocks of 128 threads: the partial layout [nblk][G][C][2] is the same)
    if (quad && gp && qv == 4 && even && pipe) bn_bwd2q_kernel<true, false, 4, true, true><<<nblk, 128, sm1, st>>>(p, nullptr, nullptr, partial);
    else if (quad && gp && qv == 4 && even) bn_bwd2q_kernel<true, false, 4, true><<<nblk, 128, sm1, st>>>(p, nullptr, nullptr, partial);
    else if (quad && gp && qv == 4) bn_bwd2q_kernel<true, false, 4><<<nblk, 128, sm1, st>>>(p, nullptr, nullptr, partial);
    else if (quad && gp) bn_bwd2q_kernel<true, false, 8><<<nblk, 128, sm1, st>>>(p, nullptr, nullptr, partial);
    else if (dual && gp) bn_bwd2_reduce_kernel<true, false><<<nblk, 256, sm2, st>>>(p, partial);
    else if (quad && qv == 4) bn_bwd2q_kernel<false, false, 4><<<nblk, 128, sm1, st>>>(p, nullptr, nullptr, partial);
    else if (quad) bn_bwd2q_kernel<false, false, 8><<<nblk, 128, sm1, st>>>(p, nullptr, nullptr, partial);
    else if (dual) bn_bwd2_reduce_kernel<false, false><<<nblk, 256, sm2, st>>>(p, partial);
    else if (plain) bn_bwd_reduce_plain_kernel<<<nblk, 256, sm2, st>>>(p.z, p.ga, p.ga_c8, ga_groups, scale, shift, mean, invstd, partial, G, npix_, C);
    else bn_bwd_reduce_kernel<<<nblk, 256, 256 * 16 * sizeof(float), st>>>(p, partial);
    FB_CUDA(cudaGetLastError());
  }
  if (phase & 2) {
    bn_bwd_finalize_kernel<<<(C + 7) / 8, 256, 0, st>>>(partial, nblk, G, C, (double)B * H * W * (double)count_scale, gamma, invstd,
                                                        mean, dgamma, dbeta, coef, grad_scale);
    FB_CUDA(cudaGetLastError());
    const size_t n = (size_t)B * H * W * (C / 8);   // per date group; 256 threads is a multiple of C/8 for every supported C
    const int g2 = ew_grid(n, 256, di.sms);
    uint4* dzo = reinterpret_cast<uint4*>(dz);
    const size_t nq = (size_t)B * ((H + 1) / 2) * ((W + 1) / 2) * (C / qv);   // quad-threads per date
    const int g1 = ew_grid(nq, 128, di.sms);
    if (quad && gp && qv == 4 && even && pipe) bn_bwd2q_kernel<true, true, 4, true, true><<<g1, 128, 0, st>>>(p, coef, dzo, nullptr);
    else if (quad && gp && qv == 4 && even) bn_bwd2q_kernel<true, true, 4, true><<<g1, 128, 0, st>>>(p, coef, dzo, nullptr);
    else if (quad && gp && qv == 4) bn_bwd2q_kernel<true, true, 4><<<g1, 128, 0, st>>>(p, coef, dzo, nullptr);
    else if (quad && gp) bn_bwd2q_kernel<true, true, 8><<<g1, 128, 0, st>>>(p, coef, dzo, nullptr);
    else if (dual && gp) bn_bwd2_apply_kernel<true, false><<<g2, 256, 0, st>>>(p, coef, dzo);
    else if (quad && qv == 4) bn_bwd2q_kernel<false, true, 4><<<g1, 128, 0, st>>>(p, coef, dzo, nullptr);
    else if (quad) bn_bwd2q_kernel<false, true, 8><<<g1, 128, 0, st>>>(p, coef, dzo, nullptr);
    else if (dual) bn_bwd2_apply_kernel<false, false><<<g2, 256, 0, st>>>(p, coef, dzo);
    else if (plain) bn_bwd_apply_plain_kernel<true><<<g2, 256, 0, st>>>(p.z, p.ga, p.ga_c8, ga_groups, scale, shift, coef, dzo, G, npix_, C);
    else bn_bwd_apply_kernel<<<ew_grid(n, 256, di.sms), 256, 0, st>>>(p, coef, reinterpret_cast<uint4*>(dz));
    FB_CUDA(cudaGetLastError());
  }
  return FB_OK;
}

int fabric_b200_bn_relu_bwd(const void* z, const void* a, const void* ga, int ga_groups, int ga_channels, int mul_other,
                            const void* gp, const float* scale, const float* shift, const float* mean, const float* invstd,
                            const float* gamma, void* dz, float* dgamma, float* dbeta, float* ws, int G, int B, int H, int W,
                            int C, void* stream) {
  return bn_relu_bwd_phases(3, z, a, ga, ga_groups, ga_channels, mul_other, gp, scale, shift, mean, invstd, gamma, dz, dgamma,
                            dbeta, ws, G, B, H, W, C, 1.f, 1.f, stream);
}

int fabric_b200_bn_relu_bwd_phase(int phase, const void* z, const void* a, const void* ga, int ga_groups, int ga_channels,
                                  int mul_other, const void* gp, const float* scale, const float* shift, const float* mean,
                                  const float* invstd, const float* gamma, void* dz, float* dgamma, float* dbeta, float* ws,
                                  int G, int B, int H, int W, int C, float count_scale, float grad_scale, void* stream) {
  if (phase != 1 && phase != 2) return fail(FB_ERR_ARG, "phase must be 1 (reduce) or 2 (finalize + apply)");
  return bn_relu_bwd_phases(phase, z, a, ga, ga_groups, ga_channels, mul_other, gp, scale, shift, mean, invstd, gamma, dz,
                            dgamma, dbeta, ws, G, B, H, W, C, count_scale, grad_scale, stream);
}

int fabric_b200_bn_bwd_from_partials(const void* z, const void* dy, const float* partial, int grid, int n_tile,
                                     const float* mean, const float* invstd, const float* gamma, void* dz, float* dgamma,
                                     float* dbeta, float* coef_ws, int G, int B, int H, int W, int C, float count_scale,
                                     float grad_scale, void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!z || !dy || !partial || !mean || !invstd || !gamma || !dz || !dgamma || !dbeta || !coef_ws) return fail(FB_ERR_ARG, "null pointer");
  if (C % 8 || 256 % (C / 8) || n_tile < 32 || C % n_tile || grid < C / n_tile || G < 1 || G > 2) return fail(FB_ERR_SHAPE, "bad shape");
  if ((double)G * B * H * W * C / 8 >= 4.0e9) return fail(FB_ERR_SHAPE, "tensor too large for 32-bit indexing");
  BnBwd p;
  p.z = reinterpret_cast<const uint4*>(z), p.a = nullptr;
  p.ga = reinterpret_cast<const uint4*>(dy), p.gp = nullptr;
  p.scale = mean, p.shift = mean, p.mean = mean, p.invstd = invstd;   // (scale / shift are not read: dy is pre-masked)
  p.ga_groups = G, p.ga_c8 = C / 8, p.mul_other = 0;
  p.G = G, p.B = B, p.H = H, p.W = W, p.C = C;
  p.premasked = 1;
  cudaStream_t st = (cudaStream_t)stream;
  bn_bwd_finalize_kernel<<<(C + 7) / 8, 256, 0, st>>>(partial, grid, G, C, (double)B * H * W * (double)count_scale, gamma, invstd,
                                                      mean, dgamma, dbeta, coef_ws, grad_scale, n_tile);
  FB_CUDA(cudaGetLastError());
  const size_t n = (size_t)B * H * W * (C / 8);
  if (bwd2q_vec() != 8)
    bn_bwd_apply_plain_kernel<false><<<ew_grid(n, 256, di.sms), 256, 0, st>>>(p.z, p.ga, p.ga_c8, G, mean, mean, coef_ws,
                                                                              reinterpret_cast<uint4*>(dz), G, (uint32_t)B * H * W, C);
  else
    bn_bwd_apply_kernel<<<ew_grid(n, 256, di.sms), 256, 0, st>>>(p, coef_ws, reinterpret_cast<uint4*>(dz));
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int64_t fabric_b200_bn_bwd_partial_floats(int G, int C) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  return (int64_t)di.sms * kBnBwdMaxBlk * G * C * 2;
}

int fabric_b200_bn_apply_relu_head(const void* z, const float* scale, const float* shift, void* a, const float* head_w,
                                   const float* head_b, float* logits, int B, int H, int W, int C, void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!z || !scale || !shift || !head_w || !head_b || !logits) return fail(FB_ERR_ARG, "null pointer");   // (a may be NULL)
  if (C != 64) return fail(FB_ERR_SHAPE, "the fused head is built for 64 channels (outc = outconv(64, 2), bidate_model.py:20)");
  if ((double)B * H * W * 8 >= 4.0e9) return fail(FB_ERR_SHAPE, "tensor too large for 32-bit indexing");
  const uint32_t npix = (uint32_t)B * H * W;
  bn_apply_head_kernel<<<ew_grid((size_t)npix * 8, 256, di.sms), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(z), scale, shift, reinterpret_cast<uint4*>(a), head_w, head_b, logits, npix, (uint32_t)H * W);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int64_t fabric_b200_bn_head_bwd_ws_floats(void) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  return (int64_t)di.sms * 2 * kHeadPartial + 3 * 64 + 128;
}

int fabric_b200_bn_head_bwd(int phase, const float* dlogits, const void* z, const float* scale, const float* shift,
                            const float* mean, const float* invstd, const float* gamma, const float* head_w, void* dz,
                            float* dgamma, float* dbeta, float* dw, float* db, float* ws, int B, int H, int W, int C,
                            float count_scale, float grad_scale, void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!dlogits || !z || !scale || !shift || !mean || !invstd || !gamma || !head_w || !dz || !dgamma || !dbeta || !dw || !db || !ws)
    return fail(FB_ERR_ARG, "null pointer");
  if (C != 64) return fail(FB_ERR_SHAPE, "the fused head is built for 64 channels");
  if (phase < 1 || phase > 3) return fail(FB_ERR_ARG, "phase must be 1 (reduce), 2 (finalize + apply) or 3 (both)");
  if ((double)B * H * W * 8 >= 4.0e9) return fail(FB_ERR_SHAPE, "tensor too large for 32-bit indexing");
  HeadBwd p;
  p.z = reinterpret_cast<const uint4*>(z), p.dlogits = dlogits, p.scale = scale, p.shift = shift, p.mean = mean, p.invstd = invstd;
  p.hw = head_w, p.npix = (uint32_t)B * H * W, p.plane = (uint32_t)H * W;
  p.plane_magic = ((unsigned long long)p.npix * p.plane < (1ull << 40)) ? ((1ull << 40) + p.plane - 1) / p.plane : 0ull;
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = di.sms * 2;
  float* partial = ws;
  float* coef = ws + (size_t)nblk * kHeadPartial;
  if (phase & 1) {
    bn_head_bwd_reduce_kernel<<<nblk, 256, 256 * 34 * sizeof(float), st>>>(p, partial);
    FB_CUDA(cudaGetLastError());
  }
  if (phase & 2) {
    bn_head_bwd_finalize_kernel<<<1, 256, 0, st>>>(partial, nblk, (double)B * H * W * (double)count_scale, gamma, invstd, mean,
                                                   dgamma, dbeta, coef, dw, db, grad_scale);
    FB_CUDA(cudaGetLastError());
    bn_head_bwd_apply_kernel<<<ew_grid((size_t)p.npix * 8, 256, di.sms), 256, 0, st>>>(p, coef, reinterpret_cast<uint4*>(dz));
    FB_CUDA(cudaGetLastError());
  }
  return FB_OK;
}

int fabric_b200_up_input_bwd(const void* dcat, void* dlow, int B, int H, int W, int Cs, int h, int w, int Cl, void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!dcat || !dlow) return fail(FB_ERR_ARG, "null pointer");
  if (Cs % 8 || Cl % 8 || 2 * h > H || 2 * w > W) return fail(FB_ERR_SHAPE, "bad shape");
  if (h <= 65535 && B <= 65535 && (double)B * H * W * (Cs + Cl) / 8 < 4.0e9) {
    const int Cl8 = Cl / 8;
    int shift = -1;
    for (int sft = 0; sft < 12; ++sft)
      if ((1 << sft) == Cl8) shift = sft;
    const float sy = (2 * h > 1) ? (float)(h - 1) / (float)(2 * h - 1) : 0.f;
    const float sx = (2 * w > 1) ? (float)(w - 1) / (float)(2 * w - 1) : 0.f;
    dim3 grid(((size_t)w * Cl8 + 255) / 256, h, B);
    up_input_bwd_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(dcat),
                                                                     reinterpret_cast<uint4*>(dlow), H, W, Cs / 8, h, w, Cl8, shift,
                                                                     (H - 2 * h) / 2, (W - 2 * w) / 2, sy, sx);
    FB_CUDA(cudaGetLastError());
    return FB_OK;
  }
  const size_t n = (size_t)B * h * w * (Cl / 8);
  up_input_bwd_kernel<<<ew_grid(n, 256, di.sms), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(dcat), reinterpret_cast<uint4*>(dlow), B, H, W, Cs, h, w, Cl);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int fabric_b200_wgrad_reduce(const float* ws, int splits, int Cout, int Cin, int CinPad, float* dw, void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!ws || !dw) return fail(FB_ERR_ARG, "null pointer");
  if (CinPad % 16 || CinPad < Cin) return fail(FB_ERR_SHAPE, "CinPad must be a multiple of 16 and >= Cin");
  const size_t n = (size_t)Cout * 9 * CinPad / 4;
  wgrad_reduce_kernel<<<ew_grid(n, 256, di.sms), 256, 0, (cudaStream_t)stream>>>(ws, splits, Cout, Cin, CinPad, dw);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int fabric_b200_wgrad_reduce_swapped(const float* ws, int splits, int Cout, int Cin, float* dw, void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!ws || !dw) return fail(FB_ERR_ARG, "null pointer");
  const size_t n = (size_t)Cout * Cin * 9;
  wgrad_reduce_swapped_kernel<<<ew_grid(n, 256, di.sms), 256, 0, (cudaStream_t)stream>>>(ws, splits, Cout, Cin, dw);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

}  // extern "C"
