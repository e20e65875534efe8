// Training-path kernels of fabric_b200 that are HBM-bound (everything except the convolutions):
// BatchNorm batch statistics / apply / backward, the segmentation losses and their gradient, the 1x1 head backward,
// and the adjoint of the decoder-input builder.  See include/fabric_b200.h for the reference constructs replaced.
#include "host_common.cuh"
#include "ptx.cuh"

using namespace fbh;
using fb::pack8;
using fb::unpack8;

namespace {

// ================================================================================================ BN forward
// reference models/unet_parts.py:14,17 (nn.BatchNorm2d in training mode), per date group.
__global__ void bn_finalize_kernel(const float* __restrict__ stats, int grid, int n_tile, int C, int G, double count,
                                   const float* __restrict__ conv_bias, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* running_mean, float* running_var,
                                   long long* nbt, float momentum, float eps, float* scale, float* shift, float* mean,
                                   float* invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int ntiles = C / n_tile, nt = c / n_tile, lc = c % n_tile;
  float rm = running_mean[c], rv = running_var[c];
  const float b = conv_bias ? conv_bias[c] : 0.f;
  for (int g = 0; g < G; ++g) {  // date 1 first, then date 2: the order in which the reference calls the encoder
    double s1 = 0.0, s2 = 0.0;
    for (int cta = nt; cta < grid; cta += ntiles) {
      const float* p = stats + (((size_t)cta * 2 + g) * n_tile + lc) * 2;
      s1 += p[0];
      s2 += p[1];
    }
    const double m = s1 / count;
    double var = s2 / count - m * m;
    if (var < 0.0) var = 0.0;
    const float inv = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * inv;
    scale[g * C + c] = sc;
    shift[g * C + c] = beta[c] - (float)m * sc;  // the conv bias cancels against the batch mean
    mean[g * C + c] = (float)m;
    invstd[g * C + c] = inv;
    rm = (1.f - momentum) * rm + momentum * ((float)m + b);
    rv = (1.f - momentum) * rv + momentum * (float)(var * (count / (count > 1.0 ? count - 1.0 : 1.0)));
  }
  running_mean[c] = rm;
  running_var[c] = rv;
  if (c == 0 && nbt) *nbt += G;
}

// a = relu(z * scale_g + shift_g) (+ 2x2 max pool).  One thread = 8 channels of one 2x2 pixel quad.
__global__ void bn_apply_kernel(const uint4* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift,
                                uint4* __restrict__ a, uint4* __restrict__ pool, int G, int B, int H, int W, int C) {
  const int C8 = C / 8, Hq = (H + 1) / 2, Wq = (W + 1) / 2, Hp = H / 2, Wp = W / 2;
  const size_t total = (size_t)G * B * Hq * Wq * C8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = i % C8;
    size_t q = i / C8;
    const int qx = q % Wq;
    q /= Wq;
    const int qy = q % Hq;
    q /= Hq;
    const int b = q % B;
    const int g = q / B;
    float sc[8], sh[8], m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = scale[g * C + c8 * 8 + j];
      sh[j] = shift[g * C + c8 * 8 + j];
      m[j] = 0.f;  // post-ReLU values are >= 0
    }
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const int y = 2 * qy + (d >> 1), x = 2 * qx + (d & 1);
      if (y < H && x < W) {
        const size_t o = ((((size_t)g * B + b) * H + y) * W + x) * C8 + c8;
        float f[8];
        unpack8(z[o], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          f[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
          m[j] = fmaxf(m[j], f[j]);
        }
        a[o] = pack8(f);
      }
    }
    if (pool && qy < Hp && qx < Wp) pool[((((size_t)g * B + b) * Hp + qy) * Wp + qx) * C8 + c8] = pack8(m);
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
                                  float gamma, float* __restrict__ dlogits, float* __restrict__ partial) {
  __shared__ float red[32];
  const size_t plane = (size_t)H * W, total = (size_t)B * plane;
  const float invn = 1.f / (float)total;
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
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += ws[(size_t)i * m + j];
    out[j] = (float)s;
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
};

// dy for the 8 channels c8 of the 2x2 quad (qy,qx) of image b in group g.  `av` = own activations of the quad (only if
// needed), returns validity per pixel.
__device__ __forceinline__ void bn_bwd_dy_quad(const BnBwd& p, int g, int b, int qy, int qx, int c8, float (&dy)[4][8],
                                               float (&zf)[4][8], bool (&valid)[4]) {
  const int C8 = p.C / 8;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sc[j] = p.scale[g * p.C + c8 * 8 + j];
    sh[j] = p.shift[g * p.C + c8 * 8 + j];
  }
  float av[4][8];
  const bool need_a = p.gp != nullptr;
#pragma unroll
  for (int d = 0; d < 4; ++d) {
    const int y = 2 * qy + (d >> 1), x = 2 * qx + (d & 1);
    valid[d] = y < p.H && x < p.W;
#pragma unroll
    for (int j = 0; j < 8; ++j) dy[d][j] = 0.f, zf[d][j] = 0.f, av[d][j] = -1.f;
    if (!valid[d]) continue;
    const size_t pix = (((size_t)b * p.H + y) * p.W + x);
    const size_t o = ((size_t)g * p.B * p.H * p.W + pix) * C8 + c8;
    unpack8(p.z[o], zf[d]);
    if (need_a) unpack8(p.a[o], av[d]);
    if (p.ga) {
      const size_t go = ((size_t)(p.ga_groups == 1 ? 0 : g) * p.B * p.H * p.W + pix) * p.ga_c8 + c8;
      unpack8(p.ga[go], dy[d]);
      if (p.mul_other) {
        float ao[8];
        unpack8(p.a[((size_t)(1 - g) * p.B * p.H * p.W + pix) * C8 + c8], ao);
#pragma unroll
        for (int j = 0; j < 8; ++j) dy[d][j] *= ao[j];
      }
    }
  }
  if (p.gp && qy < p.H / 2 && qx < p.W / 2) {
    float gpv[8];
    unpack8(p.gp[((((size_t)g * p.B + b) * (p.H / 2) + qy) * (p.W / 2) + qx) * C8 + c8], gpv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // nn.MaxPool2d backward routes to the first maximum in window scan order
      int best = 0;
      float m = av[0][j];
#pragma unroll
      for (int d = 1; d < 4; ++d)
        if (av[d][j] > m) m = av[d][j], best = d;
#pragma unroll
      for (int d = 0; d < 4; ++d)
        if (d == best) dy[d][j] += gpv[j];
    }
  }
#pragma unroll
  for (int d = 0; d < 4; ++d)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (!(fmaf(zf[d][j], sc[j], sh[j]) > 0.f)) dy[d][j] = 0.f;
}

// pass 1: partial[blk][g][c][2] = (sum dy, sum dy * xhat)
__global__ void bn_bwd_reduce_kernel(BnBwd p, float* __restrict__ partial) {
  extern __shared__ float sm[];  // [blockDim][16]
  const int C8 = p.C / 8, Hq = (p.H + 1) / 2, Wq = (p.W + 1) / 2;
  const int c8 = threadIdx.x % C8, lane_q = threadIdx.x / C8, qpb = blockDim.x / C8;
  const size_t nquads = (size_t)p.B * Hq * Wq;
  for (int g = 0; g < p.G; ++g) {
    float s1[8], s2[8], mu[8], is[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s1[j] = s2[j] = 0.f;
      mu[j] = p.mean[g * p.C + c8 * 8 + j];
      is[j] = p.invstd[g * p.C + c8 * 8 + j];
    }
    for (size_t q = (size_t)blockIdx.x * qpb + lane_q; q < nquads; q += (size_t)gridDim.x * qpb) {
      const int qx = q % Wq, qy = (q / Wq) % Hq, b = q / ((size_t)Wq * Hq);
      float dy[4][8], zf[4][8];
      bool valid[4];
      bn_bwd_dy_quad(p, g, b, qy, qx, c8, dy, zf, valid);
#pragma unroll
      for (int d = 0; d < 4; ++d)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s1[j] += dy[d][j];
          s2[j] = fmaf(dy[d][j], (zf[d][j] - mu[j]) * is[j], s2[j]);
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
      for (int l = 0; l < qpb; ++l) s += sm[(l * C8 + (c >> 3)) * 16 + k * 8 + (c & 7)];
      dst[i] = s;
    }
  }
}

// pass 2: reduce partials; dgamma, dbeta; per-group coefficients for the apply pass
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int nblk, int G, int C, double count,
                                       const float* __restrict__ gamma, const float* __restrict__ invstd,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ coef) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double dg = 0.0, db = 0.0;
  for (int g = 0; g < G; ++g) {
    double s1 = 0.0, s2 = 0.0;
    for (int b = 0; b < nblk; ++b) {
      const float* p = partial + (((size_t)b * G + g) * C + c) * 2;
      s1 += p[0];
      s2 += p[1];
    }
    dg += s2;
    db += s1;
    coef[(g * 3 + 0) * C + c] = gamma[c] * invstd[g * C + c];  // dz = k0 * (dy - k1 - xhat * k2)
    coef[(g * 3 + 1) * C + c] = (float)(s1 / count);
    coef[(g * 3 + 2) * C + c] = (float)(s2 / count);
  }
  dgamma[c] = (float)dg;
  dbeta[c] = (float)db;
}

// pass 3: dz (bf16)
__global__ void bn_bwd_apply_kernel(BnBwd p, const float* __restrict__ coef, uint4* __restrict__ dz) {
  const int C8 = p.C / 8, Hq = (p.H + 1) / 2, Wq = (p.W + 1) / 2;
  const size_t total = (size_t)p.G * p.B * Hq * Wq * C8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c8 = i % C8;
    size_t q = i / C8;
    const int qx = q % Wq;
    q /= Wq;
    const int qy = q % Hq;
    q /= Hq;
    const int b = q % p.B;
    const int g = q / p.B;
    float dy[4][8], zf[4][8];
    bool valid[4];
    bn_bwd_dy_quad(p, g, b, qy, qx, c8, dy, zf, valid);
    float k0[8], k1[8], k2[8], mu[8], is[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c8 * 8 + j;
      k0[j] = coef[(g * 3 + 0) * p.C + c];
      k1[j] = coef[(g * 3 + 1) * p.C + c];
      k2[j] = coef[(g * 3 + 2) * p.C + c];
      mu[j] = p.mean[g * p.C + c];
      is[j] = p.invstd[g * p.C + c];
    }
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      if (!valid[d]) continue;
      const int y = 2 * qy + (d >> 1), x = 2 * qx + (d & 1);
      float r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = k0[j] * (dy[d][j] - k1[j] - (zf[d][j] - mu[j]) * is[j] * k2[j]);
      dz[((((size_t)g * p.B + b) * p.H + y) * p.W + x) * C8 + c8] = pack8(r);
    }
  }
}

// ================================================================================================ decoder-input adjoint
// dlow[b][i][j][c] = sum_{u,v} wy(u,i) wx(v,j) dcat[b][u+padT][v+padL][Cs+c]   (adjoint of bilinear x2 + pad)
__global__ void up_input_bwd_kernel(const uint4* __restrict__ dcat, uint4* __restrict__ dlow, int B, int H, int W, int Cs,
                                    int h, int w, int Cl) {
  const int Ct8 = (Cs + Cl) / 8, Cs8 = Cs / 8, Cl8 = Cl / 8;
  const int padT = (H - 2 * h) / 2, padL = (W - 2 * w) / 2;
  const float sy = (2 * h > 1) ? (float)(h - 1) / (float)(2 * h - 1) : 0.f;
  const float sx = (2 * w > 1) ? (float)(w - 1) / (float)(2 * w - 1) : 0.f;
  const size_t total = (size_t)B * h * w * Cl8;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c8 = idx % Cl8;
    size_t r = idx / Cl8;
    const int j = r % w;
    r /= w;
    const int i = r % h;
    const int b = r / h;
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
        unpack8(dcat[(((size_t)b * H + u + padT) * W + v + padL) * Ct8 + Cs8 + c8], f);
        const float ww = wy * wx;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(ww, f[k], acc[k]);
      }
    }
    dlow[idx] = pack8(acc);
  }
}

// fp32 [S][Cout][9][CinPad] split-K partials -> nn.Conv2d weight gradient [Cout][Cin][3][3]
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, int S, int Cout, int Cin, int CinPad, float* __restrict__ dw) {
  const size_t n = (size_t)Cout * Cin * 9;
  const size_t slab = (size_t)Cout * 9 * CinPad;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int tap = i % 9, ci = (i / 9) % Cin, co = i / ((size_t)9 * Cin);
    const size_t o = ((size_t)co * 9 + tap) * CinPad + ci;
    float s = 0.f;
    for (int k = 0; k < S; ++k) s += ws[k * slab + o];
    dw[i] = s;
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
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      stats_ws, grid, n_tile, C, G, (double)count_per_group, conv_bias, gamma, beta, running_mean, running_var,
      reinterpret_cast<long long*>(num_batches_tracked), momentum, eps, scale, shift, mean, invstd);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int fabric_b200_bn_apply_relu(const void* z, const float* scale, const float* shift, void* a, void* pool_out, int G, int B,
                              int H, int W, int C, void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!z || !scale || !shift || !a) return fail(FB_ERR_ARG, "null pointer");
  if (C % 8) return fail(FB_ERR_SHAPE, "C must be a multiple of 8");
  const size_t n = (size_t)G * B * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
  bn_apply_kernel<<<ew_grid(n, 256, di.sms), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const uint4*>(z), scale, shift, reinterpret_cast<uint4*>(a), reinterpret_cast<uint4*>(pool_out), G, B,
      H, W, C);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int64_t fabric_b200_seg_loss_ws_floats(int B, int H, int W) {
  const int rows = B * H;
  const int rpb = (rows + 295) / 296;
  const int nblk = (rows + rpb - 1) / rpb;
  return (int64_t)nblk * 6 * W + 4 * (int64_t)W + 1024;
}

int fabric_b200_seg_loss_fwd_bwd(int kind, float alpha, float beta, float gamma, float eps, const float* logits,
                                 const int64_t* labels, int label_ndim, int B, int H, int W, float* loss_out, float* dlogits,
                                 float* ws, void* stream) {
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
    seg_loss_partial_kernel<<<nblk, 256, 0, st>>>(logits, reinterpret_cast<const long long*>(labels), B, H, W, rpb, partial);
    FB_CUDA(cudaGetLastError());
    const size_t smem = 6 * (size_t)W * sizeof(float);
    FB_CUDA(cudaFuncSetAttribute(seg_loss_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    seg_loss_finalize_kernel<<<1, 256, smem, st>>>(partial, nblk, W, label_ndim, kind, alpha, beta, eps, coef, loss_out);
    FB_CUDA(cudaGetLastError());
    seg_loss_grad_kernel<<<ew_grid(n, 256, di.sms), 256, 0, st>>>(logits, reinterpret_cast<const long long*>(labels), coef, B,
                                                                H, W, dlogits);
    FB_CUDA(cudaGetLastError());
  } else if (kind == 3 || kind == 4) {  // focal / cross entropy
    const int nblk = 296;
    focal_loss_kernel<<<nblk, 256, 0, st>>>(logits, reinterpret_cast<const long long*>(labels), B, H, W,
                                            kind == 4 ? 0.f : gamma, dlogits, ws);
    FB_CUDA(cudaGetLastError());
    reduce_partials_kernel<<<1, 32, 0, st>>>(ws, nblk, 1, loss_out);
    FB_CUDA(cudaGetLastError());
  } else {
    return fail(FB_ERR_ARG, "unknown loss kind %d", kind);
  }
  return FB_OK;
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

int64_t fabric_b200_bn_bwd_ws_floats(int G, int C) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  return (int64_t)di.sms * 4 * G * C * 2 + (int64_t)G * 3 * C;
}

int fabric_b200_bn_relu_bwd(const void* z, const void* a, const void* ga, int ga_groups, int ga_channels, int mul_other,
                            const void* gp, const float* scale, const float* shift, const float* mean, const float* invstd,
                            const float* gamma, void* dz, float* dgamma, float* dbeta, float* ws, int G, int B, int H, int W,
                            int C, void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!z || !scale || !shift || !mean || !invstd || !gamma || !dz || !dgamma || !dbeta || !ws) return fail(FB_ERR_ARG, "null pointer");
  if (!ga && !gp) return fail(FB_ERR_ARG, "no gradient source");
  if ((mul_other || gp) && !a) return fail(FB_ERR_ARG, "activation tensor needed for product / pool routing");
  if (mul_other && G != 2) return fail(FB_ERR_SHAPE, "product fusion needs both date groups");
  if (C % 8 || 256 % (C / 8) || (ga && (ga_channels % 8 || ga_channels < C))) return fail(FB_ERR_SHAPE, "bad channels");
  BnBwd p;
  p.z = reinterpret_cast<const uint4*>(z), p.a = reinterpret_cast<const uint4*>(a);
  p.ga = reinterpret_cast<const uint4*>(ga), p.gp = reinterpret_cast<const uint4*>(gp);
  p.scale = scale, p.shift = shift, p.mean = mean, p.invstd = invstd;
  p.ga_groups = ga_groups, p.ga_c8 = ga_channels / 8, p.mul_other = mul_other;
  p.G = G, p.B = B, p.H = H, p.W = W, p.C = C;
  cudaStream_t st = (cudaStream_t)stream;
  const int nblk = di.sms * 4;
  float* partial = ws;
  float* coef = ws + (size_t)nblk * G * C * 2;
  bn_bwd_reduce_kernel<<<nblk, 256, 256 * 16 * sizeof(float), st>>>(p, partial);
  FB_CUDA(cudaGetLastError());
  bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(partial, nblk, G, C, (double)B * H * W, gamma, invstd, dgamma, dbeta,
                                                          coef);
  FB_CUDA(cudaGetLastError());
  const size_t n = (size_t)G * B * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
  bn_bwd_apply_kernel<<<ew_grid(n, 256, di.sms), 256, 0, st>>>(p, coef, reinterpret_cast<uint4*>(dz));
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int fabric_b200_up_input_bwd(const void* dcat, void* dlow, int B, int H, int W, int Cs, int h, int w, int Cl, void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!dcat || !dlow) return fail(FB_ERR_ARG, "null pointer");
  if (Cs % 8 || Cl % 8 || 2 * h > H || 2 * w > W) return fail(FB_ERR_SHAPE, "bad shape");
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
  const size_t n = (size_t)Cout * Cin * 9;
  wgrad_reduce_kernel<<<ew_grid(n, 256, di.sms), 256, 0, (cudaStream_t)stream>>>(ws, splits, Cout, Cin, CinPad, dw);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

}  // extern "C"
