// Full-scene inference helpers: tile gather straight from the CHW scene into packed NHWC bf16 batches
// (with the per-band z-score of utils/dataloaders.py:94-99 fused), argmax + confusion counts (train.py:96-106), and
// the mask reassembly of utils/inference.py:184-236.
#include "host_common.cuh"
#include "ptx.cuh"

using namespace fbh;

namespace {

// scene [C][H][W] (fp32 or uint16) -> tiles bf16 [N][p][p][Cpad]; one block = up to 256 pixels of one tile row.
template <typename T>
__global__ void __launch_bounds__(256) gather_tiles_kernel(const T* __restrict__ scene, const int* __restrict__ origins,
                                                           __nv_bfloat16* __restrict__ dst, const float* __restrict__ mean,
                                                           const float* __restrict__ inv_std, int C, int Cpad, int H, int W,
                                                           int p) {
  extern __shared__ float tile[];  // [C][257]
  const int n = blockIdx.z, y = blockIdx.y, x0 = blockIdx.x * 256;
  const int nx = min(256, p - x0);
  const int oy = origins[2 * n], ox = origins[2 * n + 1];
  const T* in = scene + (size_t)(oy + y) * W + ox + x0;
  for (int i = threadIdx.x; i < C * 256; i += blockDim.x) {
    const int c = i >> 8, x = i & 255;
    if (x < nx) {
      float v = (float)in[(size_t)c * H * W + x];
      if (mean) v = (v - mean[c]) * inv_std[c];
      tile[c * 257 + x] = v;
    }
  }
  __syncthreads();
  uint4* out = reinterpret_cast<uint4*>(dst + (((size_t)n * p + y) * p + x0) * Cpad);
  const int C8 = Cpad / 8;
  for (int i = threadIdx.x; i < nx * C8; i += blockDim.x) {
    const int x = i / C8, c0 = (i % C8) * 8;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (c0 + j < C) ? tile[(c0 + j) * 257 + x] : 0.f;
    out[i] = fb::pack8(f);
  }
}

// The same gather for <= 16 bands packed to 16 channels: one thread = one pixel (coalesced 128-byte reads per band plane
// across a warp, one 32-byte store per lane) -- the smem-transpose form above is issue-bound at ~25 % of HBM bandwidth.
template <typename T>
__global__ void __launch_bounds__(256) gather_tiles16_kernel(const T* __restrict__ scene, const int* __restrict__ origins,
                                                             uint32_t* __restrict__ dst, const float* __restrict__ mean,
                                                             const float* __restrict__ inv_std, int C, int H, int W, int p) {
  const int n = blockIdx.z, y = blockIdx.y, x = blockIdx.x * 256 + threadIdx.x;
  if (x >= p) return;
  const int oy = origins[2 * n], ox = origins[2 * n + 1];
  const T* in = scene + (size_t)(oy + y) * W + ox + x;
  const size_t plane = (size_t)H * W;
  float v[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    v[c] = c < C ? (float)__ldg(in + c * plane) : 0.f;
    if (mean && c < C) v[c] = (v[c] - __ldg(mean + c)) * __ldg(inv_std + c);
  }
  uint32_t r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = fb::pack_bf16x2(v[2 * j], v[2 * j + 1]);
  uint32_t* out = dst + (((size_t)n * p + y) * p + x) * 8;
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(out), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
               "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

// logits NCHW fp32 [B][2][H][W] -> mask uint8 [B][H][W] (torch.max(.,1) index: ties -> class 0) and, with labels,
// counts[0..3] += (TP, FP, FN, TN) for the positive class 1
__global__ void argmax_metrics_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                                      unsigned char* __restrict__ mask, unsigned long long* __restrict__ counts, int B,
                                      int plane) {
  __shared__ unsigned int sc[4];
  if (threadIdx.x < 4) sc[threadIdx.x] = 0;
  __syncthreads();
  const size_t total = (size_t)B * plane;
  unsigned int c[4] = {0, 0, 0, 0};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / plane, o = i % plane;
    const float l0 = logits[(b * 2) * plane + o], l1 = logits[(b * 2 + 1) * plane + o];
    const unsigned char m = l1 > l0 ? 1 : 0;
    if (mask) mask[i] = m;
    if (labels) {
      const bool t = labels[i] != 0;
      c[m ? (t ? 0 : 1) : (t ? 2 : 3)]++;
    }
  }
  if (labels) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      unsigned int v = c[k];
      for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sc[k], v);
    }
    __syncthreads();
    if (threadIdx.x < 4 && sc[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)sc[threadIdx.x]);
  }
}

// masks uint8 [N][p][p] -> canvas uint8 [H][W] at the tiles' origins (tiles [first, first+count) must not overlap)
__global__ void scatter_tiles_kernel(const unsigned char* __restrict__ masks, const int* __restrict__ origins,
                                     unsigned char* __restrict__ canvas, int first, int count, int p, int H, int W) {
  const size_t total = (size_t)count * p * p;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = i % p, y = (i / p) % p, n = first + (int)(i / ((size_t)p * p));
    const int gy = origins[2 * n] + y, gx = origins[2 * n + 1] + x;
    if (gy < H && gx < W) canvas[(size_t)gy * W + gx] = masks[((size_t)n * p + y) * p + x];
  }
}

}  // namespace

extern "C" {

int fabric_b200_gather_tiles(const void* scene, int scene_dtype, const int* origins, void* dst, const float* mean,
                             const float* inv_std, int N, int C, int Cpad, int H, int W, int p, void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!scene || !origins || !dst) return fail(FB_ERR_ARG, "null pointer");
  if (N < 1 || C < 1 || C > 32 || Cpad < C || Cpad % 8 || p < 1 || p > H || p > W || N > 65535 || p > 65535)
    return fail(FB_ERR_SHAPE, "bad shape");
  if ((mean == nullptr) != (inv_std == nullptr)) return fail(FB_ERR_ARG, "mean and inv_std go together");
  if (!aligned16(dst)) return fail(FB_ERR_ALIGN, "dst must be 16-byte aligned");
  dim3 grid((p + 255) / 256, p, N);
  const size_t smem = (size_t)C * 257 * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  if (C <= 16 && Cpad == 16 && ((uintptr_t)dst & 31) == 0 && (scene_dtype == 0 || scene_dtype == 1)) {
    if (scene_dtype == 0)
      gather_tiles16_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(scene), origins,
                                                         reinterpret_cast<uint32_t*>(dst), mean, inv_std, C, H, W, p);
    else
      gather_tiles16_kernel<unsigned short><<<grid, 256, 0, st>>>(reinterpret_cast<const unsigned short*>(scene), origins,
                                                                  reinterpret_cast<uint32_t*>(dst), mean, inv_std, C, H, W, p);
    FB_CUDA(cudaGetLastError());
    return FB_OK;
  }
  if (scene_dtype == 0)
    gather_tiles_kernel<float><<<grid, 256, smem, st>>>(reinterpret_cast<const float*>(scene), origins,
                                                        reinterpret_cast<__nv_bfloat16*>(dst), mean, inv_std, C, Cpad, H, W, p);
  else if (scene_dtype == 1)
    gather_tiles_kernel<unsigned short><<<grid, 256, smem, st>>>(reinterpret_cast<const unsigned short*>(scene), origins,
                                                                 reinterpret_cast<__nv_bfloat16*>(dst), mean, inv_std, C, Cpad,
                                                                 H, W, p);
  else
    return fail(FB_ERR_ARG, "scene_dtype must be 0 (fp32) or 1 (uint16)");
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int fabric_b200_argmax_metrics(const float* logits, const int64_t* labels, uint8_t* mask, uint64_t* counts, int B, int H, int W,
                               void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!logits || (!mask && !labels) || (labels && !counts)) return fail(FB_ERR_ARG, "null pointer");
  const size_t n = (size_t)B * H * W;
  argmax_metrics_kernel<<<ew_grid(n, 256, di.sms), 256, 0, (cudaStream_t)stream>>>(
      logits, reinterpret_cast<const long long*>(labels), mask, reinterpret_cast<unsigned long long*>(counts), B, H * W);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int fabric_b200_scatter_tiles(const uint8_t* masks, const int* origins, uint8_t* canvas, int first, int count, int p, int H,
                              int W, void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!masks || !origins || !canvas) return fail(FB_ERR_ARG, "null pointer");
  if (count < 1) return FB_OK;
  const size_t n = (size_t)count * p * p;
  scatter_tiles_kernel<<<ew_grid(n, 256, di.sms), 256, 0, (cudaStream_t)stream>>>(masks, origins, canvas, first, count, p, H, W);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

}  // extern "C"

// ---- plain SGD over many tensors in one launch (train.py:55,95: optim.SGD(lr), no momentum / weight decay) ----------
namespace {
struct SgdChunk {
  float* p;
  const float* g;
  int n;
  int pad;
};
__global__ void sgd_multi_kernel(const SgdChunk* __restrict__ chunks, float lr_scaled) {
  const SgdChunk c = chunks[blockIdx.x];
  for (int i = threadIdx.x; i < c.n; i += blockDim.x) c.p[i] -= lr_scaled * c.g[i];
}
}  // namespace

extern "C" int fabric_b200_sgd_step(const void* chunks, int n_chunks, float lr, float grad_scale, void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!chunks || n_chunks < 1) return fail(FB_ERR_ARG, "no chunks");
  sgd_multi_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const SgdChunk*>(chunks), lr * grad_scale);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}


// ---- the whole optimizer side of a training step in ONE launch ------------------------------------------------------
// Records of 64 bytes, one CTA each (<= 65536 elements):
//   mode 0: p -= lr * grad_scale * g                                  (plain SGD, train.py:55,95)
//   mode 1: p *= stats_scale                                          (BatchNorm running statistics after the SUM all-reduce)
//   mode 2: mode 0 for a 3x3 conv weight [Cout][Cin][3][3] AND refresh of its packed bf16 copies: wf[Cout][9][CinPad]
//           (forward operand) and wd[Cin][9][Cout] with flipped taps (data-gradient operand); `off` = element offset of
//           this record inside the weight tensor.  Replaces the 35 pack_weight launches a training step used to make.
//   mode 3: the same per (16 output channels x <= 128 input channels) tile through shared memory (see the kernel)
namespace {
struct StepChunk {
  float* p;
  const float* g;
  __nv_bfloat16* wf;
  __nv_bfloat16* wd;
  int n, mode, off, Cout, Cin, CinPad, pad0, pad1;
};
static_assert(sizeof(StepChunk) == 64, "record layout is part of the ABI (fabric_b200/distributed.py packs it)");

// mode 3: one TILE of a 3x3 conv weight = kUpdCo output channels x <= kUpdCi input channels x 9 taps.  p / g = bases of the
// weight and of its gradient, off = first output channel, n = output channels in the tile, pad0 = first input channel,
// pad1 = input channels in the tile.  The fp32 update reads and writes runs of pad1 * 9 contiguous floats; the two packed
// copies are written from a shared-memory copy of the tile so that both leave in contiguous runs (wf: pad1 channels =
// <= 256 B, wd: n output channels = 32 B) -- mode 2 wrote both element by element (2-byte scattered stores: the launch
// took 0.23 ms for 0.21 GB, 0.9 TB/s).
constexpr int kUpdCo = 16, kUpdCi = 128, kUpdRow = kUpdCi * 9 + 2;   // (+2: the wd pass reads down the rows)

__global__ void __launch_bounds__(256) train_step_update_kernel(const StepChunk* __restrict__ chunks, float lr_scaled,
                                                                float stats_scale) {
  __shared__ __nv_bfloat16 tile[kUpdCo * kUpdRow];
  const StepChunk c = chunks[blockIdx.x];
  if (c.mode == 1) {
    for (int i = threadIdx.x; i < c.n; i += blockDim.x) c.p[i] *= stats_scale;
    return;
  }
  if (c.mode == 0) {
    for (int i = threadIdx.x; i < c.n; i += blockDim.x) c.p[i] -= lr_scaled * c.g[i];
    return;
  }
  if (c.mode == 3) {
    const int co0 = c.off, nco = c.n, ci0 = c.pad0, nci = c.pad1, run = nci * 9;
    for (int i = threadIdx.x; i < nco * run; i += blockDim.x) {
      const int co = i / run, r = i - co * run;
      const size_t e = ((size_t)(co0 + co) * c.Cin + ci0) * 9 + r;
      const float v = c.p[e] - lr_scaled * c.g[e];
      c.p[e] = v;
      tile[co * kUpdRow + r] = __float2bfloat16_rn(v);
    }
    __syncthreads();
    // forward operand wf[Cout][9][CinPad]: runs over the input channels
    for (int i = threadIdx.x; i < nco * run; i += blockDim.x) {
      const int ci = i % nci, t = i / nci, tap = t % 9, co = t / 9;
      c.wf[((size_t)(co0 + co) * 9 + tap) * c.CinPad + ci0 + ci] = tile[co * kUpdRow + ci * 9 + tap];
    }
    // data-gradient operand wd[Cin][9][Cout], taps flipped: runs over the output channels
    for (int i = threadIdx.x; i < nco * run; i += blockDim.x) {
      const int co = i % nco, t = i / nco, tap = t % 9, ci = t / 9;
      c.wd[((size_t)(ci0 + ci) * 9 + (8 - tap)) * c.Cout + co0 + co] = tile[co * kUpdRow + ci * 9 + tap];
    }
    return;
  }
  const int k9 = c.Cin * 9;
  for (int i = threadIdx.x; i < c.n; i += blockDim.x) {
    const float v = c.p[i] - lr_scaled * c.g[i];
    c.p[i] = v;
    const int e = c.off + i;
    const int co = e / k9, r = e - co * k9;
    const int ci = r / 9, tap = r - ci * 9;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    c.wf[((size_t)co * 9 + tap) * c.CinPad + ci] = h;
    c.wd[((size_t)ci * 9 + (8 - tap)) * c.Cout + co] = h;
  }
}
}  // namespace

extern "C" int fabric_b200_train_step_update(const void* chunks, int n_chunks, float lr, float grad_scale, float stats_scale,
                                             void* stream) {
  DeviceInfo di;
  int rc = device_info(&di);
  if (rc) return rc;
  if (!chunks || n_chunks < 1) return fail(FB_ERR_ARG, "no chunks");
  train_step_update_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const StepChunk*>(chunks),
                                                                       lr * grad_scale, stats_scale);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}
