// 3-D (volumetric) variants of the VED conv-net layers (reference nets/conv.py with ndim = 3:
// nn.Conv3d k = 1 | 3 stride 1 "same" padding, nn.MaxPool3d(2, 2), nearest x2 up-sampling --
// 'bilinear' is switched to 'nearest' for 3-D data, conv.py:127-130).  NCDHW fp32.
// No benchmarked configuration is volumetric, so these are direct kernels (one thread per output
// element, coalesced along W, weights broadcast within a warp) rather than tiled implicit GEMMs.
#include "pvb_common.cuh"

namespace {

struct Dims3 {
  int B, Cin, Cout, D, H, W, k;   // cubic kernel k = 1 | 3, padding k/2
};

constexpr int NT = 256;

// MODE 0: y[b,co,p]  = act(bias[co] + sum_{ci,t} x[b,ci,p+d_t] W[co,ci,t])
// MODE 1: dx[b,ci,p] = sum_{co,t} dpre[b,co,p-d_t] W[co,ci,t]
template <int MODE>
__global__ void __launch_bounds__(NT)
conv3d_direct_kernel(const float* __restrict__ src, const float* __restrict__ Wt,
                     const float* __restrict__ bias, float* __restrict__ dst, float* __restrict__ pre,
                     Dims3 d, int act) {
  const int Cg = MODE == 0 ? d.Cin : d.Cout, Nout = MODE == 0 ? d.Cout : d.Cin;
  const int HW = d.H * d.W, DHW = d.D * HW;
  const int taps = d.k * d.k * d.k, pad = d.k / 2;
  const int64_t total = (int64_t)d.B * Nout * DHW;
  const int64_t idx = (int64_t)blockIdx.x * NT + threadIdx.x;
  if (idx >= total) return;
  const int pos = (int)(idx % DHW);
  const int n = (int)((idx / DHW) % Nout);
  const int b = (int)(idx / ((int64_t)DHW * Nout));
  const int z = pos / HW, y = (pos - z * HW) / d.W, x = pos - z * HW - y * d.W;
  const float* sb = src + (int64_t)b * Cg * DHW;
  float s = (MODE == 0 && bias) ? __ldg(bias + n) : 0.f;
  for (int t = 0; t < taps; ++t) {
    int dz = t / (d.k * d.k) - pad, dy = (t / d.k) % d.k - pad, dx = t % d.k - pad;
    if (MODE == 1) { dz = -dz; dy = -dy; dx = -dx; }
    const int zz = z + dz, yy = y + dy, xx = x + dx;
    if (zz < 0 || zz >= d.D || yy < 0 || yy >= d.H || xx < 0 || xx >= d.W) continue;
    const float* sp = sb + zz * HW + yy * d.W + xx;
    for (int c = 0; c < Cg; ++c) {
      const float w = MODE == 0 ? __ldg(Wt + ((int64_t)n * d.Cin + c) * taps + t)
                                : __ldg(Wt + ((int64_t)c * d.Cin + n) * taps + t);
      s = fmaf(__ldg(sp + (int64_t)c * DHW), w, s);
    }
  }
  if (MODE == 0) {
    if (pre) pre[idx] = s;
    s = pvb::act_fwd(s, act);
  }
  dst[idx] = s;
}

// dW[co,ci,t] += sum_{b,p} dpre[b,co,p] x[b,ci,p+d_t];  db[co] += sum dpre   (CTA per (co, ci))
__global__ void __launch_bounds__(NT)
conv3d_wgrad_kernel(const float* __restrict__ dpre, const float* __restrict__ x, float* __restrict__ dW,
                    float* __restrict__ db, Dims3 d) {
  __shared__ float red[NT / 32][28];
  const int co = blockIdx.x / d.Cin, ci = blockIdx.x - co * d.Cin;
  const int HW = d.H * d.W, DHW = d.D * HW;
  const int taps = d.k * d.k * d.k, pad = d.k / 2;
  float acc[28];
#pragma unroll
  for (int t = 0; t < 28; ++t) acc[t] = 0.f;
  const int64_t total = (int64_t)d.B * DHW;
  for (int64_t e = threadIdx.x; e < total; e += NT) {
    const int b = (int)(e / DHW), pos = (int)(e - (int64_t)b * DHW);
    const float g = __ldg(dpre + ((int64_t)b * d.Cout + co) * DHW + pos);
    acc[27] += g;
    const int z = pos / HW, y = (pos - z * HW) / d.W, xw = pos - z * HW - y * d.W;
    const float* xb = x + ((int64_t)b * d.Cin + ci) * DHW;
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      if (t < taps) {
        const int zz = z + t / (d.k * d.k) - pad, yy = y + (t / d.k) % d.k - pad, xx = xw + t % d.k - pad;
        if (zz >= 0 && zz < d.D && yy >= 0 && yy < d.H && xx >= 0 && xx < d.W)
          acc[t] = fmaf(g, __ldg(xb + zz * HW + yy * d.W + xx), acc[t]);
      }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int t = 0; t < 28; ++t) {
    const float v = pvb::warp_sum(acc[t]);
    if (lane == 0) red[warp][t] = v;
  }
  __syncthreads();
  if (threadIdx.x < 28) {
    float s = 0.f;
    for (int w = 0; w < NT / 32; ++w) s += red[w][threadIdx.x];
    if (threadIdx.x < taps) dW[((int64_t)co * d.Cin + ci) * taps + threadIdx.x] += s;
    else if (threadIdx.x == 27 && ci == 0 && db) db[co] += s;
  }
}

// nn.MaxPool3d(2, 2): one thread per output window
__global__ void maxpool3d_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t BC, int D,
                                     int H, int W) {
  const int Do = D / 2, Ho = H / 2, Wo = W / 2;
  const int64_t total = BC * Do * Ho * Wo;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int wo = (int)(i % Wo), ho = (int)((i / Wo) % Ho), dd = (int)((i / ((int64_t)Wo * Ho)) % Do);
  const int64_t bc = i / ((int64_t)Wo * Ho * Do);
  const float* p = x + ((bc * D + 2 * dd) * H + 2 * ho) * W + 2 * wo;
  float m = -INFINITY;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float v = p[((k >> 2) * H + ((k >> 1) & 1)) * W + (k & 1)];
    m = (v > m || v != v) ? v : m;
  }
  y[i] = m;
}

// gradient goes to the first maximum of the window in (d, h, w) scan order, as torch's argmax
// does; dx must be zero-filled beforehand when a dimension is odd (uncovered border)
__global__ void maxpool3d_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                     float* __restrict__ dx, int64_t BC, int D, int H, int W) {
  const int Do = D / 2, Ho = H / 2, Wo = W / 2;
  const int64_t total = BC * Do * Ho * Wo;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int wo = (int)(i % Wo), ho = (int)((i / Wo) % Ho), dd = (int)((i / ((int64_t)Wo * Ho)) % Do);
  const int64_t bc = i / ((int64_t)Wo * Ho * Do);
  const int64_t base = ((bc * D + 2 * dd) * H + 2 * ho) * W + 2 * wo;
  float m = -INFINITY;
  int arg = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float v = x[base + ((k >> 2) * H + ((k >> 1) & 1)) * W + (k & 1)];
    if (v > m || v != v) { m = v; arg = k; }
  }
  const float g = dy[i];
#pragma unroll
  for (int k = 0; k < 8; ++k)
    dx[base + ((k >> 2) * H + ((k >> 1) & 1)) * W + (k & 1)] = k == arg ? g : 0.f;
}

// nearest x2: y[bc, d, h, w] = x[bc, d/2, h/2, w/2];  backward sums the 8 children
__global__ void upsample3d_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t BC, int D,
                                      int H, int W) {
  const int64_t total = BC * 8 * D * H * W;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int w = (int)(i % (2 * W)), h = (int)((i / (2 * W)) % (2 * H));
  const int dd = (int)((i / ((int64_t)4 * W * H)) % (2 * D));
  const int64_t bc = i / ((int64_t)8 * W * H * D);
  y[i] = x[((bc * D + dd / 2) * H + h / 2) * W + w / 2];
}
__global__ void upsample3d_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int64_t BC,
                                      int D, int H, int W) {
  const int64_t total = BC * D * H * W;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int w = (int)(i % W), h = (int)((i / W) % H), dd = (int)((i / ((int64_t)W * H)) % D);
  const int64_t bc = i / ((int64_t)W * H * D);
  const float* p = dy + ((bc * 2 * D + 2 * dd) * 2 * H + 2 * h) * 2 * W + 2 * w;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += p[((k >> 2) * 2 * H + ((k >> 1) & 1)) * 2 * W + (k & 1)];
  dx[i] = s;
}

inline bool dims_ok(int B, int Cin, int Cout, int D, int H, int W, int k) {
  return B >= 0 && Cin > 0 && Cout > 0 && D > 0 && H > 0 && W > 0 && (k == 1 || k == 3) &&
         (int64_t)D * H * W < (1ll << 31);
}

}  // namespace

extern "C" int pvb_conv3d_fwd(const float* x, const float* W, const float* b, float* y, float* pre, int B,
                              int Cin, int Cout, int D, int H, int Wd, int k, int act, void* stream) {
  PVB_CHECK_ARG(x && W && y, "pvb_conv3d_fwd: null pointer");
  PVB_CHECK_ARG(dims_ok(B, Cin, Cout, D, H, Wd, k), "pvb_conv3d_fwd: bad shape (k must be 1 or 3)");
  PVB_CHECK_ARG(act != PVB_ACT_GELU || pre, "pvb_conv3d_fwd: gelu needs the pre-activation buffer");
  const int64_t total = (int64_t)B * Cout * D * H * Wd;
  if (total == 0) return 0;
  Dims3 d{B, Cin, Cout, D, H, Wd, k};
  conv3d_direct_kernel<0><<<pvb::cdiv(total, NT), NT, 0, (cudaStream_t)stream>>>(x, W, b, y, pre, d, act);
  pvb::count_launch();
  return pvb::launch_status();
}

extern "C" int pvb_conv3d_bwd_data(const float* dpre, const float* W, float* dx, int B, int Cin, int Cout,
                                   int D, int H, int Wd, int k, void* stream) {
  PVB_CHECK_ARG(dpre && W && dx, "pvb_conv3d_bwd_data: null pointer");
  PVB_CHECK_ARG(dims_ok(B, Cin, Cout, D, H, Wd, k), "pvb_conv3d_bwd_data: bad shape");
  const int64_t total = (int64_t)B * Cin * D * H * Wd;
  if (total == 0) return 0;
  Dims3 d{B, Cin, Cout, D, H, Wd, k};
  conv3d_direct_kernel<1><<<pvb::cdiv(total, NT), NT, 0, (cudaStream_t)stream>>>(dpre, W, nullptr, dx,
                                                                                 nullptr, d, 0);
  pvb::count_launch();
  return pvb::launch_status();
}

extern "C" int pvb_conv3d_bwd_weight(const float* dpre, const float* x, float* dW, float* db, int B,
                                     int Cin, int Cout, int D, int H, int Wd, int k, void* stream) {
  PVB_CHECK_ARG(dpre && x && dW, "pvb_conv3d_bwd_weight: null pointer");
  PVB_CHECK_ARG(dims_ok(B, Cin, Cout, D, H, Wd, k), "pvb_conv3d_bwd_weight: bad shape");
  if (B == 0) return 0;
  Dims3 d{B, Cin, Cout, D, H, Wd, k};
  conv3d_wgrad_kernel<<<Cin * Cout, NT, 0, (cudaStream_t)stream>>>(dpre, x, dW, db, d);
  pvb::count_launch();
  return pvb::launch_status();
}

extern "C" int pvb_maxpool3d_fwd(const float* x, float* y, int64_t BC, int D, int H, int Wd, void* stream) {
  PVB_CHECK_ARG(x && y && BC >= 0 && D > 1 && H > 1 && Wd > 1, "pvb_maxpool3d_fwd: bad argument");
  const int64_t n = BC * (D / 2) * (H / 2) * (Wd / 2);
  if (n == 0) return 0;
  maxpool3d_fwd_kernel<<<pvb::cdiv(n, NT), NT, 0, (cudaStream_t)stream>>>(x, y, BC, D, H, Wd);
  pvb::count_launch();
  return pvb::launch_status();
}

extern "C" int pvb_maxpool3d_bwd(const float* x, const float* dy, float* dx, int64_t BC, int D, int H,
                                 int Wd, void* stream) {
  PVB_CHECK_ARG(x && dy && dx && BC >= 0 && D > 1 && H > 1 && Wd > 1, "pvb_maxpool3d_bwd: bad argument");
  const int64_t n = BC * (D / 2) * (H / 2) * (Wd / 2);
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if ((D | H | Wd) & 1) {   // odd sizes leave a border that no window covers
    cudaError_t e = cudaMemsetAsync(dx, 0, sizeof(float) * BC * D * H * Wd, st);
    if (e != cudaSuccess) return (int)e;
  }
  maxpool3d_bwd_kernel<<<pvb::cdiv(n, NT), NT, 0, st>>>(x, dy, dx, BC, D, H, Wd);
  pvb::count_launch();
  return pvb::launch_status();
}

extern "C" int pvb_upsample3d_fwd(const float* x, float* y, int64_t BC, int D, int H, int Wd, void* stream) {
  PVB_CHECK_ARG(x && y && BC >= 0 && D > 0 && H > 0 && Wd > 0, "pvb_upsample3d_fwd: bad argument");
  const int64_t n = BC * 8 * D * H * Wd;
  if (n == 0) return 0;
  upsample3d_fwd_kernel<<<pvb::cdiv(n, NT), NT, 0, (cudaStream_t)stream>>>(x, y, BC, D, H, Wd);
  pvb::count_launch();
  return pvb::launch_status();
}

extern "C" int pvb_upsample3d_bwd(const float* dy, float* dx, int64_t BC, int D, int H, int Wd,
                                  void* stream) {
  PVB_CHECK_ARG(dy && dx && BC >= 0 && D > 0 && H > 0 && Wd > 0, "pvb_upsample3d_bwd: bad argument");
  const int64_t n = BC * D * H * Wd;
  if (n == 0) return 0;
  upsample3d_bwd_kernel<<<pvb::cdiv(n, NT), NT, 0, (cudaStream_t)stream>>>(dy, dx, BC, D, H, Wd);
  pvb::count_launch();
  return pvb::launch_status();
}
