// Convolutional layers of the VED encoder / decoder (reference nets/conv.py:146-249):
// k x k (k = 1 or 3, stride 1, "same" zero padding) convolutions in 1-D / 2-D as implicit
// GEMMs with on-the-fly gathers (no im2col buffer), 2x max-pooling, 2x nearest / bilinear
// up-sampling, each with its hand-written backward.  NCHW fp32 like the reference; a 1-D
// signal [B, C, L] is the 2-D case H = 1 (kernel taps along W only).
//
//   forward        y[b,co,h,w] = act(bias[co] + sum_{ci,kh,kw} x[b,ci,h+kh-p,w+kw-p] W[co,ci,kh,kw])
//   backward data  dx[b,ci,h,w] = sum_{co,kh,kw} dpre[b,co,h-kh+p,w-kw+p] W[co,ci,kh,kw]
//   backward weight dW[co,ci,kh,kw] += sum_{b,h,w} dpre[b,co,h,w] x[b,ci,h+kh-p,w+kw-p];  db[co] += sum dpre
//
// GEMM view: 64 x 64 x 16 tiles, 256 threads, 4 x 4 outputs per thread (same inner loop as
// pvb_gemm.cu).  A per-CTA look-up table maps the flattened (channel, tap) index to its
// address offset and tap displacement, so a gather costs two adds and a bounds test.
#include "pvb_common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;
constexpr int MAX_CK = 2304;   // max (channels * taps) on the gathered side: 256 channels x 9

struct ConvDims {
  int B, Cin, Cout, H, W, kh, kw;   // kh = 1 for 1-D signals; pads = kh/2, kw/2
};

struct Tap { int off; short dh, dw; };

// element (channel c, tap t) of the gathered tensor with C channels: offset c*H*W + dh*W + dw
__device__ __forceinline__ void build_lut(Tap* lut, int n_ck, const ConvDims& d, int sign) {
  const int taps = d.kh * d.kw, ph = d.kh / 2, pw = d.kw / 2;
  for (int k = threadIdx.x; k < n_ck; k += blockDim.x) {
    int c = k / taps, t = k - c * taps;
    int dh = sign * (t / d.kw - ph), dw = sign * (t % d.kw - pw);
    lut[k].off = c * d.H * d.W + dh * d.W + dw;
    lut[k].dh = (short)dh;
    lut[k].dw = (short)dw;
  }
}

__device__ __forceinline__ void mma_tile(const float (*As)[BM + 4], const float (*Bs)[BN + 4], int tx,
                                         int ty, float acc[4][4]) {
#pragma unroll
  for (int k = 0; k < BK; ++k) {
    float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
    float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
    const float a[4] = {a4.x, a4.y, a4.z, a4.w};
    const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

// MODE 0: forward.   M = B*H*W pixels, N = Cout, K = Cin*taps; A gathered from x, B = W[n][k]
// MODE 1: backward data. M = pixels, N = Cin, K = Cout*taps; A gathered from dpre with mirrored
//         taps, B[k = (co,t)][n = ci] = W[co][ci][t]
template <int MODE>
__global__ void __launch_bounds__(NT)
conv_pix_kernel(const float* __restrict__ src, const float* __restrict__ Wt,
                const float* __restrict__ bias, float* __restrict__ dst, float* __restrict__ pre,
                ConvDims d, int act) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Tap* lut = reinterpret_cast<Tap*>(smem_raw);
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int taps = d.kh * d.kw;
  const int Cg = (MODE == 0) ? d.Cin : d.Cout;      // channels of the gathered tensor
  const int Nn = (MODE == 0) ? d.Cout : d.Cin;      // output channels of this GEMM
  const int K = Cg * taps;
  const int HW = d.H * d.W;
  const int64_t Mtot = (int64_t)d.B * HW;
  build_lut(lut, K, d, MODE == 0 ? 1 : -1);
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  // this thread's gather pixel (A loader: m fixed, k = tid/64 + 4*it)
  const int64_t gm = m0 + (tid & 63);
  const bool m_ok = gm < Mtot;
  const int gb = m_ok ? (int)(gm / HW) : 0;
  const int gr = m_ok ? (int)(gm - (int64_t)gb * HW) : 0;
  const int gh = gr / d.W, gw = gr - gh * d.W;
  const float* gbase = src + (int64_t)gb * Cg * HW + gr;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  __syncthreads();
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      int kl = (tid >> 6) + 4 * it;
      int k = k0 + kl;
      float v = 0.f;
      if (m_ok && k < K) {
        Tap t = lut[k];
        int hh = gh + t.dh, ww = gw + t.dw;
        if (hh >= 0 && hh < d.H && ww >= 0 && ww < d.W) v = __ldg(gbase + t.off);
      }
      As[kl][tid & 63] = v;
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      int idx = tid + it * NT;
      int kl = idx & 15, nl = idx >> 4;
      int k = k0 + kl, n = n0 + nl;
      float v = 0.f;
      if (k < K && n < Nn) {
        if (MODE == 0) {
          v = __ldg(Wt + (int64_t)n * K + k);
        } else {
          int co = k / taps, t = k - co * taps;
          v = __ldg(Wt + ((int64_t)co * d.Cin + n) * taps + t);
        }
      }
      Bs[kl][nl] = v;
    }
    __syncthreads();
    mma_tile(As, Bs, tx, ty, acc);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= Mtot) continue;
    int b = (int)(m / HW);
    int r = (int)(m - (int64_t)b * HW);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= Nn) continue;
      float v = acc[i][j];
      int64_t o = ((int64_t)b * Nn + n) * HW + r;
      if (MODE == 0) {
        if (bias) v += __ldg(bias + n);
        if (pre) pre[o] = v;
        v = pvb::act_fwd(v, act);
      }
      dst[o] = v;
    }
  }
}

// backward weight: dW[co][(ci,t)] += sum over pixels; M' = Cout, N' = Cin*taps, K' = B*H*W,
// split over gridDim.z pixel ranges (atomic accumulation); db from the n-tile-0 CTAs.
__global__ void __launch_bounds__(NT)
conv_wgrad_kernel(const float* __restrict__ dpre, const float* __restrict__ x, float* __restrict__ dW,
                  float* __restrict__ db, ConvDims d, int64_t pix_per_split) {
  __shared__ Tap lut[BN];
  __shared__ float As[BK][BM + 4];   // [pixel][co]
  __shared__ float Bs[BK][BN + 4];   // [pixel][(ci,t)]
  const int taps = d.kh * d.kw, ph = d.kh / 2, pw = d.kw / 2;
  const int Np = d.Cin * taps;
  const int HW = d.H * d.W;
  const int64_t Mtot = (int64_t)d.B * HW;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * BN, co0 = blockIdx.y * BM;
  if (tid < BN) {
    int n = n0 + tid;
    Tap t;
    t.off = 0; t.dh = 0; t.dw = 0;
    if (n < Np) {
      int c = n / taps, tt = n - c * taps;
      int dh = tt / d.kw - ph, dw = tt % d.kw - pw;
      t.off = c * HW + dh * d.W + dw;
      t.dh = (short)dh;
      t.dw = (short)dw;
    }
    lut[tid] = t;
  }
  __syncthreads();
  const int64_t p_begin = (int64_t)blockIdx.z * pix_per_split;
  const int64_t p_end = (p_begin + pix_per_split < Mtot) ? p_begin + pix_per_split : Mtot;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;
  const bool do_bias = db && blockIdx.x == 0;
  const int kl = tid & 15;        // this thread's pixel within the chunk (fixed)
  for (int64_t p0 = p_begin; p0 < p_end; p0 += BK) {
    const int64_t pm = p0 + kl;
    const bool p_ok = pm < p_end;
    const int b = p_ok ? (int)(pm / HW) : 0;
    const int r = p_ok ? (int)(pm - (int64_t)b * HW) : 0;
    const int h = r / d.W, w = r - h * d.W;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      int ml = (tid >> 4) + 16 * it;            // co within tile
      int co = co0 + ml;
      float v = 0.f;
      if (p_ok && co < d.Cout) v = __ldg(dpre + ((int64_t)b * d.Cout + co) * HW + r);
      As[kl][ml] = v;
    }
    const float* xb = x + (int64_t)b * d.Cin * HW + r;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      int nl = (tid >> 4) + 16 * it;
      float v = 0.f;
      if (p_ok && n0 + nl < Np) {
        Tap t = lut[nl];
        int hh = h + t.dh, ww = w + t.dw;
        if (hh >= 0 && hh < d.H && ww >= 0 && ww < d.W) v = __ldg(xb + t.off);
      }
      Bs[kl][nl] = v;
    }
    __syncthreads();
    mma_tile(As, Bs, tx, ty, acc);
    if (do_bias && tid < BM) {
#pragma unroll
      for (int k = 0; k < BK; ++k) bsum += As[k][tid];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int co = co0 + ty * 4 + i;
    if (co >= d.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < Np) atomicAdd(dW + (int64_t)co * Np + n, acc[i][j]);
    }
  }
  if (do_bias && tid < BM && co0 + tid < d.Cout) atomicAdd(db + co0 + tid, bsum);
}

// ---- 2x max-pool (nn.MaxPool{1,2}d(2, 2), floor mode) ------------------------------------------
__global__ void maxpool2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t BC,
                                    int H, int W, int Ho, int Wo, int two_d) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BC * Ho * Wo) return;
  int wo = (int)(i % Wo);
  int ho = (int)((i / Wo) % Ho);
  int64_t bc = i / ((int64_t)Wo * Ho);
  const float* p = x + (bc * H + (two_d ? 2 * ho : ho)) * W + 2 * wo;
  float m = fmaxf(p[0], p[1]);
  if (two_d) m = fmaxf(m, fmaxf(p[W], p[W + 1]));
  y[i] = m;
}
// 2-D, W % 8 == 0, even H, 16-byte aligned planes: one thread per FOUR horizontally adjacent windows --
// two 32-byte row segments in (four 16-byte loads), one 16-byte store; a warp reads 2 x 1 KB of
// contiguous bytes.  HBM-bound: 5 B per input element.
__global__ void __launch_bounds__(256)
maxpool2_fwd_vec_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n_quads, int W,
                        int Wo4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_quads) return;
  const int64_t rowo = i / Wo4;                  // output row index over all planes (b, c, ho)
  const int q = (int)(i - rowo * Wo4);
  const float4* p0 = reinterpret_cast<const float4*>(x + (2 * rowo) * W + 8 * q);
  const float4* p1 = reinterpret_cast<const float4*>(x + (2 * rowo + 1) * W + 8 * q);
  const float4 a0 = __ldg(p0), a1 = __ldg(p0 + 1), b0 = __ldg(p1), b1 = __ldg(p1 + 1);
  float4 o;
  o.x = fmaxf(fmaxf(a0.x, a0.y), fmaxf(b0.x, b0.y));
  o.y = fmaxf(fmaxf(a0.z, a0.w), fmaxf(b0.z, b0.w));
  o.z = fmaxf(fmaxf(a1.x, a1.y), fmaxf(b1.x, b1.y));
  o.w = fmaxf(fmaxf(a1.z, a1.w), fmaxf(b1.z, b1.w));
  reinterpret_cast<float4*>(y)[i] = o;
}
// backward of the same case: four windows per thread, 16-byte accesses throughout
__global__ void __launch_bounds__(256)
maxpool2_bwd_vec_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx,
                        int64_t n_quads, int W, int Wo4, int act) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_quads) return;
  const int64_t rowo = i / Wo4;
  const int q = (int)(i - rowo * Wo4);
  const int64_t o0 = (2 * rowo) * W + 8 * q, o1 = o0 + W;
  const float4 a0 = __ldg(reinterpret_cast<const float4*>(x + o0)),
               a1 = __ldg(reinterpret_cast<const float4*>(x + o0) + 1),
               b0 = __ldg(reinterpret_cast<const float4*>(x + o1)),
               b1 = __ldg(reinterpret_cast<const float4*>(x + o1) + 1);
  const float4 g4 = __ldg(reinterpret_cast<const float4*>(dy) + i);
  const float top[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  const float bot[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
  float dt[8], db[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    // first maximal element in row-major window order (torch's tie-breaking)
    const float v[4] = {top[2 * k], top[2 * k + 1], bot[2 * k], bot[2 * k + 1]};
    int best = 0;
    float vb = v[0];
#pragma unroll
    for (int e = 1; e < 4; ++e)
      if (v[e] > vb) { vb = v[e]; best = e; }
    const float g = gg[k] * (act ? pvb::act_grad(vb, 0.f, act) : 1.f);
    dt[2 * k] = best == 0 ? g : 0.f;
    dt[2 * k + 1] = best == 1 ? g : 0.f;
    db[2 * k] = best == 2 ? g : 0.f;
    db[2 * k + 1] = best == 3 ? g : 0.f;
  }
  reinterpret_cast<float4*>(dx + o0)[0] = make_float4(dt[0], dt[1], dt[2], dt[3]);
  reinterpret_cast<float4*>(dx + o0)[1] = make_float4(dt[4], dt[5], dt[6], dt[7]);
  reinterpret_cast<float4*>(dx + o1)[0] = make_float4(db[0], db[1], db[2], db[3]);
  reinterpret_cast<float4*>(dx + o1)[1] = make_float4(db[4], db[5], db[6], db[7]);
}
// dx = dy routed to the first maximal element of each window (torch tie-breaking: first in
// row-major window order).  One thread per output window (reads its 2 / 4 inputs once, writes the
// whole window); elements outside any window (odd sizes) are zeroed by the trailing threads.
__global__ void maxpool2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                    float* __restrict__ dx, int64_t BC, int H, int W, int Ho, int Wo,
                                    int two_d, int act) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n_win = BC * Ho * Wo;
  if (i < n_win) {
    int wo = (int)(i % Wo);
    int ho = (int)((i / Wo) % Ho);
    int64_t bc = i / ((int64_t)Wo * Ho);
    const int64_t o = (bc * H + (two_d ? 2 * ho : ho)) * W + 2 * wo;
    float v[4] = {x[o], x[o + 1], two_d ? x[o + W] : -INFINITY, two_d ? x[o + W + 1] : -INFINITY};
    int best = 0;
#pragma unroll
    for (int k = 1; k < 4; ++k)
      if (v[k] > v[best]) best = k;
    // act != none: x is the activation output feeding the pool; the gradient that reaches the
    // maximal element continues through that activation (dx = its dpre), saving a separate pass
    const float g = dy[i] * (act ? pvb::act_grad(v[best], 0.f, act) : 1.f);
    float2 top = make_float2(best == 0 ? g : 0.f, best == 1 ? g : 0.f);
    if ((o & 1) == 0) {
      *reinterpret_cast<float2*>(dx + o) = top;
    } else {
      dx[o] = top.x;
      dx[o + 1] = top.y;
    }
    if (two_d) {
      float2 bot = make_float2(best == 2 ? g : 0.f, best == 3 ? g : 0.f);
      if (((o + W) & 1) == 0) {
        *reinterpret_cast<float2*>(dx + o + W) = bot;
      } else {
        dx[o + W] = bot.x;
        dx[o + W + 1] = bot.y;
      }
    }
    return;
  }
  // odd sizes: the last column / row belongs to no window
  i -= n_win;
  const int odd_w = W & 1, odd_h = two_d ? (H & 1) : 0;
  const int64_t per_plane = (int64_t)odd_w * H + (int64_t)odd_h * (W - odd_w);
  if (per_plane == 0 || i >= BC * per_plane) return;
  int64_t bc = i / per_plane;
  int r = (int)(i - bc * per_plane);
  int h, w;
  if (r < odd_w * H) { h = r; w = W - 1; }
  else { h = H - 1; w = r - odd_w * H; }
  dx[(bc * H + h) * W + w] = 0.f;
}

// ---- 2x up-sampling (F.interpolate(scale_factor=2), nearest; bilinear for 2-D, align_corners=False)
__device__ __forceinline__ void bilin_src(int o, int n_in, int& i0, int& i1, float& l1) {
  float s = fmaxf(((float)o + 0.5f) * 0.5f - 0.5f, 0.f);   // area_pixel_compute_source_index
  i0 = (int)s;
  i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
  l1 = s - (float)i0;
}
__global__ void upsample2_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t BC,
                                     int H, int W, int two_d, int bilinear) {
  const int Ho = two_d ? 2 * H : H, Wo = 2 * W;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BC * Ho * Wo) return;
  int wo = (int)(i % Wo);
  int ho = (int)((i / Wo) % Ho);
  int64_t bc = i / ((int64_t)Wo * Ho);
  const float* p = x + bc * H * W;
  if (!bilinear) {
    y[i] = p[(two_d ? ho >> 1 : ho) * W + (wo >> 1)];
    return;
  }
  int h0, h1, w0, w1;
  float lh, lw;
  bilin_src(ho, H, h0, h1, lh);
  bilin_src(wo, W, w0, w1, lw);
  y[i] = (1.f - lh) * ((1.f - lw) * p[h0 * W + w0] + lw * p[h0 * W + w1]) +
         lh * ((1.f - lw) * p[h1 * W + w0] + lw * p[h1 * W + w1]);
}
// gather form of the adjoint: each input element collects from the <= 3 x 3 outputs that read it
// y_below (optional) = output of the layer below with activation `act`: dx is multiplied by act'(y_below), i.e.
// it comes out as that layer's dpre (saves a separate pass over dx)
__global__ void upsample2_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int64_t BC,
                                     int H, int W, int two_d, int bilinear,
                                     const float* __restrict__ y_below, int act) {
  const int Ho = two_d ? 2 * H : H, Wo = 2 * W;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BC * H * W) return;
  int w = (int)(i % W);
  int h = (int)((i / W) % H);
  int64_t bc = i / ((int64_t)W * H);
  const float* g = dy + bc * Ho * Wo;
  float s = 0.f;
  if (!bilinear) {
    if (two_d) s = g[(2 * h) * Wo + 2 * w] + g[(2 * h) * Wo + 2 * w + 1] +
                   g[(2 * h + 1) * Wo + 2 * w] + g[(2 * h + 1) * Wo + 2 * w + 1];
    else s = g[h * Wo + 2 * w] + g[h * Wo + 2 * w + 1];
    dx[i] = y_below ? s * pvb::act_grad(y_below[i], 0.f, act) : s;
    return;
  }
  for (int ho = max(2 * h - 2, 0); ho <= min(2 * h + 2, Ho - 1); ++ho) {
    int h0, h1;
    float lh;
    bilin_src(ho, H, h0, h1, lh);
    float ch = (h0 == h ? 1.f - lh : 0.f) + (h1 == h ? lh : 0.f);
    if (ch == 0.f) continue;
    for (int wo = max(2 * w - 2, 0); wo <= min(2 * w + 2, Wo - 1); ++wo) {
      int w0, w1;
      float lw;
      bilin_src(wo, W, w0, w1, lw);
      float cw = (w0 == w ? 1.f - lw : 0.f) + (w1 == w ? lw : 0.f);
      if (cw != 0.f) s = fmaf(ch * cw, g[ho * Wo + wo], s);
    }
  }
  dx[i] = y_below ? s * pvb::act_grad(y_below[i], 0.f, act) : s;
}

// dy and dpre may be the SAME buffer (the engine applies the derivative in place): no __restrict__ on the
// pair and plain loads of dy (a thread reads its element before it writes it)
__global__ void act_bwd_flat_kernel(const float* dy, const float* __restrict__ y,
                                    const float* __restrict__ pre, float* dpre, int64_t n, int act) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // 16-byte accesses over the aligned body, scalar tail
  const bool vec = ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(y) |
                     reinterpret_cast<uintptr_t>(dpre) | (pre ? reinterpret_cast<uintptr_t>(pre) : 0)) & 15) == 0;
  const int64_t n4 = vec ? n / 4 : 0;
  for (int64_t k = i; k < n4; k += stride) {
    const float4 g = reinterpret_cast<const float4*>(dy)[k];
    const float4 v = __ldg(reinterpret_cast<const float4*>(y) + k);
    const float4 p = pre ? __ldg(reinterpret_cast<const float4*>(pre) + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    reinterpret_cast<float4*>(dpre)[k] =
        make_float4(g.x * pvb::act_grad(v.x, p.x, act), g.y * pvb::act_grad(v.y, p.y, act),
                    g.z * pvb::act_grad(v.z, p.z, act), g.w * pvb::act_grad(v.w, p.w, act));
  }
  for (int64_t k = 4 * n4 + i; k < n; k += stride) dpre[k] = dy[k] * pvb::act_grad(y[k], pre ? pre[k] : 0.f, act);
}

// Cout == 1, 1 x 1 kernel (the last decoder layer, nets/conv.py:141: hidden -> one output channel): a weighted
// sum of Cin planes -- HBM-bound, one pass over x; four pixels per thread, 16-byte accesses (HW % 4 == 0).
// The generic pixel-GEMM kernels spent 24 + 13 + 40 us on it at batch 512 x 32 x 128 (8 MB).
__global__ void __launch_bounds__(256)
conv_o1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ b,
                   float* __restrict__ y, float* __restrict__ pre, int64_t n4, int Cin, int HW4, int act) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  const int64_t bi = i / HW4;
  const int p4 = (int)(i - bi * HW4);
  const float4* xp = reinterpret_cast<const float4*>(x) + bi * Cin * HW4 + p4;
  const float b0 = b ? __ldg(b) : 0.f;
  float4 acc = make_float4(b0, b0, b0, b0);
#pragma unroll 8
  for (int c = 0; c < Cin; ++c) {
    const float4 v = __ldg(xp + (int64_t)c * HW4);
    const float w = __ldg(W + c);
    acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
    acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
  }
  if (pre) reinterpret_cast<float4*>(pre)[i] = acc;
  reinterpret_cast<float4*>(y)[i] = make_float4(pvb::act_fwd(acc.x, act), pvb::act_fwd(acc.y, act),
                                                pvb::act_fwd(acc.z, act), pvb::act_fwd(acc.w, act));
}
// dx[b][c][p] = W[c] dpre[b][p]  (times act'(y_below[b][c][p]) when the layer below's output is given)
__global__ void __launch_bounds__(256)
conv_o1_bwd_data_kernel(const float* __restrict__ dpre, const float* __restrict__ W, float* __restrict__ dx,
                        int64_t n4, int Cin, int HW4, const float* __restrict__ y_below, int act) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;     // over B x Cin x HW4
  if (i >= n4) return;
  const int64_t bc = i / HW4;
  const int p4 = (int)(i - bc * HW4);
  const int64_t bi = bc / Cin;
  const int c = (int)(bc - bi * Cin);
  const float4 g = __ldg(reinterpret_cast<const float4*>(dpre) + bi * HW4 + p4);
  const float w = __ldg(W + c);
  float4 o = make_float4(w * g.x, w * g.y, w * g.z, w * g.w);
  if (y_below) {
    const float4 yb = __ldg(reinterpret_cast<const float4*>(y_below) + i);
    o.x *= pvb::act_grad(yb.x, 0.f, act); o.y *= pvb::act_grad(yb.y, 0.f, act);
    o.z *= pvb::act_grad(yb.z, 0.f, act); o.w *= pvb::act_grad(yb.w, 0.f, act);
  }
  reinterpret_cast<float4*>(dx)[i] = o;
}
// dW[c] += sum_{b,p} dpre[b][p] x[b][c][p];  db += sum dpre.   grid (Cin, batch splits)
__global__ void __launch_bounds__(256)
conv_o1_wgrad_kernel(const float* __restrict__ dpre, const float* __restrict__ x, float* __restrict__ dW,
                     float* __restrict__ db, int B, int Cin, int HW4, int rows_per_split) {
  __shared__ float red[32];
  const int c = blockIdx.x;
  const int b0 = (int)blockIdx.y * rows_per_split, b1 = min(B, b0 + rows_per_split);
  const int64_t items = (int64_t)(b1 - b0) * HW4;
  float acc = 0.f, accb = 0.f;
  for (int64_t i = threadIdx.x; i < items; i += 256) {
    const int64_t bi = b0 + i / HW4;
    const int p4 = (int)(i % HW4);
    const float4 g = __ldg(reinterpret_cast<const float4*>(dpre) + bi * HW4 + p4);
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + (bi * Cin + c) * HW4 + p4);
    acc = fmaf(g.x, v.x, fmaf(g.y, v.y, fmaf(g.z, v.z, fmaf(g.w, v.w, acc))));
    accb += (g.x + g.y) + (g.z + g.w);
  }
  acc = pvb::block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(dW + c, acc);
  if (db && c == 0) {
    accb = pvb::block_sum(accb, red);
    if (threadIdx.x == 0) atomicAdd(db, accb);
  }
}

// Cin == 1 (the first encoder layer): HBM-write bound, so no GEMM machinery -- one thread per
// pixel keeps its <= 9 inputs in registers and streams the Cout outputs (coalesced per channel)
__global__ void __launch_bounds__(256)
conv_c1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ Wt,
                   const float* __restrict__ bias, float* __restrict__ y, float* __restrict__ pre,
                   ConvDims d, int act) {
  extern __shared__ float wsm[];   // [Cout][taps] + [Cout]
  const int taps = d.kh * d.kw, ph = d.kh / 2, pw = d.kw / 2;
  for (int k = threadIdx.x; k < d.Cout * taps; k += blockDim.x) wsm[k] = Wt[k];
  for (int k = threadIdx.x; k < d.Cout; k += blockDim.x) wsm[d.Cout * taps + k] = bias ? bias[k] : 0.f;
  __syncthreads();
  const int HW = d.H * d.W;
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= (int64_t)d.B * HW) return;
  const int b = (int)(m / HW), r = (int)(m - (int64_t)b * HW);
  const int h = r / d.W, w = r - h * d.W;
  float in[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    in[t] = 0.f;
    if (t < taps) {
      int hh = h + t / d.kw - ph, ww = w + t % d.kw - pw;
      if (hh >= 0 && hh < d.H && ww >= 0 && ww < d.W) in[t] = __ldg(x + (int64_t)b * HW + hh * d.W + ww);
    }
  }
  float* o = y + (int64_t)b * d.Cout * HW + r;
  float* po = pre ? pre + (int64_t)b * d.Cout * HW + r : nullptr;
  for (int co = 0; co < d.Cout; ++co) {
    float v = wsm[d.Cout * taps + co];
    const float* wr = wsm + co * taps;
#pragma unroll
    for (int t = 0; t < 9; ++t)
      if (t < taps) v = fmaf(in[t], wr[t], v);
    if (po) po[(int64_t)co * HW] = v;
    o[(int64_t)co * HW] = pvb::act_fwd(v, act);
  }
}

// Same layer, four horizontally adjacent pixels per thread (W % 4 == 0, 16-byte aligned tensors): every
// weight fetched from shared memory feeds four FMAs and every output channel is one 16-byte store --
// the one-pixel kernel spends its time on 9 * Cout shared-memory loads and Cout 4-byte stores per pixel.
__global__ void __launch_bounds__(256)
conv_c1_fwd4_kernel(const float* __restrict__ x, const float* __restrict__ Wt,
                    const float* __restrict__ bias, float* __restrict__ y, float* __restrict__ pre,
                    ConvDims d, int act) {
  extern __shared__ float wsm[];   // [Cout][12]: 9 taps (row-major kh x kw, zero padded) + bias
  const int taps = d.kh * d.kw, ph = d.kh / 2, pw = d.kw / 2;
  for (int k = threadIdx.x; k < d.Cout * 12; k += blockDim.x) {
    const int co = k / 12, j = k - co * 12;
    float v = 0.f;
    if (j < 9) {                    // slot (r, c) of the 3 x 3 window <- tap (r - 1 + ph, c - 1 + pw)
      const int r = j / 3 - 1 + ph, c = j % 3 - 1 + pw;
      if (r >= 0 && r < d.kh && c >= 0 && c < d.kw) v = Wt[co * taps + r * d.kw + c];
    } else if (j == 9) {
      v = bias ? bias[co] : 0.f;
    }
    wsm[k] = v;
  }
  __syncthreads();
  const int HW = d.H * d.W, W4 = d.W / 4;
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // group of 4 pixels
  if (q >= (int64_t)d.B * d.H * W4) return;
  const int b = (int)(q / (d.H * W4)), rq = (int)(q - (int64_t)b * d.H * W4);
  const int h = rq / W4, w0 = (rq - h * W4) * 4;
  // 3 rows x 6 columns of inputs around the 4 pixels (zero outside the image)
  float in[3][6];
  const float* xb = x + (int64_t)b * HW;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int hh = h + r - 1;
    const bool rok = hh >= 0 && hh < d.H;
    const float4 mid = rok ? __ldg(reinterpret_cast<const float4*>(xb + hh * d.W + w0)) : make_float4(0.f, 0.f, 0.f, 0.f);
    in[r][0] = (rok && w0 > 0) ? __ldg(xb + hh * d.W + w0 - 1) : 0.f;
    in[r][1] = mid.x; in[r][2] = mid.y; in[r][3] = mid.z; in[r][4] = mid.w;
    in[r][5] = (rok && w0 + 4 < d.W) ? __ldg(xb + hh * d.W + w0 + 4) : 0.f;
  }
  const int64_t obase = (int64_t)b * d.Cout * HW + h * d.W + w0;
  for (int co = 0; co < d.Cout; ++co) {
    const float4 wa = *reinterpret_cast<const float4*>(wsm + co * 12);
    const float4 wb = *reinterpret_cast<const float4*>(wsm + co * 12 + 4);
    const float4 wc = *reinterpret_cast<const float4*>(wsm + co * 12 + 8);
    const float wk[9] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x};
    float v[4] = {wc.y, wc.y, wc.y, wc.y};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int p = 0; p < 4; ++p) v[p] = fmaf(in[r][p + c], wk[r * 3 + c], v[p]);
    if (pre) *reinterpret_cast<float4*>(pre + obase + (int64_t)co * HW) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(y + obase + (int64_t)co * HW) =
        make_float4(pvb::act_fwd(v[0], act), pvb::act_fwd(v[1], act), pvb::act_fwd(v[2], act),
                    pvb::act_fwd(v[3], act));
  }
}

int check_dims(const ConvDims& d, const char* who) {
  PVB_CHECK_ARG(d.B >= 0 && d.Cin > 0 && d.Cout > 0 && d.H > 0 && d.W > 0, "%s: bad dims", who);
  PVB_CHECK_ARG((d.kh == 1 || d.kh == 3) && (d.kw == 1 || d.kw == 3), "%s: kernel size must be 1 or 3", who);
  PVB_CHECK_ARG(d.Cin * d.kh * d.kw <= MAX_CK && d.Cout * d.kh * d.kw <= MAX_CK,
                "%s: channels * taps > %d", who, MAX_CK);
  return 0;
}

}  // namespace

// Weight gradient of a Cin == 1 layer (the first encoder layer): dW[co][t] = sum_px dpre[co][px]
// x[px + d_t] has Cout * taps outputs and reads Cout planes of dpre once -- HBM-bound, so no GEMM
// machinery: a thread keeps the <= 9 neighbours of its pixel in registers and accumulates the 8 output
// channels of its CTA's channel group; one block reduction + atomics per CTA at the end.
constexpr int C1_CO = 8;
__global__ void __launch_bounds__(256)
conv_c1_wgrad_kernel(const float* __restrict__ dpre, const float* __restrict__ x, float* __restrict__ dW,
                     float* __restrict__ db, ConvDims d) {
  __shared__ float red[8][C1_CO * 10];
  const int taps = d.kh * d.kw, ph = d.kh / 2, pw = d.kw / 2;
  const int HW = d.H * d.W;
  const int64_t Mtot = (int64_t)d.B * HW;
  const int co0 = blockIdx.y * C1_CO;
  float acc[C1_CO][10];
#pragma unroll
  for (int c = 0; c < C1_CO; ++c)
#pragma unroll
    for (int t = 0; t < 10; ++t) acc[c][t] = 0.f;
  // 32-bit pixel arithmetic (host checks B*H*W < 2^31); (b, r) advance incrementally
  const unsigned stride = gridDim.x * 256u;
  const unsigned sb = stride / (unsigned)HW, sr = stride - sb * (unsigned)HW;
  unsigned m = blockIdx.x * 256u + threadIdx.x;
  unsigned b = m / (unsigned)HW, r = m - b * (unsigned)HW;
  for (; m < (unsigned)Mtot; m += stride) {
    const int h = (int)(r / (unsigned)d.W), w = (int)r - h * d.W;
    const float* xp = x + (int64_t)b * HW + r;
    float in[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      in[t] = 0.f;
      if (t < taps) {
        const int dh = t / d.kw - ph, dw = t % d.kw - pw;
        if ((unsigned)(h + dh) < (unsigned)d.H && (unsigned)(w + dw) < (unsigned)d.W)
          in[t] = __ldg(xp + dh * d.W + dw);
      }
    }
    const float* gp = dpre + ((int64_t)b * d.Cout + co0) * HW + r;
    float g[C1_CO];
#pragma unroll
    for (int c = 0; c < C1_CO; ++c) g[c] = co0 + c < d.Cout ? __ldg(gp + (int64_t)c * HW) : 0.f;
#pragma unroll
    for (int c = 0; c < C1_CO; ++c) {
      acc[c][9] += g[c];
#pragma unroll
      for (int t = 0; t < 9; ++t) acc[c][t] = fmaf(g[c], in[t], acc[c][t]);
    }
    r += sr;
    b += sb;
    if (r >= (unsigned)HW) { r -= (unsigned)HW; ++b; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < C1_CO; ++c)
#pragma unroll
    for (int t = 0; t < 10; ++t) {
      const float v = pvb::warp_sum(acc[c][t]);
      if (lane == 0) red[warp][c * 10 + t] = v;
    }
  __syncthreads();
  if (threadIdx.x < C1_CO * 10) {
    float s = 0.f;
    for (int w8 = 0; w8 < 8; ++w8) s += red[w8][threadIdx.x];
    const int c = threadIdx.x / 10, t = threadIdx.x - c * 10;
    if (co0 + c < d.Cout) {
      if (t < taps) atomicAdd(dW + (int64_t)(co0 + c) * taps + t, s);
      else if (t == 9 && db) atomicAdd(db + co0 + c, s);
    }
  }
}

// Same gradient, four horizontally adjacent pixels per iteration (W % 4 == 0, 16-byte aligned): the
// Cout planes of dpre are read with 16-byte loads, the 3 x 6 input window once per group, and four output
// channels per CTA keep the accumulators (4 x 10) + window + gradients under 100 registers.
constexpr int C1_CO4 = 4;
__global__ void __launch_bounds__(256)
conv_c1_wgrad4_kernel(const float* __restrict__ dpre, const float* __restrict__ x, float* __restrict__ dW,
                      float* __restrict__ db, ConvDims d) {
  __shared__ float red[8][C1_CO4 * 10];
  const int taps = d.kh * d.kw, ph = d.kh / 2, pw = d.kw / 2;
  const int HW = d.H * d.W, W4 = d.W / 4;
  const unsigned G = (unsigned)d.B * (unsigned)d.H * (unsigned)W4;     // groups of 4 pixels
  const int co0 = blockIdx.y * C1_CO4;
  float acc[C1_CO4][10];
#pragma unroll
  for (int c = 0; c < C1_CO4; ++c)
#pragma unroll
    for (int t = 0; t < 10; ++t) acc[c][t] = 0.f;
  for (unsigned q = blockIdx.x * 256u + threadIdx.x; q < G; q += gridDim.x * 256u) {
    const unsigned b = q / (unsigned)(d.H * W4), rq = q - b * (unsigned)(d.H * W4);
    const int h = (int)(rq / (unsigned)W4), w0 = (int)(rq - (unsigned)h * (unsigned)W4) * 4;
    float in[3][6];
    const float* xb = x + (int64_t)b * HW;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int hh = h + r - 1;
      const bool rok = hh >= 0 && hh < d.H;
      const float4 mid = rok ? __ldg(reinterpret_cast<const float4*>(xb + hh * d.W + w0)) : make_float4(0.f, 0.f, 0.f, 0.f);
      in[r][0] = (rok && w0 > 0) ? __ldg(xb + hh * d.W + w0 - 1) : 0.f;
      in[r][1] = mid.x; in[r][2] = mid.y; in[r][3] = mid.z; in[r][4] = mid.w;
      in[r][5] = (rok && w0 + 4 < d.W) ? __ldg(xb + hh * d.W + w0 + 4) : 0.f;
    }
    const float* gp = dpre + ((int64_t)b * d.Cout + co0) * HW + h * d.W + w0;
    float4 g[C1_CO4];
#pragma unroll
    for (int c = 0; c < C1_CO4; ++c)
      g[c] = co0 + c < d.Cout ? __ldg(reinterpret_cast<const float4*>(gp + (int64_t)c * HW)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c = 0; c < C1_CO4; ++c) {
      const float gv[4] = {g[c].x, g[c].y, g[c].z, g[c].w};
      acc[c][9] += (gv[0] + gv[1]) + (gv[2] + gv[3]);
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
#pragma unroll
          for (int p = 0; p < 4; ++p) acc[c][r * 3 + cc] = fmaf(gv[p], in[r][p + cc], acc[c][r * 3 + cc]);
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int c = 0; c < C1_CO4; ++c)
#pragma unroll
    for (int t = 0; t < 10; ++t) {
      const float v = pvb::warp_sum(acc[c][t]);
      if (lane == 0) red[warp][c * 10 + t] = v;
    }
  __syncthreads();
  if (threadIdx.x < C1_CO4 * 10) {
    float s = 0.f;
    for (int w8 = 0; w8 < 8; ++w8) s += red[w8][threadIdx.x];
    const int c = threadIdx.x / 10, j = threadIdx.x - c * 10;
    if (co0 + c < d.Cout) {
      if (j == 9) {
        if (db) atomicAdd(db + co0 + c, s);
      } else {                       // window slot (r, cc) -> tap (r - 1 + ph, cc - 1 + pw)
        const int r = j / 3 - 1 + ph, cc = j % 3 - 1 + pw;
        if (r >= 0 && r < d.kh && cc >= 0 && cc < d.kw) atomicAdd(dW + (int64_t)(co0 + c) * taps + r * d.kw + cc, s);
      }
    }
  }
}

extern "C" int pvb_conv_fwd(const float* x, const float* W, const float* b, float* y, float* pre,
                            int B, int Cin, int Cout, int H, int Wd, int kh, int kw, int act,
                            void* stream) {
  ConvDims d{B, Cin, Cout, H, Wd, kh, kw};
  if (check_dims(d, "pvb_conv_fwd")) return -1;
  PVB_CHECK_ARG(x && W && y, "pvb_conv_fwd: null pointer");
  PVB_CHECK_ARG(act >= 0 && act <= PVB_ACT_SIGMOID, "pvb_conv_fwd: unknown activation %d", act);
  if (B == 0) return 0;
  int64_t M = (int64_t)B * H * Wd;
  if (Cout == 1 && kh * kw == 1 && (H * Wd) % 4 == 0 &&
      (((uintptr_t)x | (uintptr_t)y | (uintptr_t)pre) & 15) == 0) {
    conv_o1_fwd_kernel<<<pvb::cdiv(M / 4, 256), 256, 0, (cudaStream_t)stream>>>(x, W, b, y, pre, M / 4, Cin,
                                                                                H * Wd / 4, act);
    pvb::count_launch();
    return pvb::launch_status();
  }
  if (Cin == 1 && Wd % 4 == 0 && Cout * 12 * sizeof(float) <= 40 * 1024 &&
      (((uintptr_t)x | (uintptr_t)y | (uintptr_t)pre) & 15) == 0) {
    const int64_t groups = M / 4;
    conv_c1_fwd4_kernel<<<pvb::cdiv(groups, 256), 256, (size_t)Cout * 12 * sizeof(float), (cudaStream_t)stream>>>(
        x, W, b, y, pre, d, act);
    pvb::count_launch();
    return pvb::launch_status();
  }
  if (Cin == 1 && Cout * (kh * kw + 1) * sizeof(float) <= 40 * 1024) {
    size_t smem1 = (size_t)Cout * (kh * kw + 1) * sizeof(float);
    conv_c1_fwd_kernel<<<pvb::cdiv(M, 256), 256, smem1, (cudaStream_t)stream>>>(x, W, b, y, pre, d, act);
    pvb::count_launch();
    return pvb::launch_status();
  }
  dim3 grid((unsigned)((M + BM - 1) / BM), (Cout + BN - 1) / BN);
  size_t smem = (size_t)Cin * kh * kw * sizeof(Tap);
  conv_pix_kernel<0><<<grid, NT, smem, (cudaStream_t)stream>>>(x, W, b, y, pre, d, act);
  pvb::count_launch();
  return pvb::launch_status();
}

extern "C" int pvb_conv_bwd_data(const float* dpre, const float* W, float* dx, int B, int Cin,
                                 int Cout, int H, int Wd, int kh, int kw, const float* y_below, int act,
                                 void* stream) {
  ConvDims d{B, Cin, Cout, H, Wd, kh, kw};
  if (check_dims(d, "pvb_conv_bwd_data")) return -1;
  PVB_CHECK_ARG(dpre && W && dx, "pvb_conv_bwd_data: null pointer");
  PVB_CHECK_ARG(!y_below || (act >= 0 && act <= PVB_ACT_SIGMOID && act != PVB_ACT_GELU),
                "pvb_conv_bwd_data: fused activation derivative needs an activation expressed from its output");
  if (B == 0) return 0;
  int64_t M = (int64_t)B * H * Wd;
  if (Cout == 1 && kh * kw == 1 && (H * Wd) % 4 == 0 &&
      (((uintptr_t)dpre | (uintptr_t)dx | (uintptr_t)y_below) & 15) == 0) {
    const int64_t n4 = M / 4 * Cin;
    conv_o1_bwd_data_kernel<<<pvb::cdiv(n4, 256), 256, 0, (cudaStream_t)stream>>>(dpre, W, dx, n4, Cin, H * Wd / 4,
                                                                                  y_below, act);
    pvb::count_launch();
    return pvb::launch_status();
  }
  dim3 grid((unsigned)((M + BM - 1) / BM), (Cin + BN - 1) / BN);
  size_t smem = (size_t)Cout * kh * kw * sizeof(Tap);
  conv_pix_kernel<1><<<grid, NT, smem, (cudaStream_t)stream>>>(dpre, W, nullptr, dx, nullptr, d, 0);
  pvb::count_launch();
  if (y_below) {      // no fused form on this path: a separate (16-byte) pass
    const int64_t n = M * Cin;
    int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    act_bwd_flat_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dx, y_below, nullptr, dx, n, act);
    pvb::count_launch();
  }
  return pvb::launch_status();
}

extern "C" int pvb_conv_bwd_weight(const float* dpre, const float* x, float* dW, float* db, int B,
                                   int Cin, int Cout, int H, int Wd, int kh, int kw, void* stream) {
  ConvDims d{B, Cin, Cout, H, Wd, kh, kw};
  if (check_dims(d, "pvb_conv_bwd_weight")) return -1;
  PVB_CHECK_ARG(dpre && x && dW, "pvb_conv_bwd_weight: null pointer");
  if (B == 0) return 0;
  int64_t M = (int64_t)B * H * Wd;
  if (Cout == 1 && kh * kw == 1 && (H * Wd) % 4 == 0 && Cin <= 65535 && (((uintptr_t)x | (uintptr_t)dpre) & 15) == 0) {
    int splits = (148 * 4 + Cin - 1) / Cin;
    if (splits > B) splits = B;
    const int rows = (B + splits - 1) / splits;
    dim3 grid_o1((unsigned)Cin, (unsigned)((B + rows - 1) / rows));
    conv_o1_wgrad_kernel<<<grid_o1, 256, 0, (cudaStream_t)stream>>>(dpre, x, dW, db, B, Cin, H * Wd / 4, rows);
    pvb::count_launch();
    return pvb::launch_status();
  }
  if (Cin == 1 && M < (1ll << 31) - (1 << 20) && Wd % 4 == 0 &&
      (((uintptr_t)x | (uintptr_t)dpre) & 15) == 0) {
    const int cgroups4 = (Cout + C1_CO4 - 1) / C1_CO4;
    int64_t px_blocks = (148 * 4 + cgroups4 - 1) / cgroups4;
    if (px_blocks > (M / 4 + 255) / 256) px_blocks = (M / 4 + 255) / 256;
    dim3 grid4((unsigned)px_blocks, cgroups4);
    conv_c1_wgrad4_kernel<<<grid4, 256, 0, (cudaStream_t)stream>>>(dpre, x, dW, db, d);
    pvb::count_launch();
    return pvb::launch_status();
  }
  if (Cin == 1 && M < (1ll << 31) - (1 << 20)) {
    const int cgroups = (Cout + C1_CO - 1) / C1_CO;
    int64_t px_blocks = (148 * 4 + cgroups - 1) / cgroups;
    if (px_blocks > (M + 255) / 256) px_blocks = (M + 255) / 256;
    dim3 grid1((unsigned)px_blocks, cgroups);
    conv_c1_wgrad_kernel<<<grid1, 256, 0, (cudaStream_t)stream>>>(dpre, x, dW, db, d);
    pvb::count_launch();
    return pvb::launch_status();
  }
  int tiles = ((Cin * kh * kw + BN - 1) / BN) * ((Cout + BM - 1) / BM);
  int64_t splits = (148 * 4 + tiles - 1) / tiles;
  int64_t max_splits = (M + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int64_t per = ((M + splits - 1) / splits + BK - 1) / BK * BK;
  splits = (M + per - 1) / per;
  dim3 grid((Cin * kh * kw + BN - 1) / BN, (Cout + BM - 1) / BM, (unsigned)splits);
  conv_wgrad_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(dpre, x, dW, db, d, per);
  pvb::count_launch();
  return pvb::launch_status();
}

extern "C" int pvb_act_bwd(const float* dy, const float* y, const float* pre, float* dpre, int64_t n,
                           int act, void* stream) {
  PVB_CHECK_ARG(dy && y && dpre && n >= 0, "pvb_act_bwd: bad argument");
  PVB_CHECK_ARG(act != PVB_ACT_GELU || pre, "pvb_act_bwd: gelu needs the pre-activation");
  if (n == 0) return 0;
  int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
  act_bwd_flat_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dy, y, pre, dpre, n, act);
  pvb::count_launch();
  return pvb::launch_status();
}

extern "C" int pvb_maxpool2_fwd(const float* x, float* y, int64_t BC, int H, int Wd, int two_d,
                                void* stream) {
  PVB_CHECK_ARG(x && y && BC >= 0 && H > 0 && Wd > 1 && (!two_d || H > 1), "pvb_maxpool2_fwd: bad argument");
  int Ho = two_d ? H / 2 : H, Wo = Wd / 2;
  int64_t n = BC * Ho * Wo;
  if (n == 0) return 0;
  if (two_d && (Wd % 8) == 0 && (H % 2) == 0 && (((uintptr_t)x | (uintptr_t)y) & 15) == 0) {
    const int64_t nq = n / 4;
    maxpool2_fwd_vec_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, y, nq, Wd,
                                                                                             Wo / 4);
    pvb::count_launch();
    return pvb::launch_status();
  }
  maxpool2_fwd_kernel<<<pvb::cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, y, BC, H, Wd, Ho, Wo, two_d);
  pvb::count_launch();
  return pvb::launch_status();
}

extern "C" int pvb_maxpool2_bwd(const float* x, const float* dy, float* dx, int64_t BC, int H, int Wd,
                                int two_d, int act, void* stream) {
  PVB_CHECK_ARG(x && dy && dx && BC >= 0 && H > 0 && Wd > 1 && (!two_d || H > 1), "pvb_maxpool2_bwd: bad argument");
  PVB_CHECK_ARG(act >= 0 && act <= PVB_ACT_SIGMOID && act != PVB_ACT_GELU,
                "pvb_maxpool2_bwd: fused activation derivative must be computable from the output");
  int Ho = two_d ? H / 2 : H, Wo = Wd / 2;
  const int odd_w = Wd & 1, odd_h = two_d ? (H & 1) : 0;
  int64_t n = BC * Ho * Wo + BC * ((int64_t)odd_w * H + (int64_t)odd_h * (Wd - odd_w));
  if (n == 0) return 0;
  if (two_d && (Wd % 8) == 0 && (H % 2) == 0 &&
      (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx) & 15) == 0) {
    const int64_t nq = BC * Ho * Wo / 4;
    maxpool2_bwd_vec_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        x, dy, dx, nq, Wd, Wo / 4, act);
    pvb::count_launch();
    return pvb::launch_status();
  }
  maxpool2_bwd_kernel<<<pvb::cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, dy, dx, BC, H, Wd, Ho, Wo, two_d, act);
  pvb::count_launch();
  return pvb::launch_status();
}

extern "C" int pvb_upsample2_fwd(const float* x, float* y, int64_t BC, int H, int Wd, int two_d,
                                 int bilinear, void* stream) {
  PVB_CHECK_ARG(x && y && BC >= 0 && H > 0 && Wd > 0, "pvb_upsample2_fwd: bad argument");
  PVB_CHECK_ARG(!bilinear || two_d, "pvb_upsample2_fwd: bilinear is 2-D only (reference nets/conv.py:128-130)");
  int64_t n = BC * (two_d ? 2 * H : H) * 2 * Wd;
  if (n == 0) return 0;
  upsample2_fwd_kernel<<<pvb::cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, y, BC, H, Wd, two_d, bilinear);
  pvb::count_launch();
  return pvb::launch_status();
}

extern "C" int pvb_upsample2_bwd(const float* dy, float* dx, int64_t BC, int H, int Wd, int two_d,
                                 int bilinear, const float* y_below, int act, void* stream) {
  PVB_CHECK_ARG(dy && dx && BC >= 0 && H > 0 && Wd > 0, "pvb_upsample2_bwd: bad argument");
  PVB_CHECK_ARG(!y_below || (act >= 0 && act <= PVB_ACT_SIGMOID && act != PVB_ACT_GELU),
                "pvb_upsample2_bwd: fused activation derivative needs an activation expressed from its output");
  PVB_CHECK_ARG(!bilinear || two_d, "pvb_upsample2_bwd: bilinear is 2-D only");
  int64_t n = BC * H * Wd;
  if (n == 0) return 0;
  upsample2_bwd_kernel<<<pvb::cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(dy, dx, BC, H, Wd, two_d, bilinear,
                                                                          y_below, act);
  pvb::count_launch();
  return pvb::launch_status();
}
