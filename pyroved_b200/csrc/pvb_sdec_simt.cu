// Spatial decoder, generic fp32 path: fused grid + affine + first layer,
// its backward, per-pixel observation log-likelihood and the ELBO reduction.
#include "pvb_common.cuh"

namespace {

// h0[r, h] = tanh(U0[i,h] gx_p + U1[i,h] gy_p + v[i,h]),  r = i*N + p.
// One thread produces 4 consecutive hidden units of one row -> float4 store,
// a warp writes 512 contiguous bytes (HBM-write bound: R*Hd*4 bytes).
__global__ void __launch_bounds__(256)
sdec_h0_fwd_kernel(const float* __restrict__ Uv, float* __restrict__ h0, int64_t R, int N, int H,
                   int W, int ndim, int Hd) {
  const int hq = Hd >> 2;  // float4 groups per row
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = R * hq;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; t < total; t += stride) {
    int64_t r = t / hq;
    int h = (int)(t - r * hq) * 4;
    int64_t i = r / N;
    int p = (int)(r - i * N);
    float gx, gy;
    pvb::grid_xy(p, H, W, ndim, gx, gy);
    const float* u = Uv + i * 3 * Hd + h;
    float4 u0 = __ldg(reinterpret_cast<const float4*>(u));
    float4 u1 = __ldg(reinterpret_cast<const float4*>(u + Hd));
    float4 vv = __ldg(reinterpret_cast<const float4*>(u + 2 * Hd));
    float4 o;
    o.x = tanhf(fmaf(u0.x, gx, fmaf(u1.x, gy, vv.x)));
    o.y = tanhf(fmaf(u0.y, gx, fmaf(u1.y, gy, vv.y)));
    o.z = tanhf(fmaf(u0.z, gx, fmaf(u1.z, gy, vv.z)));
    o.w = tanhf(fmaf(u0.w, gx, fmaf(u1.w, gy, vv.w)));
    __stcs(reinterpret_cast<float4*>(h0 + r * Hd + h), o);
  }
}

// gUv[i, {0,1,2}, h] = sum_p dpre0[i,p,h] * {gx_p, gy_p, 1},  dpre0 = dh0 (1 - h0^2)
// one CTA per (instance, 32-wide slice of hidden units); 8 row lanes.
__global__ void __launch_bounds__(256)
sdec_h0_bwd_kernel(const float* __restrict__ dh0, const float* __restrict__ h0,
                   float* __restrict__ gUv, int N, int H, int W, int ndim, int Hd) {
  __shared__ float s0[8][33], s1[8][33], s2[8][33];
  const int64_t i = blockIdx.y;
  const int h = blockIdx.x * 32 + threadIdx.x;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  if (h < Hd) {
    for (int p = threadIdx.y; p < N; p += 8) {
      int64_t o = (i * N + p) * Hd + h;
      float hv = h0[o];
      float d = dh0[o] * (1.f - hv * hv);
      float gx, gy;
      pvb::grid_xy(p, H, W, ndim, gx, gy);
      a0 = fmaf(d, gx, a0);
      a1 = fmaf(d, gy, a1);
      a2 += d;
    }
  }
  s0[threadIdx.y][threadIdx.x] = a0;
  s1[threadIdx.y][threadIdx.x] = a1;
  s2[threadIdx.y][threadIdx.x] = a2;
  __syncthreads();
  if (threadIdx.y == 0 && h < Hd) {
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { t0 += s0[k][threadIdx.x]; t1 += s1[k][threadIdx.x]; t2 += s2[k][threadIdx.x]; }
    float* g = gUv + i * 3 * Hd;
    g[h] = t0;
    g[Hd + h] = t1;
    g[2 * Hd + h] = t2;
  }
}

// ---- observation log-likelihood (obs_terms lives in pvb_common.cuh) ----------
using pvb::obs_terms;

__global__ void __launch_bounds__(256)
obs_loglik_kernel(const float* __restrict__ logit, const float* __restrict__ x,
                  const float* __restrict__ w, float* __restrict__ rowll,
                  float* __restrict__ dlogit, float* __restrict__ loc, int64_t R, int64_t B, int N,
                  int sampler, int sigmoid_d, float sig) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; r < R; r += stride) {
    int64_t i = r / N;
    int p = (int)(r - i * N);
    float xv = x[(i % B) * N + p];
    float ll, dn, lc;
    obs_terms(logit[r], xv, sampler, sigmoid_d, sig, ll, dn, lc);
    rowll[r] = ll;
    if (dlogit) dlogit[r] = (w ? w[i] : 1.f) * dn;
    if (loc) loc[r] = lc;
  }
}

// ll[i] = sum_p rowll[i,p]: one warp per instance, lane-strided partial sums
// then a shuffle tree (fixed order -> bitwise reproducible).
__global__ void __launch_bounds__(256)
rowll_reduce_kernel(const float* __restrict__ rowll, float* __restrict__ ll, int64_t I, int N) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (i >= I) return;
  const float* row = rowll + i * N;
  float s = 0.f;
  for (int p = lane; p < N; p += 32) s += row[p];
  s = pvb::warp_sum(s);
  if (lane == 0) ll[i] = s;
}

// loss_out[0] (+)= -sum_i w_i (ll_i + beta kl_i): single block, fixed order
__global__ void __launch_bounds__(1024)
elbo_total_kernel(const float* __restrict__ ll, const float* __restrict__ kl,
                  const float* __restrict__ w, float beta, float* __restrict__ loss_out,
                  int accumulate, int64_t I) {
  __shared__ float sm[32];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < I; i += blockDim.x) {
    float t = ll[i] + (kl ? beta * kl[i] : 0.f);
    s += (w ? w[i] : 1.f) * t;
  }
  s = pvb::block_sum(s, sm);
  if (threadIdx.x == 0) loss_out[0] = accumulate ? loss_out[0] - s : -s;
}

}  // namespace

extern "C" int pvb_sdec_h0_fwd(const float* Uv, float* h0, int64_t I, int H, int W, int ndim,
                               int Hd, void* stream) {
  PVB_CHECK_ARG(Uv && h0 && I >= 0 && H > 0 && W > 0, "pvb_sdec_h0_fwd: bad argument");
  PVB_CHECK_ARG(ndim == 1 || ndim == 2, "pvb_sdec_h0_fwd: ndim must be 1 or 2");
  PVB_CHECK_ARG(Hd > 0 && Hd % 4 == 0, "pvb_sdec_h0_fwd: hidden size must be a multiple of 4");
  int N = (ndim == 1) ? H : H * W;
  int64_t R = I * N;
  if (R == 0) return 0;
  int64_t total = R * (Hd / 4);
  int64_t blocks = (total + 255) / 256;
  int64_t cap = 148LL * 8 * 4;  // multiple of the SM count, grid-stride beyond
  if (blocks > cap) blocks = cap;
  sdec_h0_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(Uv, h0, R, N, H, W, ndim, Hd); pvb::count_launch();
  return pvb::launch_status();
}

extern "C" int pvb_sdec_h0_bwd(const float* dh0, const float* h0, float* gUv, int64_t I, int H,
                               int W, int ndim, int Hd, void* stream) {
  PVB_CHECK_ARG(dh0 && h0 && gUv && I >= 0 && H > 0 && W > 0 && Hd > 0, "pvb_sdec_h0_bwd: bad argument");
  PVB_CHECK_ARG(ndim == 1 || ndim == 2, "pvb_sdec_h0_bwd: ndim must be 1 or 2");
  PVB_CHECK_ARG(I <= 65535 * 1024LL, "pvb_sdec_h0_bwd: too many instances");
  if (I == 0) return 0;
  int N = (ndim == 1) ? H : H * W;
  dim3 blk(32, 8);
  for (int64_t i0 = 0; i0 < I; i0 += 65535) {
    int64_t ni = I - i0 < 65535 ? I - i0 : 65535;
    dim3 grid((Hd + 31) / 32, (unsigned)ni);
    sdec_h0_bwd_kernel<<<grid, blk, 0, (cudaStream_t)stream>>>(
        dh0 + i0 * N * Hd, h0 + i0 * N * Hd, gUv + i0 * 3 * Hd, N, H, W, ndim, Hd); pvb::count_launch();
  }
  return pvb::launch_status();
}

extern "C" int pvb_obs_loglik(const float* logit, const float* x, const float* w, float* rowll,
                              float* dlogit, float* loc, int64_t I, int64_t B, int N, int sampler,
                              int sigmoid_d, float decoder_sig, void* stream) {
  PVB_CHECK_ARG(logit && x && rowll && I >= 0 && B > 0 && N > 0, "pvb_obs_loglik: bad argument");
  PVB_CHECK_ARG(sampler >= PVB_SAMPLER_BERNOULLI && sampler <= PVB_SAMPLER_CONT_BERNOULLI,
                "pvb_obs_loglik: unknown sampler %d", sampler);
  PVB_CHECK_ARG(sampler != PVB_SAMPLER_GAUSSIAN || decoder_sig > 0.f, "pvb_obs_loglik: decoder_sig must be > 0");
  int64_t R = I * N;
  if (R == 0) return 0;
  int64_t blocks = (R + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  obs_loglik_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      logit, x, w, rowll, dlogit, loc, R, B, N, sampler, sigmoid_d, decoder_sig); pvb::count_launch();
  return pvb::launch_status();
}

extern "C" int pvb_elbo_reduce(const float* rowll, const float* kl, const float* w, float beta,
                               float* ll, float* loss_out, int accumulate, int64_t I, int N,
                               void* stream) {
  PVB_CHECK_ARG(rowll && ll && I >= 0 && N > 0, "pvb_elbo_reduce: bad argument");
  if (I == 0) return 0;
  rowll_reduce_kernel<<<pvb::cdiv(I * 32, 256), 256, 0, (cudaStream_t)stream>>>(rowll, ll, I, N); pvb::count_launch();
  if (loss_out) {
    elbo_total_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(ll, kl, w, beta, loss_out, accumulate, I);
    pvb::count_launch();
  }
  return pvb::launch_status();
}
