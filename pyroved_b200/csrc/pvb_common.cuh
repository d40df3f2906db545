// Shared device/host helpers for libpvb (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include "../../include/pvb.h"

namespace pvb {

void set_error(const char* fmt, ...);
void count_launch();  // statistics only: kernels launched by this library

#define PVB_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      pvb::set_error(__VA_ARGS__);          \
      return -1;                            \
    }                                       \
  } while (0)

static inline int launch_status() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("CUDA launch failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- activations (torch module defaults) ---------------------------------
__device__ __forceinline__ float softplus_f(float x) {
  return x > 20.f ? x : log1pf(expf(x));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float act_fwd(float x, int act) {
  switch (act) {
    case PVB_ACT_TANH: return tanhf(x);
    case PVB_ACT_RELU: return x > 0.f ? x : 0.f;
    case PVB_ACT_LRELU: return x > 0.f ? x : 0.01f * x;
    case PVB_ACT_SOFTPLUS: return softplus_f(x);
    case PVB_ACT_GELU: return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
    case PVB_ACT_SIGMOID: return sigmoid_f(x);
    default: return x;
  }
}
// derivative expressed from the saved OUTPUT y (pre only needed for gelu)
__device__ __forceinline__ float act_grad(float y, float pre, int act) {
  switch (act) {
    case PVB_ACT_TANH: return 1.f - y * y;
    case PVB_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case PVB_ACT_LRELU: return y > 0.f ? 1.f : 0.01f;
    case PVB_ACT_SOFTPLUS: return 1.f - expf(-y);  // sigmoid(pre), y = log(1+e^pre)
    case PVB_ACT_GELU: {
      float c = 0.5f * (1.f + erff(pre * 0.70710678118654752f));
      return c + pre * 0.3989422804014327f * expf(-0.5f * pre * pre);
    }
    case PVB_ACT_SIGMOID: return y * (1.f - y);
    default: return 1.f;
  }
}

// fp32 -> (hi, lo) TF32 pair by truncation: hi = the top 11 significant bits, lo = a - hi (exact; the tensor
// core reads its top 11 bits), so a = hi + lo to 2^-22 and three MMAs (lo*hi, hi*lo, hi*hi) give an
// fp32-grade product on the warp-level tensor-core path.  (cvt.rna.tf32 is emulated on sm_100 -- seven ALU
// instructions per conversion, which made the loop issue-bound; the mask + subtract is two.)
__device__ __forceinline__ void split_tf32(float a, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(a) & 0xffffe000u;
  lo = __float_as_uint(a - __uint_as_float(hi));
}
// D(16x8) += A(16x8, row) B(8x8, col), TF32 operands, fp32 accumulators
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum for blockDim.x <= 1024 (result valid in every thread)
__device__ __forceinline__ float block_sum(float v, float* smem32) {
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem32[wid] = v;
  __syncthreads();
  int nw = (blockDim.x + 31) >> 5;
  float r = (lane < nw) ? smem32[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

// pixel p -> grid coordinates, utils/coord.py:14-18 (2-D) and :43 (1-D).
// torch.linspace(a, b, n)[i] = a + i*(b-a)/(n-1) for the first half and
// b - (n-1-i)*step for the second half (symmetric evaluation).
__device__ __forceinline__ float linspace_at(float a, float b, int n, int i) {
  if (n == 1) return a;
  float step = (b - a) / (float)(n - 1);
  return (i < n / 2) ? a + step * (float)i : b - step * (float)(n - 1 - i);
}
__device__ __forceinline__ void grid_xy(int p, int H, int W, int ndim, float& gx, float& gy) {
  if (ndim == 1) {
    gx = linspace_at(1.f, -1.f, H, p);
    gy = 0.f;
  } else {
    int i = p / W, j = p - i * W;
    gx = linspace_at(-1.f, 1.f, H, i);
    gy = linspace_at(1.f, -1.f, W, j);
  }
}

// ---- per-pixel observation terms (utils/prob.py:25-29 + torch log_prob) -------
// torch clamp_probs eps for fp32 and the matching logit bound log((1-eps)/eps)
#define PVB_PROB_EPS 1.1920928955078125e-07f

// ContinuousBernoulli(probs = p).log_prob(x) and its derivative wrt p, as torch computes them
// (torch/distributions/continuous_bernoulli.py: clamp_probs, _cont_bern_log_norm with the
// Taylor branch inside (0.499, 0.501], xlogy / xlog1py).  FAST selects the fast intrinsics.
template <bool FAST>
__device__ __forceinline__ void cont_bernoulli(float p, float x, float& ll, float& dll_dp) {
  auto lg = [](float v) { return FAST ? __logf(v) : logf(v); };
  auto l1p = [](float v) { return FAST ? __logf(1.f + v) : log1pf(v); };
  auto rcp = [](float v) { return FAST ? __fdividef(1.f, v) : 1.f / v; };
  const bool unst = (p <= 0.499f) || (p > 0.501f);
  const float cp = unst ? p : 0.499f;
  const float A = l1p(-cp) - lg(cp);
  const float ln_unst = lg(fabsf(A)) - (cp <= 0.5f ? l1p(-2.f * cp) : lg(2.f * cp - 1.f));
  const float u = p - 0.5f, x2 = u * u;
  const float taylor = 0.6931471805599453f + (4.f / 3.f + 104.f / 45.f * x2) * x2;
  const float logC = unst ? ln_unst : taylor;
  ll = (x == 0.f ? 0.f : x * lg(p)) + (x == 1.f ? 0.f : (1.f - x) * l1p(-p)) + logC;
  const float dC = unst ? (-rcp(1.f - p) - rcp(p)) * rcp(A) + 2.f * rcp(1.f - 2.f * p)
                        : 2.f * u * (4.f / 3.f + 208.f / 45.f * x2);
  dll_dp = x * rcp(p) - (1.f - x) * rcp(1.f - p) + dC;
}

__device__ __forceinline__ void obs_terms(float l, float x, int sampler, int sigmoid_d, float sig,
                                          float& ll, float& dnll_dl, float& loc) {
  if (sampler == PVB_SAMPLER_BERNOULLI) {
    if (sigmoid_d) {
      // probs = sigmoid(l) -> clamp(eps, 1-eps) -> logits; log_prob = x*lg - softplus(lg)
      float p = pvb::sigmoid_f(l);
      loc = p;
      bool in = (p >= PVB_PROB_EPS) && (p <= 1.f - PVB_PROB_EPS);
      float pc = fminf(fmaxf(p, PVB_PROB_EPS), 1.f - PVB_PROB_EPS);
      float lg = in ? l : logf(pc) - log1pf(-pc);
      float sp = lg > 0.f ? lg + log1pf(expf(-lg)) : log1pf(expf(lg));
      ll = x * lg - sp;
      dnll_dl = in ? (p - x) : 0.f;
    } else {
      float p = l;
      loc = p;
      bool in = (p >= PVB_PROB_EPS) && (p <= 1.f - PVB_PROB_EPS);
      float pc = fminf(fmaxf(p, PVB_PROB_EPS), 1.f - PVB_PROB_EPS);
      float lg = logf(pc) - log1pf(-pc);
      float sp = lg > 0.f ? lg + log1pf(expf(-lg)) : log1pf(expf(lg));
      ll = x * lg - sp;
      dnll_dl = in ? (pc - x) / (pc * (1.f - pc)) : 0.f;
    }
  } else if (sampler == PVB_SAMPLER_CONT_BERNOULLI) {
    float p0 = sigmoid_d ? pvb::sigmoid_f(l) : l;
    loc = p0;
    bool in = (p0 >= PVB_PROB_EPS) && (p0 <= 1.f - PVB_PROB_EPS);
    float p = fminf(fmaxf(p0, PVB_PROB_EPS), 1.f - PVB_PROB_EPS);
    float dll_dp;
    cont_bernoulli<false>(p, x, ll, dll_dp);
    dnll_dl = in ? -dll_dp * (sigmoid_d ? p * (1.f - p) : 1.f) : 0.f;
  } else {  // gaussian, Normal(loc, sig).log_prob(x)
    float m = sigmoid_d ? pvb::sigmoid_f(l) : l;
    loc = m;
    float d = x - m;
    float inv_var = 1.f / (sig * sig);
    ll = -0.5f * d * d * inv_var - logf(sig) - 0.91893853320467274f;
    float dm = -d * inv_var;  // d(-ll)/dm
    dnll_dl = sigmoid_d ? dm * m * (1.f - m) : dm;
  }
}

// d(-log p(x|l))/dl alone: the shortest dependent chain (two MUFU ops), for the critical path
__device__ __forceinline__ float obs_dnll_fast(float l, float x, int sampler, int sigmoid_d, float sig) {
  if (sampler == PVB_SAMPLER_BERNOULLI) {
    float p = sigmoid_d ? __fdividef(1.f, 1.f + __expf(-l)) : l;
    bool in = (p >= PVB_PROB_EPS) && (p <= 1.f - PVB_PROB_EPS);
    if (sigmoid_d) return in ? (p - x) : 0.f;
    return in ? __fdividef(p - x, p * (1.f - p)) : 0.f;
  }
  if (sampler == PVB_SAMPLER_CONT_BERNOULLI) {
    float p0 = sigmoid_d ? __fdividef(1.f, 1.f + __expf(-l)) : l;
    bool in = (p0 >= PVB_PROB_EPS) && (p0 <= 1.f - PVB_PROB_EPS);
    float p = fminf(fmaxf(p0, PVB_PROB_EPS), 1.f - PVB_PROB_EPS);
    float ll_unused, dll_dp;
    cont_bernoulli<true>(p, x, ll_unused, dll_dp);
    return in ? -dll_dp * (sigmoid_d ? p * (1.f - p) : 1.f) : 0.f;
  }
  float m = sigmoid_d ? __fdividef(1.f, 1.f + __expf(-l)) : l;
  float dm = __fdividef(m - x, sig * sig);
  return sigmoid_d ? dm * m * (1.f - m) : dm;
}
// obs_terms with fast intrinsics (__expf/__logf/__fdividef: ~1e-6 relative), same case analysis.
// Used by the fused tensor-core decoder, whose operands are fp16 anyway; its log-likelihood and
// reconstruction stay ~1e-6 relative to the exact path (tolerance of the path: 1e-3).
__device__ __forceinline__ float softplus_fast(float lg) {
  return fmaxf(lg, 0.f) + __logf(1.f + __expf(-fabsf(lg)));
}
__device__ __forceinline__ void obs_terms_fast(float l, float x, int sampler, int sigmoid_d, float sig,
                                               float& ll, float& dnll_dl, float& loc) {
  if (sampler == PVB_SAMPLER_BERNOULLI) {
    float p = sigmoid_d ? __fdividef(1.f, 1.f + __expf(-l)) : l;
    loc = p;
    bool in = (p >= PVB_PROB_EPS) && (p <= 1.f - PVB_PROB_EPS);
    float pc = fminf(fmaxf(p, PVB_PROB_EPS), 1.f - PVB_PROB_EPS);
    float lg = (sigmoid_d && in) ? l : __logf(pc) - __logf(1.f - pc);
    ll = x * lg - softplus_fast(lg);
    if (sigmoid_d) dnll_dl = in ? (p - x) : 0.f;
    else dnll_dl = in ? __fdividef(pc - x, pc * (1.f - pc)) : 0.f;
  } else if (sampler == PVB_SAMPLER_CONT_BERNOULLI) {
    float p0 = sigmoid_d ? __fdividef(1.f, 1.f + __expf(-l)) : l;
    loc = p0;
    bool in = (p0 >= PVB_PROB_EPS) && (p0 <= 1.f - PVB_PROB_EPS);
    float p = fminf(fmaxf(p0, PVB_PROB_EPS), 1.f - PVB_PROB_EPS);
    float dll_dp;
    cont_bernoulli<true>(p, x, ll, dll_dp);
    dnll_dl = in ? -dll_dp * (sigmoid_d ? p * (1.f - p) : 1.f) : 0.f;
  } else {
    float m = sigmoid_d ? __fdividef(1.f, 1.f + __expf(-l)) : l;
    loc = m;
    float d = x - m;
    float inv_var = __fdividef(1.f, sig * sig);
    ll = -0.5f * d * d * inv_var - __logf(sig) - 0.91893853320467274f;
    float dm = -d * inv_var;
    dnll_dl = sigmoid_d ? dm * m * (1.f - m) : dm;
  }
}

}  // namespace pvb
