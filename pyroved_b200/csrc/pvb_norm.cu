// nn.BatchNorm{1,2}d of the VED convolutional nets (reference nets/conv.py:187,240 insert
// get_bnorm(ndim)(ch) after every conv + activation when batchnorm=True; utils/nn.py:103-105).
// x [B, C, HW] fp32 (NCHW).  HBM-bound: the training forward reads x twice (statistics, apply)
// and writes y once; the backward reads dy and x twice and writes dx once.
//
//   forward (training)   mean_c, var_c (biased) over the B*HW elements of channel c;
//                        y = (x - mean_c) / sqrt(var_c + eps) * gamma_c + beta_c;
//                        running_mean = (1-m) running_mean + m mean;
//                        running_var  = (1-m) running_var  + m var n/(n-1);  num_batches_tracked += 1
//   forward (eval)       y = (x - running_mean_c) / sqrt(running_var_c + eps) * gamma_c + beta_c
//   backward (training)  dbeta_c += sum dy;  dgamma_c += sum dy xhat;
//                        dx = gamma_c invstd_c (dy - sum dy / n - xhat sum(dy xhat) / n)
//
// Sums: NS fixed slices per channel reduced in a fixed order (deterministic), slice sums in fp32
// about a per-channel shift (the channel's first element) so var = E[d^2] - E[d]^2 does not
// cancel, combined in fp64.
#include "pvb_common.cuh"

namespace {

constexpr int NS = 64;        // slices per channel
constexpr int NT = 256;

__device__ __forceinline__ void block_sum2(float& a, float& b, float* sh) {
  a = pvb::warp_sum(a);
  b = pvb::warp_sum(b);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[warp] = a; sh[8 + warp] = b; }
  __syncthreads();
  if (warp == 0) {
    a = lane < NT / 32 ? sh[lane] : 0.f;
    b = lane < NT / 32 ? sh[8 + lane] : 0.f;
    a = pvb::warp_sum(a);
    b = pvb::warp_sum(b);
  }
}

// MODE 0: (sum (x - k), sum (x - k)^2), k = x[0, c, 0]
// MODE 1: (sum dy, sum dy (x - mean_c) invstd_c)
template <int MODE, int VEC>
__global__ void __launch_bounds__(NT)
bn_partial_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                  const float* __restrict__ mean, const float* __restrict__ invstd,
                  float* __restrict__ part, int C, int HW, int n, int chunk) {
  __shared__ float sh[16];
  const int c = blockIdx.y, s = blockIdx.x;
  const float k = MODE == 0 ? __ldg(x + (int64_t)c * HW) : __ldg(mean + c);
  const float is = MODE == 0 ? 1.f : __ldg(invstd + c);
  const int e_end = min(n, (s + 1) * chunk);
  float a = 0.f, b = 0.f;
  for (int e = s * chunk + threadIdx.x * VEC; e < e_end; e += NT * VEC) {
    const int bi = e / HW, r = e - bi * HW;
    const int64_t off = ((int64_t)bi * C + c) * HW + r;
    if (VEC == 4) {
      const float4 xv = *reinterpret_cast<const float4*>(x + off);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
      float gs[4] = {0.f, 0.f, 0.f, 0.f};
      if (MODE == 1) {
        const float4 gv = *reinterpret_cast<const float4*>(dy + off);
        gs[0] = gv.x; gs[1] = gv.y; gs[2] = gv.z; gs[3] = gv.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float d = xs[j] - k;
        if (MODE == 0) { a += d; b = fmaf(d, d, b); }
        else { a += gs[j]; b = fmaf(gs[j], d * is, b); }
      }
    } else {
      const float d = x[off] - k;
      if (MODE == 0) { a += d; b = fmaf(d, d, b); }
      else { const float g = dy[off]; a += g; b = fmaf(g, d * is, b); }
    }
  }
  block_sum2(a, b, sh);
  if (threadIdx.x == 0) {
    part[((int64_t)c * NS + s) * 2] = a;
    part[((int64_t)c * NS + s) * 2 + 1] = b;
  }
}

__global__ void bn_stats_finalize_kernel(const float* __restrict__ x, const float* __restrict__ part,
                                         float* __restrict__ running_mean, float* __restrict__ running_var,
                                         long long* __restrict__ nbt, float* __restrict__ save_mean,
                                         float* __restrict__ save_invstd, int C, int HW, int n, float eps,
                                         float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && nbt) *nbt += 1;
  if (c >= C) return;
  double s1 = 0.0, s2 = 0.0;
  for (int s = 0; s < NS; ++s) {
    s1 += (double)part[((int64_t)c * NS + s) * 2];
    s2 += (double)part[((int64_t)c * NS + s) * 2 + 1];
  }
  const double m1 = s1 / n;
  double var = s2 / n - m1 * m1;
  if (var < 0.0) var = 0.0;
  const float mean = (float)((double)x[(int64_t)c * HW] + m1);
  save_mean[c] = mean;
  save_invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
  if (running_var) {
    const double unbiased = n > 1 ? var * (double)n / (double)(n - 1) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// eval mode: (mean, invstd) from the running statistics
__global__ void bn_eval_stats_kernel(const float* __restrict__ running_mean,
                                     const float* __restrict__ running_var, float* __restrict__ save_mean,
                                     float* __restrict__ save_invstd, int C, float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  save_mean[c] = running_mean[c];
  save_invstd[c] = rsqrtf(running_var[c] + eps);
}

// dgamma / dbeta accumulate; (sum dy, sum dy xhat) / n kept for the element-wise pass
__global__ void bn_bwd_finalize_kernel(float* __restrict__ part, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, int C, int n, int training) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s1 = 0.0, s2 = 0.0;
  for (int s = 0; s < NS; ++s) {
    s1 += (double)part[((int64_t)c * NS + s) * 2];
    s2 += (double)part[((int64_t)c * NS + s) * 2 + 1];
  }
  if (dbeta) dbeta[c] += (float)s1;
  if (dgamma) dgamma[c] += (float)s2;
  // eval mode: the statistics are constants, dx = gamma invstd dy (no mean terms)
  part[(int64_t)c * NS * 2] = training ? (float)(s1 / n) : 0.f;
  part[(int64_t)c * NS * 2 + 1] = training ? (float)(s2 / n) : 0.f;
}

// MODE 0: y = (x - mean) invstd gamma + beta
// MODE 1: dx = gamma invstd (dy - m_dy - xhat m_dyx), (m_dy, m_dyx) = part[c][0]
template <int MODE, int VEC>
__global__ void __launch_bounds__(NT)
bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma,
                const float* __restrict__ beta, const float* __restrict__ mean,
                const float* __restrict__ invstd, const float* __restrict__ part, float* __restrict__ out,
                int C, int HW, int64_t total) {
  int64_t i = ((int64_t)blockIdx.x * NT + threadIdx.x) * VEC;
  const int64_t stride = (int64_t)gridDim.x * NT * VEC;
  for (; i < total; i += stride) {
    const int c = (int)((i / HW) % C);
    const float mu = __ldg(mean + c), is = __ldg(invstd + c);
    const float g = gamma ? __ldg(gamma + c) : 1.f;
    float m_dy = 0.f, m_dyx = 0.f, sh = 0.f;
    if (MODE == 0) sh = beta ? __ldg(beta + c) : 0.f;
    else { m_dy = __ldg(part + (int64_t)c * NS * 2); m_dyx = __ldg(part + (int64_t)c * NS * 2 + 1); }
    if (VEC == 4) {
      const float4 xv = *reinterpret_cast<const float4*>(x + i);
      float v[4] = {xv.x, xv.y, xv.z, xv.w};
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = fmaf((v[j] - mu) * is, g, sh);
      } else {
        const float4 gv = *reinterpret_cast<const float4*>(dy + i);
        const float gs[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = g * is * (gs[j] - m_dy - (v[j] - mu) * is * m_dyx);
      }
      *reinterpret_cast<float4*>(out + i) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      const float xh = (x[i] - mu) * is;
      out[i] = MODE == 0 ? fmaf(xh, g, sh) : g * is * (dy[i] - m_dy - xh * m_dyx);
    }
  }
}

inline bool vec_ok(const void* a, const void* b, const void* c, int64_t HW) {
  return HW % 4 == 0 && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) == 0;
}

inline int apply_blocks(int64_t total, int vec) {
  int64_t b = (total / vec + NT - 1) / NT;
  return (int)(b < 148 * 8 ? (b > 0 ? b : 1) : 148 * 8);
}

}  // namespace

extern "C" int64_t pvb_bn_workspace_bytes(int C) { return (int64_t)C * NS * 2 * sizeof(float); }

extern "C" int pvb_bn_fwd(const float* x, const float* gamma, const float* beta, float* running_mean,
                          float* running_var, int64_t* num_batches_tracked, float* y, float* save_mean,
                          float* save_invstd, void* workspace, int B, int C, int64_t HW, float eps,
                          float momentum, int training, void* stream) {
  PVB_CHECK_ARG(x && y && save_mean && save_invstd && workspace, "pvb_bn_fwd: null pointer");
  PVB_CHECK_ARG(B >= 0 && C > 0 && HW > 0 && (int64_t)B * HW < (1ll << 31), "pvb_bn_fwd: bad shape");
  PVB_CHECK_ARG(training || (running_mean && running_var), "pvb_bn_fwd: eval mode needs running statistics");
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int n = (int)(B * HW);
  const bool vec = vec_ok(x, y, nullptr, HW);
  float* part = reinterpret_cast<float*>(workspace);
  if (training) {
    PVB_CHECK_ARG(n > 1, "pvb_bn_fwd: training needs more than one value per channel");
    int chunk = (n + NS - 1) / NS;
    chunk = (chunk + 3) / 4 * 4;
    dim3 grid(NS, C);
    if (vec) bn_partial_kernel<0, 4><<<grid, NT, 0, st>>>(x, nullptr, nullptr, nullptr, part, C, (int)HW, n, chunk);
    else bn_partial_kernel<0, 1><<<grid, NT, 0, st>>>(x, nullptr, nullptr, nullptr, part, C, (int)HW, n, chunk);
    pvb::count_launch();
    bn_stats_finalize_kernel<<<pvb::cdiv(C, 128), 128, 0, st>>>(
        x, part, running_mean, running_var, reinterpret_cast<long long*>(num_batches_tracked), save_mean,
        save_invstd, C, (int)HW, n, eps, momentum);
    pvb::count_launch();
  } else {
    bn_eval_stats_kernel<<<pvb::cdiv(C, 128), 128, 0, st>>>(running_mean, running_var, save_mean,
                                                            save_invstd, C, eps);
    pvb::count_launch();
  }
  const int64_t total = (int64_t)B * C * HW;
  if (vec)
    bn_apply_kernel<0, 4><<<apply_blocks(total, 4), NT, 0, st>>>(x, nullptr, gamma, beta, save_mean,
                                                                save_invstd, nullptr, y, C, (int)HW, total);
  else
    bn_apply_kernel<0, 1><<<apply_blocks(total, 1), NT, 0, st>>>(x, nullptr, gamma, beta, save_mean,
                                                                save_invstd, nullptr, y, C, (int)HW, total);
  pvb::count_launch();
  return pvb::launch_status();
}

extern "C" int pvb_bn_bwd(const float* dy, const float* x, const float* gamma, const float* save_mean,
                          const float* save_invstd, float* dx, float* dgamma, float* dbeta, void* workspace,
                          int B, int C, int64_t HW, int training, void* stream) {
  PVB_CHECK_ARG(dy && x && save_mean && save_invstd && dx && workspace, "pvb_bn_bwd: null pointer");
  PVB_CHECK_ARG(B >= 0 && C > 0 && HW > 0 && (int64_t)B * HW < (1ll << 31), "pvb_bn_bwd: bad shape");
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int n = (int)(B * HW);
  const bool vec = vec_ok(x, dy, dx, HW);
  float* part = reinterpret_cast<float*>(workspace);
  int chunk = (n + NS - 1) / NS;
  chunk = (chunk + 3) / 4 * 4;
  dim3 grid(NS, C);
  if (vec) bn_partial_kernel<1, 4><<<grid, NT, 0, st>>>(x, dy, save_mean, save_invstd, part, C, (int)HW, n, chunk);
  else bn_partial_kernel<1, 1><<<grid, NT, 0, st>>>(x, dy, save_mean, save_invstd, part, C, (int)HW, n, chunk);
  pvb::count_launch();
  bn_bwd_finalize_kernel<<<pvb::cdiv(C, 128), 128, 0, st>>>(part, dgamma, dbeta, C, n, training);
  pvb::count_launch();
  const int64_t total = (int64_t)B * C * HW;
  if (vec)
    bn_apply_kernel<1, 4><<<apply_blocks(total, 4), NT, 0, st>>>(x, dy, gamma, nullptr, save_mean,
                                                                save_invstd, part, dx, C, (int)HW, total);
  else
    bn_apply_kernel<1, 1><<<apply_blocks(total, 1), NT, 0, st>>>(x, dy, gamma, nullptr, save_mean,
                                                                save_invstd, part, dx, C, (int)HW, total);
  pvb::count_launch();
  return pvb::launch_status();
}
