// Fused small-batch MLP kernels for the encoder / classifier side of the SVI step.
//
// At SVI batch sizes the guide's encoder (nets/fc.py:51-61, 97-108, 264-271) is
// launch-latency bound: [M <= a few thousand] x [<= 256] activations, a dozen tiny
// GEMM / bias / activation / column-sum launches per direction.  Here:
//   pvb_mlp_tail_fwd    hidden layers >= 1, the linear heads, the reparameterised sample
//                       (+ Philox noise), the sampled KL and the coordinate-transform fold
//                       -- one launch, RB batch rows per CTA, activations stay in smem
//   pvb_mlp_chain_bwd   heads -> hidden stack backward (gradients wrt pre-activations)
//   pvb_mlp_wgrad       every dW / db of the stack as one grouped launch
//   pvb_latent_side_bwd dUv gather + fold backward + latent backward per instance
// The first (wide) layer stays a plain GEMM (pvb_linear_fwd with fused bias + activation).
#include "pvb_common.cuh"
#include "pvb_fold.cuh"

namespace {

constexpr int RB = 4;        // batch rows per CTA (2 and 8 measured at batch 512: tail 18.5 / 16.5 us against 12.6)
constexpr int MT = 256;      // threads per CTA (thread n owns output column n of all RB rows)
constexpr int MAXW = PVB_MLP_MAX_WIDTH;
constexpr int MAXHD = PVB_MLP_MAX_HEAD_DIM;
constexpr int KS = 128;      // K extent of the staged weight block
constexpr int WLD = KS + 4;  // row stride of the staged block (16-byte rows, conflict-free float4 reads)
constexpr int WS_BYTES = MAXW * WLD * 4;

__device__ __forceinline__ void cpa16(float* smem_dst, const float* gsrc) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cpa4(float* smem_dst, const float* gsrc) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cpa_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// Ws[n][0..kc) = W[n][k0..k0+kc) for n < N (W [N][K] row-major): every copy of the block is in
// flight at once, so the block costs one memory latency
__device__ __forceinline__ void stage_rows(const float* __restrict__ W, int N, int K, int k0, int kc,
                                           float* Ws) {
  const int tid = threadIdx.x;
  if ((K & 3) == 0 && (kc & 3) == 0) {
    const int q4 = kc >> 2;
    for (int idx = tid; idx < N * q4; idx += MT) {
      int n = idx / q4, q = idx - n * q4;
      cpa16(Ws + n * WLD + 4 * q, W + (size_t)n * K + k0 + 4 * q);
    }
  } else {
    for (int idx = tid; idx < N * kc; idx += MT) {
      int n = idx / kc, k = idx - n * kc;
      cpa4(Ws + n * WLD + k, W + (size_t)n * K + k0 + k);
    }
  }
  cpa_wait_all();
}

// out_s[r][n] = act(b[n] + sum_k in_s[r][k] W[n][k]),  r < RB, n < N <= 256, K <= 256
__device__ __forceinline__ void dense_rows(const float* __restrict__ W, const float* __restrict__ b,
                                           int N, int K, const float* in_s, float* out_s, int act,
                                           float* __restrict__ h_g, float* __restrict__ pre_g,
                                           int64_t row0, int64_t M, float* Ws) {
  const int n = threadIdx.x;
  float acc[RB];
#pragma unroll
  for (int r = 0; r < RB; ++r) acc[r] = 0.f;
  for (int k0 = 0; k0 < K; k0 += KS) {
    const int kc = (K - k0 < KS) ? K - k0 : KS;
    __syncthreads();   // previous block consumed; in_s complete
    stage_rows(W, N, K, k0, kc, Ws);
    __syncthreads();
    if (n < N) {
      const float* w = Ws + n * WLD;
      int k = 0;
      for (; k + 4 <= kc; k += 4) {
        float4 w4 = *reinterpret_cast<const float4*>(w + k);
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          float4 x4 = *reinterpret_cast<const float4*>(in_s + r * MAXW + k0 + k);
          acc[r] = fmaf(x4.x, w4.x, fmaf(x4.y, w4.y, fmaf(x4.z, w4.z, fmaf(x4.w, w4.w, acc[r]))));
        }
      }
      for (; k < kc; ++k)
#pragma unroll
        for (int r = 0; r < RB; ++r) acc[r] = fmaf(in_s[r * MAXW + k0 + k], w[k], acc[r]);
    }
  }
  if (n < N) {
    float bn = b ? __ldg(b + n) : 0.f;
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      float v = acc[r] + bn;
      float a = pvb::act_fwd(v, act);
      out_s[r * MAXW + n] = a;
      if (row0 + r < M) {
        if (h_g) h_g[(row0 + r) * N + n] = a;
        if (pre_g) pre_g[(row0 + r) * N + n] = v;
      }
    }
  }
}

// small linear head: out_s[r][n] = b[n] + in_s[r][:] . W[n][:]  -- one warp per (row, output)
__device__ __forceinline__ void head_rows(const float* __restrict__ W, const float* __restrict__ b,
                                          int N, int K, const float* in_s, float* out_s,
                                          float* __restrict__ out_g, int64_t row0, int64_t M) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int idx = warp; idx < RB * N; idx += MT / 32) {
    int r = idx / N, n = idx - r * N;
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s = fmaf(in_s[r * MAXW + k], __ldg(W + (size_t)n * K + k), s);
    s = pvb::warp_sum(s);
    if (lane == 0) {
      float v = s + (b ? __ldg(b + n) : 0.f);
      out_s[r * MAXHD + n] = v;
      if (row0 + r < M && out_g) out_g[(row0 + r) * N + n] = v;
    }
  }
}

__global__ void __launch_bounds__(MT) mlp_tail_fwd_kernel(pvb_mlp_tail_args a) {
  extern __shared__ __align__(16) float Ws[];   // [MAXW][WLD]
  __shared__ __align__(16) float bufA[RB * MAXW];
  __shared__ __align__(16) float bufB[RB * MAXW];
  __shared__ float heads[PVB_MLP_MAX_HEADS][RB * MAXHD];
  __shared__ float zs[RB * MAXHD];
  __shared__ pvb::Xform xf[RB];
  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * RB;
  // stage the input rows (zero beyond M)
  for (int idx = tid; idx < RB * a.w_in; idx += MT) {
    int r = idx / a.w_in, k = idx % a.w_in;
    bufA[r * MAXW + k] = (row0 + r < a.M) ? a.h_in[(row0 + r) * a.w_in + k] : 0.f;
  }
  float* cur = bufA;
  float* nxt = bufB;
  int K = a.w_in;
  for (int l = 0; l < a.n_layers; ++l) {
    dense_rows(a.W[l], a.b[l], a.width[l], K, cur, nxt, a.act, a.h[l], a.pre[l], row0, a.M, Ws);
    float* t = cur; cur = nxt; nxt = t;
    K = a.width[l];
  }
  __syncthreads();
  for (int k = 0; k < a.n_heads; ++k)
    head_rows(a.hW[k], a.hb[k], a.hdim[k], K, cur, heads[k], a.hout[k], row0, a.M);
  if (!a.gauss) return;
  __syncthreads();
  // reparameterised sample + sampled KL (pvb_randn + pvb_latent_fwd)
  const int Z = a.hdim[0];
  if (tid < RB && row0 + tid < a.M) {
    const int r = tid;
    const int64_t i = row0 + r;
    const uint32_t step = (a.gen_eps && a.step_counter) ? (uint32_t)(*a.step_counter) : 0u;
    float acc = 0.f;
    for (int d = 0; d < Z; ++d) {
      const int64_t o = i * Z + d;
      float e;
      if (a.gen_eps) {
        e = pvb::philox_randn((uint64_t)(a.first_index + o), step, a.seed);
        a.eps[o] = e;
      } else {
        e = a.eps[o];
      }
      float sg = pvb::softplus_f(heads[1][r * MAXHD + d]);
      float zz = fmaf(sg, e, heads[0][r * MAXHD + d]);
      a.sigma[o] = sg;
      a.z[o] = zz;
      zs[r * MAXHD + d] = zz;
      acc += -0.5f * zz * zz + 0.5f * e * e + logf(sg);
    }
    a.kl[i] = acc;
    if (a.fold) xf[r] = pvb::xform_of(a.cfg, pvb::split_of(a.cfg), zs + r * MAXHD);
  }
  if (!a.fold) return;
  __syncthreads();
  const pvb::Split sp = pvb::split_of(a.cfg);
  const int Hd = a.cfg.hidden;
  for (int idx = tid; idx < RB * Hd; idx += MT) {
    int r = idx / Hd, h = idx % Hd;
    int64_t i = row0 + r;
    if (i >= a.M) continue;
    const float* cond_i = a.cond ? a.cond + i * a.cfg.cond_dim : nullptr;
    pvb::fold_unit(a.cfg, sp, xf[r], zs + r * MAXHD, cond_i, a.Wc, a.bc, a.Wz, h,
                   a.Uv + i * 3 * Hd);
  }
}

// ---- heads + hidden stack backward ---------------------------------------------------------
__global__ void __launch_bounds__(MT) mlp_chain_bwd_kernel(pvb_mlp_chain_args a) {
  extern __shared__ __align__(16) float Ws[];   // [MAXW][WLD]: rows n of W[l], K block
  __shared__ float gs[PVB_MLP_MAX_HEADS][RB * MAXHD];
  __shared__ __align__(16) float dp[RB * MAXW];
  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * RB;
  for (int k = 0; k < a.n_heads; ++k)
    for (int idx = tid; idx < RB * a.hdim[k]; idx += MT) {
      int r = idx / a.hdim[k], n = idx % a.hdim[k];
      gs[k][r * MAXHD + n] = (row0 + r < a.M) ? a.g[k][(row0 + r) * a.hdim[k] + n] : 0.f;
    }
  __syncthreads();
  float acc[RB];
#pragma unroll
  for (int r = 0; r < RB; ++r) acc[r] = 0.f;
  int l = a.n_layers - 1;
  // through the heads: dh[r][k] = sum_heads sum_n g[r][n] hW[n][k]   (coalesced over k)
  if (tid < a.width[l]) {
    for (int k = 0; k < a.n_heads; ++k) {
      const float* W = a.hW[k];
      const int Kw = a.width[l];
#pragma unroll 8
      for (int n = 0; n < a.hdim[k]; ++n) {
        float w = __ldg(W + (size_t)n * Kw + tid);
#pragma unroll
        for (int r = 0; r < RB; ++r) acc[r] = fmaf(gs[k][r * MAXHD + n], w, acc[r]);
      }
    }
  }
  for (; l >= 0; --l) {
    const int Wd = a.width[l];
    // dpre = dh * act'(h)
    if (tid < Wd) {
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        float d = 0.f;
        if (row0 + r < a.M) {
          int64_t o = (row0 + r) * Wd + tid;
          float p = a.pre[l] ? a.pre[l][o] : 0.f;
          d = acc[r] * pvb::act_grad(a.h[l][o], p, a.act);
          a.dpre[l][o] = d;
        }
        dp[r * MAXW + tid] = d;
      }
    }
    if (l == 0) break;
    // dh_prev[r][k] = sum_n dpre[r][n] W[l][n][k]: W rows staged once per K block
    const int Kp = a.width[l - 1];
#pragma unroll
    for (int r = 0; r < RB; ++r) acc[r] = 0.f;
    for (int k0 = 0; k0 < Kp; k0 += KS) {
      const int kc = (Kp - k0 < KS) ? Kp - k0 : KS;
      __syncthreads();   // dp complete; previous block consumed
      stage_rows(a.W[l], Wd, Kp, k0, kc, Ws);
      __syncthreads();
      const int k = tid - k0;
      if (k >= 0 && k < kc) {
        int n = 0;
        for (; n + 4 <= Wd; n += 4) {
          float w0 = Ws[n * WLD + k], w1 = Ws[(n + 1) * WLD + k];
          float w2 = Ws[(n + 2) * WLD + k], w3 = Ws[(n + 3) * WLD + k];
#pragma unroll
          for (int r = 0; r < RB; ++r) {
            float4 d4 = *reinterpret_cast<const float4*>(dp + r * MAXW + n);
            acc[r] = fmaf(d4.x, w0, fmaf(d4.y, w1, fmaf(d4.z, w2, fmaf(d4.w, w3, acc[r]))));
          }
        }
        for (; n < Wd; ++n)
#pragma unroll
          for (int r = 0; r < RB; ++r) acc[r] = fmaf(dp[r * MAXW + n], Ws[n * WLD + k], acc[r]);
      }
    }
    __syncthreads();   // dp consumed before the next layer overwrites it
  }
}

// ---- grouped weight gradients ------------------------------------------------------------------
// One 32 x 32 tile of one dW per CTA; the reduction over the M batch rows is split over WG_G = 4
// thread groups of the CTA (own cp.async ring + named barrier each), partial tiles summed in a
// fixed order.
// (two cp.async stages: with the 40-float rows three would take 123 KB and leave one CTA per SM; the wide
// first layer of the 64 x 64 models is 532 tiles)
constexpr int WG_T = 32, WG_NT = 128, WG_MAXP = 8, WG_ST = 2, WG_G = 4;
constexpr int WG_THREADS = WG_G * WG_NT;
constexpr int WG_LD = WG_T + 8;     // row stride = 8 mod 32 banks: the MMA fragment loads (4 rows x 8 columns) are conflict-free
constexpr int WG_SMEM = WG_G * WG_ST * 2 * WG_T * WG_LD * 4;
struct WgradArgs {
  int64_t M;
  int n_prob;
  pvb_wgrad_problem p[WG_MAXP];
  int tile0[WG_MAXP + 1];
};

__device__ __forceinline__ void cpa4z(float* smem_dst, const float* gsrc, bool pred) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  int sz = pred ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}

__global__ void __launch_bounds__(WG_THREADS) mlp_wgrad_kernel(WgradArgs a) {
  extern __shared__ __align__(16) float wg_smem[];
  const int g = threadIdx.x >> 7, tid = threadIdx.x & 127;
  float (*As)[WG_T][WG_LD] = reinterpret_cast<float (*)[WG_T][WG_LD]>(wg_smem + g * WG_ST * 2 * WG_T * WG_LD);
  float (*Bs)[WG_T][WG_LD] = reinterpret_cast<float (*)[WG_T][WG_LD]>(wg_smem + g * WG_ST * 2 * WG_T * WG_LD +
                                                                     WG_ST * WG_T * WG_LD);
  int pi = 0;
  while (pi + 1 < a.n_prob && (int)blockIdx.x >= a.tile0[pi + 1]) ++pi;
  const pvb_wgrad_problem pr = a.p[pi];
  const int t = blockIdx.x - a.tile0[pi];
  const int tk_n = (pr.K + WG_T - 1) / WG_T;
  const int n0 = (t / tk_n) * WG_T, kk0 = (t % tk_n) * WG_T;
  // warp w of a group owns the 16 x 16 sub-tile (n rows 16 (w & 1), k columns 16 (w >> 1)) of the group's partial
  // tile as two m16n8k8 MMAs per 8 batch rows, 3 x TF32 (fp32-grade, see split_tf32); as FFMA on a 2 x 4 register
  // tile the loop issued 3 shared-memory loads per 8 FMA and the kernel was bound by its instruction stream
  const int wq = tid >> 5, lane = tid & 31, fg = lane >> 2, ft = lane & 3;
  const int wr = (wq & 1) * 16, wc = (wq >> 1) * 16;
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  float bsum = 0.f;   // threads 0..31 of each group in the k-tile 0 CTA: column sum of d
  const bool do_bias = pr.db && kk0 == 0;
  const int n_chunks = (int)((a.M + WG_T - 1) / WG_T);
  const int my_n = (n_chunks - g + WG_G - 1) / WG_G;
  const bool vec = ((pr.N & 3) == 0) && ((pr.K & 3) == 0) && (((uintptr_t)pr.d & 15) == 0) &&
                   (((uintptr_t)pr.x & 15) == 0);
  auto stage = [&](int ci) {
    const int buf = ci % WG_ST;
    const int64_t m0 = (int64_t)(g + ci * WG_G) * WG_T;
    if (vec) {
      // 256 16-byte pieces per operand tile, 2 per thread (a piece is all-in or all-out)
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        int idx = tid + it * WG_NT;
        int m = idx >> 3, q = idx & 7;
        int64_t gm = m0 + m;
        bool oka = gm < a.M && n0 + 4 * q < pr.N, okb = gm < a.M && kk0 + 4 * q < pr.K;
        unsigned da = (unsigned)__cvta_generic_to_shared(&As[buf][m][4 * q]);
        unsigned db = (unsigned)__cvta_generic_to_shared(&Bs[buf][m][4 * q]);
        const float* sa = oka ? pr.d + gm * pr.N + n0 + 4 * q : pr.d;
        const float* sb = okb ? pr.x + gm * pr.K + kk0 + 4 * q : pr.x;
        int za = oka ? 16 : 0, zb = okb ? 16 : 0;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(da), "l"(sa), "r"(za) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(db), "l"(sb), "r"(zb) : "memory");
      }
      return;
    }
#pragma unroll
    for (int it = 0; it < (WG_T * WG_T) / WG_NT; ++it) {
      int idx = tid + it * WG_NT;
      int c = idx % WG_T, m = idx / WG_T;
      int64_t gm = m0 + m;
      bool oka = gm < a.M && n0 + c < pr.N, okb = gm < a.M && kk0 + c < pr.K;
      cpa4z(&As[buf][m][c], oka ? pr.d + gm * pr.N + n0 + c : pr.d, oka);
      cpa4z(&Bs[buf][m][c], okb ? pr.x + gm * pr.K + kk0 + c : pr.x, okb);
    }
  };
#pragma unroll
  for (int c = 0; c < WG_ST - 1; ++c) {
    if (c < my_n) stage(c);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  for (int c = 0; c < my_n; ++c) {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(WG_ST - 2) : "memory");
    asm volatile("bar.sync %0, 128;\n" ::"r"(g + 1) : "memory");
    if (c + WG_ST - 1 < my_n) stage(c + WG_ST - 1);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    const int buf = c % WG_ST;
#pragma unroll
    for (int m = 0; m < WG_T; m += 8) {
      // A = d^T: element (row n, col m) = As[m][n]; B: element (row m, col k) = Bs[m][k]
      uint32_t ah[4], al[4];
      pvb::split_tf32(As[buf][m + ft][wr + fg], ah[0], al[0]);
      pvb::split_tf32(As[buf][m + ft][wr + fg + 8], ah[1], al[1]);
      pvb::split_tf32(As[buf][m + ft + 4][wr + fg], ah[2], al[2]);
      pvb::split_tf32(As[buf][m + ft + 4][wr + fg + 8], ah[3], al[3]);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint32_t bh[2], bl[2];
        pvb::split_tf32(Bs[buf][m + ft][wc + 8 * j + fg], bh[0], bl[0]);
        pvb::split_tf32(Bs[buf][m + ft + 4][wc + 8 * j + fg], bh[1], bl[1]);
        pvb::mma_tf32(acc[j], al, bh);      // small terms first
        pvb::mma_tf32(acc[j], ah, bl);
        pvb::mma_tf32(acc[j], ah, bh);
      }
    }
    if (do_bias && tid < WG_T) {
#pragma unroll
      for (int m = 0; m < WG_T; ++m) bsum += As[buf][m][tid];
    }
  }
  // cross-group reduction (aliases the stage buffers)
  asm volatile("cp.async.wait_all;\n" ::: "memory");
  __syncthreads();
  float* red = wg_smem;                       // [WG_G][32][33]
  float* bred = wg_smem + WG_G * WG_T * (WG_T + 1);   // [WG_G][32]
#pragma unroll
  for (int j = 0; j < 2; ++j) {     // accumulator fragment: (fg, 2 ft), (fg, 2 ft + 1), (fg + 8, 2 ft), (fg + 8, 2 ft + 1)
    float* r0 = red + (g * WG_T + wr + fg) * (WG_T + 1) + wc + 8 * j + 2 * ft;
    r0[0] = acc[j][0];
    r0[1] = acc[j][1];
    r0[8 * (WG_T + 1)] = acc[j][2];
    r0[8 * (WG_T + 1) + 1] = acc[j][3];
  }
  if (do_bias && tid < WG_T) bred[g * WG_T + tid] = bsum;
  __syncthreads();
  for (int idx = threadIdx.x; idx < WG_T * WG_T; idx += WG_THREADS) {
    int r = idx / WG_T, c = idx - r * WG_T;
    int n = n0 + r, k = kk0 + c;
    if (n >= pr.N || k >= pr.K) continue;
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < WG_G; ++q) v += red[(q * WG_T + r) * (WG_T + 1) + c];
    pr.dW[(size_t)n * pr.K + k] += v;
  }
  if (do_bias && threadIdx.x < WG_T && n0 + threadIdx.x < pr.N) {
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < WG_G; ++q) v += bred[q * WG_T + threadIdx.x];
    pr.db[n0 + threadIdx.x] += v;
  }
}

// ---- per-instance latent-side backward ---------------------------------------------------------
// up to LS_G CTAs, one instance at a time.  An instance pass is a chain of dependent L2 round trips (~8 us), so
// the kernel's time is passes per CTA x that: with 512 CTAs the 10 240 instances of the jiVAE benchmark took
// 20 passes = 174 us; one full wave of co-resident CTAs and the batched tile loop below halve that.
// (bounding the registers for 12 or 16 CTAs per SM was measured: no change -- the pass is not occupancy-bound)
constexpr int LS_OCC = 8;                    // CTAs per SM at 64 registers x 128 threads
constexpr int LS_G = 148 * LS_OCC, LS_T = 128, LS_MAXR = 40;    // one wave

__global__ void __launch_bounds__(LS_T, LS_OCC)
latent_side_bwd_kernel(pvb_fold_cfg cfg, const float* __restrict__ z, const float* __restrict__ cond,
                       const float* __restrict__ Wc, const float* __restrict__ Wz,
                       const float* __restrict__ gUv, const float* __restrict__ gUv_part, int N,
                       float* __restrict__ gz, float* __restrict__ gcond, float* __restrict__ part,
                       const float* __restrict__ eps, const float* __restrict__ sigma,
                       const float* __restrict__ s_pre, const float* __restrict__ w, float beta,
                       float* __restrict__ gmu, float* __restrict__ gs_pre, int64_t I) {
  extern __shared__ float sh[];  // [Hd*(ndim+1+LC)] weight-grad accumulators + reduction scratch
  const pvb::Split sp = pvb::split_of(cfg);
  const int Z = sp.off_c + cfg.latent_dim;
  const int LC = cfg.latent_dim + cfg.cond_dim;
  const int Hd = cfg.hidden;
  const int nd = cfg.ndim;
  const int per_h = nd + 1 + LC;
  float* acc = sh;                        // [Hd][per_h]
  float* red = sh + (size_t)Hd * per_h;   // [4][LS_MAXR]
  float* gzs = red + 4 * LS_MAXR;         // [LS_MAXR] dz of the current instance
  for (int k = threadIdx.x; k < Hd * per_h; k += blockDim.x) acc[k] = 0.f;
  __syncthreads();
  const int NR = 4 + LC;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  constexpr int TILE = PVB_TC_TILE, SLOTS = PVB_TC_MAX_SLOTS;

  for (int64_t i = blockIdx.x; i < I; i += gridDim.x) {
    const float* zi = z + i * Z;
    const pvb::Xform t = pvb::xform_of(cfg, sp, zi);
    const float c = t.c, sn = t.sn, s = t.s, dx = t.dx, dy = t.dy;
    float r[LS_MAXR];
#pragma unroll
    for (int k = 0; k < LS_MAXR; ++k) r[k] = 0.f;
    const int64_t t0 = gUv_part ? (i * N) / TILE : 0, t1 = gUv_part ? ((i + 1) * N - 1) / TILE : 0;
    for (int h = threadIdx.x; h < Hd; h += blockDim.x) {
      float g0, g1, gv;
      if (gUv_part) {
        // sum over the tiles touching instance i of its slot partial (fixed order)
        // (loads of eight tiles issued together, added in tile order: 33 tiles per instance at 64 x 64)
        // slot of instance i in tile tt = i - floor(tt * TILE / N): ONE division per instance (the first tile),
        // then the quotient is carried forward -- as a 64-bit division per tile and thread it was most of the
        // kernel's 4.9 k warp instructions per instance (ncu, jiVAE benchmark)
        g0 = g1 = gv = 0.f;
        const uint64_t a_first = (uint64_t)t0 * TILE;
        int64_t q_first = a_first < (1ull << 32) ? (int64_t)((uint32_t)a_first / (uint32_t)N) : (int64_t)(a_first / (uint64_t)N);
        int rem = (int)(a_first - (uint64_t)q_first * (uint64_t)N);
        for (int64_t tb = t0; tb <= t1; tb += 8) {
          float a0[8], a1[8], av[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int64_t tt = tb + u;
            if (tt <= t1) {
              const int slot = (int)(i - q_first);
              rem += TILE;                       // next tile: (tt + 1) * TILE = q * N + rem
              while (rem >= N) {
                rem -= N;
                ++q_first;
              }
              const float* p = gUv_part + tt * (SLOTS * 3 * Hd) + slot * 3 * Hd;
              a0[u] = __ldg(p + h);
              a1[u] = __ldg(p + Hd + h);
              av[u] = __ldg(p + 2 * Hd + h);
            } else {
              a0[u] = a1[u] = av[u] = 0.f;
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            g0 += a0[u];
            g1 += a1[u];
            gv += av[u];
          }
        }
      } else {
        const float* g = gUv + i * 3 * Hd;
        g0 = g[h]; g1 = g[Hd + h]; gv = g[2 * Hd + h];
      }
      float* a = acc + (size_t)h * per_h;
      if (nd == 2) {
        float w0 = Wc[h * 2], w1 = Wc[h * 2 + 1];
        a[0] += gv * dx + s * (g0 * c - g1 * sn);
        a[1] += gv * dy + s * (g0 * sn + g1 * c);
        r[0] += s * (g0 * (-w0 * sn + w1 * c) + g1 * (-w0 * c - w1 * sn));  // d/dphi
        r[1] += gv * w0;                                                      // d/d(dx)
        r[2] += gv * w1;                                                      // d/d(dy)
        r[3] += g0 * (w0 * c + w1 * sn) + g1 * (-w0 * sn + w1 * c);           // d/ds
      } else {
        float w0 = Wc[h];
        a[0] += gv * dx + g0;
        r[1] += gv * w0;
      }
      a[nd] += gv;  // bias
      for (int j = 0; j < cfg.latent_dim; ++j) {
        a[nd + 1 + j] += gv * zi[sp.off_c + j];
        if (4 + j < LS_MAXR) r[4 + j] += gv * Wz[h * LC + j];
      }
      for (int j = 0; j < cfg.cond_dim; ++j) {
        int jj = cfg.latent_dim + j;
        a[nd + 1 + jj] += gv * cond[i * cfg.cond_dim + j];
        if (4 + jj < LS_MAXR) r[4 + jj] += gv * Wz[h * LC + jj];
      }
    }
    __syncthreads();
    for (int k = 0; k < NR; ++k) {
      float v = pvb::warp_sum(r[k]);
      if (lane == 0) red[wid * LS_MAXR + k] = v;
    }
    if (threadIdx.x < Z) gzs[threadIdx.x] = 0.f;
    __syncthreads();
    if (threadIdx.x < NR) {
      int k = threadIdx.x;
      float v = 0.f;
      for (int ww = 0; ww < LS_T / 32; ++ww) v += red[ww * LS_MAXR + k];
      int zo = -1;
      if (k == 0) { if (sp.off_phi >= 0) zo = sp.off_phi; }
      else if (k == 1) { if (sp.off_t >= 0) { zo = sp.off_t; v *= cfg.dx_prior; } }
      else if (k == 2) { if (sp.off_t >= 0 && nd == 2) { zo = sp.off_t + 1; v *= cfg.dy_prior; } }
      else if (k == 3) { if (sp.off_s >= 0) { zo = sp.off_s; v *= cfg.sc_prior; } }
      else if (k - 4 < cfg.latent_dim) zo = sp.off_c + (k - 4);
      else if (gcond) gcond[i * cfg.cond_dim + (k - 4 - cfg.latent_dim)] = v;
      if (zo >= 0) {
        gz[i * Z + zo] = v;
        gzs[zo] = v;
      }
    }
    __syncthreads();
    if (gmu && threadIdx.x < Z) {
      // pvb_latent_bwd: loss = -sum w (ll + beta kl)
      const int64_t o = i * Z + threadIdx.x;
      const float bw = beta * (w ? w[i] : 1.f);
      const float zz = z[o];
      const float g = gzs[threadIdx.x] + bw * zz;
      gmu[o] = g;
      const float gsig = g * eps[o] - bw / sigma[o];
      gs_pre[o] = gsig * pvb::sigmoid_f(s_pre[o]);
    }
    __syncthreads();
  }
  float* p = part + (size_t)blockIdx.x * Hd * per_h;
  // partial layout: gWc[Hd][nd] | gbc[Hd] | gWz[Hd][LC]
  for (int k = threadIdx.x; k < Hd * per_h; k += blockDim.x) {
    int h = k / per_h, q = k % per_h;
    float v = acc[k];
    if (q < nd) p[h * nd + q] = v;
    else if (q == nd) p[Hd * nd + h] = v;
    else p[Hd * (nd + 1) + h * LC + (q - nd - 1)] = v;
  }
}

}  // namespace

extern "C" int pvb_mlp_tail_fwd(const pvb_mlp_tail_args* a, void* stream) {
  PVB_CHECK_ARG(a && a->M >= 0 && a->h_in, "pvb_mlp_tail_fwd: bad argument");
  PVB_CHECK_ARG(a->n_layers >= 0 && a->n_layers <= 3 && a->n_heads >= 1 && a->n_heads <= PVB_MLP_MAX_HEADS,
                "pvb_mlp_tail_fwd: up to 3 hidden layers and 1..3 heads");
  PVB_CHECK_ARG(a->w_in > 0 && a->w_in <= MAXW, "pvb_mlp_tail_fwd: input width %d > %d", a->w_in, MAXW);
  for (int l = 0; l < a->n_layers; ++l)
    PVB_CHECK_ARG(a->width[l] > 0 && a->width[l] <= MAXW && a->W[l] && a->h[l],
                  "pvb_mlp_tail_fwd: bad layer %d", l);
  for (int k = 0; k < a->n_heads; ++k)
    PVB_CHECK_ARG(a->hdim[k] > 0 && a->hdim[k] <= MAXHD && a->hW[k] && a->hout[k],
                  "pvb_mlp_tail_fwd: bad head %d", k);
  PVB_CHECK_ARG(a->act >= 0 && a->act <= PVB_ACT_SIGMOID, "pvb_mlp_tail_fwd: unknown activation");
  if (a->gauss) {
    PVB_CHECK_ARG(a->n_heads >= 2 && a->hdim[0] == a->hdim[1] && a->eps && a->sigma && a->z && a->kl,
                  "pvb_mlp_tail_fwd: gaussian head needs mu / s_pre heads and output buffers");
  }
  if (a->fold) {
    PVB_CHECK_ARG(a->gauss && a->Wc && a->bc && a->Uv, "pvb_mlp_tail_fwd: fold needs the sample");
    PVB_CHECK_ARG(a->cfg.cond_dim == 0 || a->cond, "pvb_mlp_tail_fwd: cond required");
    pvb::Split sp = pvb::split_of(a->cfg);
    PVB_CHECK_ARG(sp.off_c + a->cfg.latent_dim == a->hdim[0], "pvb_mlp_tail_fwd: latent split != Z");
  }
  if (a->M == 0) return 0;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(mlp_tail_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_BYTES);
    attr = true;
  }
  mlp_tail_fwd_kernel<<<pvb::cdiv(a->M, RB), MT, WS_BYTES, (cudaStream_t)stream>>>(*a);
  pvb::count_launch();
  return pvb::launch_status();
}

extern "C" int pvb_mlp_chain_bwd(const pvb_mlp_chain_args* a, void* stream) {
  PVB_CHECK_ARG(a && a->M >= 0, "pvb_mlp_chain_bwd: bad argument");
  PVB_CHECK_ARG(a->n_layers >= 1 && a->n_layers <= PVB_MLP_MAX_LAYERS && a->n_heads >= 1 &&
                    a->n_heads <= PVB_MLP_MAX_HEADS,
                "pvb_mlp_chain_bwd: 1..4 layers and 1..3 heads");
  for (int l = 0; l < a->n_layers; ++l)
    PVB_CHECK_ARG(a->width[l] > 0 && a->width[l] <= MAXW && a->h[l] && a->dpre[l] && (l == 0 || a->W[l]),
                  "pvb_mlp_chain_bwd: bad layer %d", l);
  for (int k = 0; k < a->n_heads; ++k)
    PVB_CHECK_ARG(a->hdim[k] > 0 && a->hdim[k] <= MAXHD && a->hW[k] && a->g[k],
                  "pvb_mlp_chain_bwd: bad head %d", k);
  PVB_CHECK_ARG(a->act != PVB_ACT_GELU || a->pre[0], "pvb_mlp_chain_bwd: gelu needs the pre-activations");
  if (a->M == 0) return 0;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(mlp_chain_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WS_BYTES);
    attr = true;
  }
  mlp_chain_bwd_kernel<<<pvb::cdiv(a->M, RB), MT, WS_BYTES, (cudaStream_t)stream>>>(*a);
  pvb::count_launch();
  return pvb::launch_status();
}

extern "C" int pvb_mlp_wgrad(const pvb_wgrad_problem* problems, int n_problems, int64_t M,
                             void* stream) {
  PVB_CHECK_ARG(problems && n_problems >= 1 && n_problems <= WG_MAXP && M >= 0,
                "pvb_mlp_wgrad: 1..8 problems");
  if (M == 0) return 0;
  WgradArgs a;
  a.M = M;
  a.n_prob = n_problems;
  int tiles = 0;
  for (int i = 0; i < n_problems; ++i) {
    PVB_CHECK_ARG(problems[i].d && problems[i].x && problems[i].dW && problems[i].N > 0 && problems[i].K > 0,
                  "pvb_mlp_wgrad: bad problem %d", i);
    a.p[i] = problems[i];
    a.tile0[i] = tiles;
    tiles += ((problems[i].N + WG_T - 1) / WG_T) * ((problems[i].K + WG_T - 1) / WG_T);
  }
  a.tile0[n_problems] = tiles;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(mlp_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
    attr = true;
  }
  mlp_wgrad_kernel<<<tiles, WG_THREADS, WG_SMEM, (cudaStream_t)stream>>>(a);
  pvb::count_launch();
  return pvb::launch_status();
}

extern "C" int pvb_latent_side_num_partials(int64_t I) { return (int)(I < LS_G ? (I > 0 ? I : 1) : LS_G); }

extern "C" int pvb_latent_side_bwd(const pvb_fold_cfg* cfg, const float* z, const float* cond,
                                   const float* Wc, const float* Wz, const float* gUv,
                                   const float* gUv_part, int N, float* gz, float* gcond,
                                   float* part, const float* eps, const float* sigma,
                                   const float* s_pre, const float* w, float beta, float* gmu,
                                   float* gs_pre, int64_t I, void* stream) {
  PVB_CHECK_ARG(cfg && z && Wc && (gUv || gUv_part) && gz && part && I >= 0,
                "pvb_latent_side_bwd: bad argument");
  PVB_CHECK_ARG(cfg->ndim == 1 || cfg->ndim == 2, "pvb_latent_side_bwd: ndim must be 1 or 2");
  PVB_CHECK_ARG(cfg->latent_dim + cfg->cond_dim + 4 <= LS_MAXR, "pvb_latent_side_bwd: latent_dim + cond_dim > 36");
  PVB_CHECK_ARG(!gUv_part || N >= 32, "pvb_latent_side_bwd: tile partials need N >= 32");
  PVB_CHECK_ARG(!gmu || (gs_pre && eps && sigma && s_pre), "pvb_latent_side_bwd: latent backward needs eps/sigma/s_pre");
  if (I == 0) return 0;
  int per_h = cfg->ndim + 1 + cfg->latent_dim + cfg->cond_dim;
  size_t smem = ((size_t)cfg->hidden * per_h + 5 * LS_MAXR) * sizeof(float);
  PVB_CHECK_ARG(smem <= 200 * 1024, "pvb_latent_side_bwd: hidden*(dims) too large");
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(latent_side_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_set = true;
  }
  const int grid = (int)(I < LS_G ? I : LS_G);
  latent_side_bwd_kernel<<<grid, LS_T, smem, (cudaStream_t)stream>>>(
      *cfg, z, cond, Wc, Wz, gUv, gUv_part, N, gz, gcond, part, eps, sigma, s_pre, w, beta, gmu,
      gs_pre, I);
  pvb::count_launch();
  return pvb::launch_status();
}
