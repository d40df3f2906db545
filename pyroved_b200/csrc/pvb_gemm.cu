// Generic fp32 SIMT GEMM + nn.Linear forward/backward built on it.
// Used for the (small) encoder / classifier layers and by the generic
// spatial-decoder path; the hot 128x128 decoder layers run on tcgen05
// (pvb_sdec_tc.cu).
#include <cooperative_groups.h>

#include "pvb_common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

// C[M,N] = epi( sum_k opA(A)[m,k] * opB(B)[k,n] )   row-major everywhere
//   TA == 0: A is [M,K] (lda)   TA == 1: A is [K,M] (lda)   (op = transpose)
//   TB == 0: B is [K,N] (ldb)   TB == 1: B is [N,K] (ldb)
// epilogue: + bias[n], activation, optional pre-activation store, accumulate.
// gridDim.z > 1: split-K, partial sums atomically added into C (C must be
// pre-initialised; bias/act must be off).
template <int TA, int TB>
__global__ void __launch_bounds__(NT)
sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
             float* __restrict__ Cpre, const float* __restrict__ bias, int64_t M, int N, int64_t K,
             int64_t lda, int64_t ldb, int64_t ldc, int act, int accumulate, int64_t k_chunk) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  const int64_t kb = (int64_t)blockIdx.z * k_chunk;
  const int64_t ke = (kb + k_chunk < K) ? kb + k_chunk : K;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t k0 = kb; k0 < ke; k0 += BK) {
    // ---- stage A tile (BM x BK) ----
#pragma unroll
    for (int it = 0; it < (BM * BK) / NT; ++it) {
      int idx = tid + it * NT;
      int m, k;
      if (TA == 0) { k = idx % BK; m = idx / BK; }   // k fastest (contiguous in memory)
      else         { m = idx % BM; k = idx / BM; }   // m fastest
      int64_t gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < ke) v = (TA == 0) ? A[gm * lda + gk] : A[gk * lda + gm];
      As[k][m] = v;
    }
    // ---- stage B tile (BK x BN) ----
#pragma unroll
    for (int it = 0; it < (BN * BK) / NT; ++it) {
      int idx = tid + it * NT;
      int n, k;
      if (TB == 0) { n = idx % BN; k = idx / BN; }
      else         { k = idx % BK; n = idx / BK; }
      int gn = n0 + n;
      int64_t gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < ke) v = (TB == 0) ? B[gk * ldb + gn] : B[(int64_t)gn * ldb + gk];
      Bs[k][n] = v;
    }
    __syncthreads();
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
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      float* c = C + gm * ldc + gn;
      if (gridDim.z > 1) {
        atomicAdd(c, v);
      } else {
        if (bias) v += bias[gn];
        if (Cpre) Cpre[gm * ldc + gn] = v;
        v = pvb::act_fwd(v, act);
        *c = accumulate ? *c + v : v;
      }
    }
  }
}

template <int TA, int TB>
int launch_sgemm_small(const float* A, const float* B, float* C, float* Cpre, const float* bias,
                       int M, int N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int act,
                       int accumulate, cudaStream_t st);

template <int TA, int TB>
int launch_sgemm(const float* A, const float* B, float* C, float* Cpre, const float* bias,
                 int64_t M, int N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int act,
                 int accumulate, int splits, cudaStream_t st) {
  if (M == 0 || N == 0) return 0;
  if (((M + BM - 1) / BM) * ((N + BN - 1) / BN) < 148 && M < (1 << 30) && ldc == N)
    return launch_sgemm_small<TA, TB>(A, B, C, Cpre, bias, (int)M, N, K, lda, ldb, ldc, act,
                                      accumulate, st);
  int64_t gy = (M + BM - 1) / BM;
  PVB_CHECK_ARG(gy <= 65535 * 16, "sgemm: M too large");
  int64_t k_chunk = K;
  if (splits > 1) {
    k_chunk = ((K + splits - 1) / splits + BK - 1) / BK * BK;
    splits = (int)((K + k_chunk - 1) / k_chunk);
  }
  // gridDim.y limit is 65535: fold extra row-blocks by looping on the host
  for (int64_t y0 = 0; y0 < gy; y0 += 65535) {
    int64_t ny = (gy - y0 < 65535) ? gy - y0 : 65535;
    dim3 grid((N + BN - 1) / BN, (unsigned)ny, splits > 1 ? splits : 1);
    int64_t moff = y0 * BM;
    const float* Ao = (TA == 0) ? A + moff * lda : A + moff;
    sgemm_kernel<TA, TB><<<grid, NT, 0, st>>>(Ao, B, C + moff * ldc, Cpre ? Cpre + moff * ldc : nullptr,
                                              bias, M - moff, N, K, lda, ldb, ldc, act, accumulate,
                                              k_chunk); pvb::count_launch();
  }
  return pvb::launch_status();
}

// ---- small-problem variant ------------------------------------------------------
// 32x32 output tiles, 128 threads (2x4 outputs each) and split-K over
// gridDim.z so that the encoder-sized GEMMs (M = batch = 512, N = 128,
// K = 784 ...) still fill the 148 SMs.  With splits > 1 partial sums are
// atomically added into C (pre-zeroed unless accumulating) and bias +
// activation run in a second, elementwise pass.
constexpr int SBM = 32, SBN = 32, SBK = 32, SNT = 128;
constexpr int SST = 4;   // cp.async stages: these GEMMs are latency-, not throughput-bound

__device__ __forceinline__ void cp_async_f32(float* smem_dst, const float* gsrc, bool pred) {
  // 4-byte async copy with zero fill when !pred (src-size 0)
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  int sz = pred ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}

template <int TA, int TB>
__global__ void __launch_bounds__(SNT)
sgemm_small_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                   float* __restrict__ Cpre, const float* __restrict__ bias, int M, int N,
                   int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int act, int accumulate,
                   int64_t k_chunk) {
  __shared__ float As[SST][SBK][SBM + 4];
  __shared__ float Bs[SST][SBK][SBN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 7, ty = tid >> 3;   // 8 column quads x 16 row pairs
  const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;
  const int64_t kb = (int64_t)blockIdx.z * k_chunk;
  const int64_t ke = (kb + k_chunk < K) ? kb + k_chunk : K;
  const int n_chunks = (int)((ke - kb + SBK - 1) / SBK);
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};

  auto stage = [&](int chunk) {
    const int buf = chunk % SST;
    const int64_t k0 = kb + (int64_t)chunk * SBK;
#pragma unroll
    for (int it = 0; it < (SBM * SBK) / SNT; ++it) {
      int idx = tid + it * SNT;
      int m, k;
      if (TA == 0) { k = idx % SBK; m = idx / SBK; }
      else         { m = idx % SBM; k = idx / SBM; }
      int gm = m0 + m;
      int64_t gk = k0 + k;
      bool ok = gm < M && gk < ke;
      const float* src = ok ? ((TA == 0) ? A + (int64_t)gm * lda + gk : A + gk * lda + gm) : A;
      cp_async_f32(&As[buf][k][m], src, ok);
    }
#pragma unroll
    for (int it = 0; it < (SBN * SBK) / SNT; ++it) {
      int idx = tid + it * SNT;
      int n, k;
      if (TB == 0) { n = idx % SBN; k = idx / SBN; }
      else         { k = idx % SBK; n = idx / SBK; }
      int gn = n0 + n;
      int64_t gk = k0 + k;
      bool ok = gn < N && gk < ke;
      const float* src = ok ? ((TB == 0) ? B + gk * ldb + gn : B + (int64_t)gn * ldb + gk) : B;
      cp_async_f32(&Bs[buf][k][n], src, ok);
    }
  };

  // prologue: SST-1 chunks in flight (empty commit groups keep the accounting uniform)
#pragma unroll
  for (int c = 0; c < SST - 1; ++c) {
    if (c < n_chunks) stage(c);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  for (int c = 0; c < n_chunks; ++c) {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(SST - 2) : "memory");
    __syncthreads();   // chunk c landed for everyone; buffer (c-1)%SST is free again
    if (c + SST - 1 < n_chunks) stage(c + SST - 1);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    const int buf = c % SST;
#pragma unroll
    for (int k = 0; k < SBK; ++k) {
      float2 a2 = *reinterpret_cast<const float2*>(&As[buf][k][ty * 2]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      acc[0][0] = fmaf(a2.x, b4.x, acc[0][0]); acc[0][1] = fmaf(a2.x, b4.y, acc[0][1]);
      acc[0][2] = fmaf(a2.x, b4.z, acc[0][2]); acc[0][3] = fmaf(a2.x, b4.w, acc[0][3]);
      acc[1][0] = fmaf(a2.y, b4.x, acc[1][0]); acc[1][1] = fmaf(a2.y, b4.y, acc[1][1]);
      acc[1][2] = fmaf(a2.y, b4.z, acc[1][2]); acc[1][3] = fmaf(a2.y, b4.w, acc[1][3]);
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int gm = m0 + ty * 2 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      float* c = C + (int64_t)gm * ldc + gn;
      if (gridDim.z > 1) {
        atomicAdd(c, v);
      } else {
        if (bias) v += bias[gn];
        if (Cpre) Cpre[(int64_t)gm * ldc + gn] = v;
        v = pvb::act_fwd(v, act);
        *c = accumulate ? *c + v : v;
      }
    }
  }
}

// ---- forward layer for small batches: C = act(A[M,K] B[N,K]^T + bias) ------------------------
// Both operands are K-contiguous: tiles are staged with 16-byte cp.async into K-major smem.
// A 32 x 32 output tile per CTA; the K loop is split over LKG = 4 thread groups of the same CTA (group g
// takes chunks g, g+4, ... through its own cp.async ring and named barrier); the four partial tiles
// are summed in a fixed order (deterministic), then bias + activation.
// The products run on the warp-level tensor-core path (mma.sync m16n8k8, 3 x TF32 = fp32-grade): as FFMA
// on a 2 x 4 register tile the loop issued 6 LDS.128 per 32 FMA and was bound by shared-memory bandwidth
// (12.7 us warm for the 784 -> 128 layer at batch 512).
constexpr int FLD = SBK + 4;
#ifndef PVB_LKG
#define PVB_LKG 4
#endif
constexpr int LKG = PVB_LKG;  // K groups per CTA
constexpr int LGS = 3;       // cp.async stages per group
constexpr int LS_THREADS = LKG * SNT;
// BN = 32, or 16 when 32-wide tiles would leave more than half of the SMs idle (the 784 -> 128 layer at
// batch 512: 64 tiles of 32 x 32, 128 of 32 x 16 -- the kernel is bound by its own instruction stream, so
// spreading it over twice the SMs is worth the repeated A fragments)
template <int BN>
constexpr int ls_smem_bytes() { return LKG * LGS * (SBM + BN) * FLD * 4; }
static_assert(LKG * SBM * (32 + 1) * 4 <= ls_smem_bytes<16>(), "reduction scratch aliases the stage buffers");

template <int BN>
__global__ void __launch_bounds__(LS_THREADS)
linear_small_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                    float* __restrict__ Cpre, const float* __restrict__ bias, int M, int N, int K,
                    int act) {
  extern __shared__ __align__(16) float ls_smem[];
  const int g = threadIdx.x >> 7, tid = threadIdx.x & 127;
  float (*As)[SBM][FLD] = reinterpret_cast<float (*)[SBM][FLD]>(ls_smem + g * LGS * (SBM + BN) * FLD);
  float (*Bs)[BN][FLD] = reinterpret_cast<float (*)[BN][FLD]>(ls_smem + g * LGS * (SBM + BN) * FLD +
                                                               LGS * SBM * FLD);
  // warp w of a group owns the 16 x (BN / 2) sub-tile (rows 16 (w & 1), columns (BN / 2) (w >> 1)) as
  // NJ = BN / 16 m16n8k8 MMAs
  constexpr int NJ = BN / 16;
  const int wq = tid >> 5, lane = tid & 31, fg = lane >> 2, ft = lane & 3;
  const int wr = (wq & 1) * 16, wc = (wq >> 1) * (BN / 2);
  const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * BN;
  const int n_chunks = (K + SBK - 1) / SBK;
  const int my_n = (n_chunks - g + LKG - 1) / LKG;   // chunks g, g + LKG, ...
  float acc[NJ][4] = {};
  auto stage = [&](int ci) {
    const int buf = ci % LGS;
    const int k0 = (g + ci * LKG) * SBK;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      int idx = tid + it * SNT;          // 256 16-byte pieces of the A tile, 8 * BN of the B tile
      int r = idx >> 3, q = idx & 7;
      int gk = k0 + 4 * q;
      {
        bool ok = (m0 + r < M) && (gk < K);   // K % 4 == 0: a piece is all-in or all-out
        unsigned d = (unsigned)__cvta_generic_to_shared(&As[buf][r][4 * q]);
        const float* src = ok ? A + (int64_t)(m0 + r) * K + gk : A;
        int sz = ok ? 16 : 0;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src), "r"(sz) : "memory");
      }
      if (idx < 8 * BN) {
        bool ok = (n0 + r < N) && (gk < K);
        unsigned d = (unsigned)__cvta_generic_to_shared(&Bs[buf][r][4 * q]);
        const float* src = ok ? B + (int64_t)(n0 + r) * K + gk : B;
        int sz = ok ? 16 : 0;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src), "r"(sz) : "memory");
      }
    }
  };
#pragma unroll
  for (int c = 0; c < LGS - 1; ++c) {
    if (c < my_n) stage(c);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  for (int c = 0; c < my_n; ++c) {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(LGS - 2) : "memory");
    asm volatile("bar.sync %0, 128;\n" ::"r"(g + 1) : "memory");   // this group's chunk c landed
    if (c + LGS - 1 < my_n) stage(c + LGS - 1);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    const int buf = c % LGS;
#pragma unroll
    for (int k = 0; k < SBK; k += 8) {
      // fragments straight from the K-major tiles: bank = (4 fg + ft) mod 32 with the 36-float row stride
      uint32_t ah[4], al[4];
      pvb::split_tf32(As[buf][wr + fg][k + ft], ah[0], al[0]);
      pvb::split_tf32(As[buf][wr + fg + 8][k + ft], ah[1], al[1]);
      pvb::split_tf32(As[buf][wr + fg][k + ft + 4], ah[2], al[2]);
      pvb::split_tf32(As[buf][wr + fg + 8][k + ft + 4], ah[3], al[3]);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        uint32_t bh[2], bl[2];
        pvb::split_tf32(Bs[buf][wc + 8 * j + fg][k + ft], bh[0], bl[0]);
        pvb::split_tf32(Bs[buf][wc + 8 * j + fg][k + ft + 4], bh[1], bl[1]);
        pvb::mma_tf32(acc[j], al, bh);      // small terms first
        pvb::mma_tf32(acc[j], ah, bl);
        pvb::mma_tf32(acc[j], ah, bh);
      }
    }
  }
  // cross-group reduction through smem (aliases the stage buffers: everyone is done with them)
  asm volatile("cp.async.wait_all;\n" ::: "memory");
  __syncthreads();
  float* red = ls_smem;   // [LKG][SBM][BN + 1]
#pragma unroll
  for (int j = 0; j < NJ; ++j) {     // accumulator fragment: (fg, 2 ft), (fg, 2 ft + 1), (fg + 8, 2 ft), (fg + 8, 2 ft + 1)
    float* r0 = red + (g * SBM + wr + fg) * (BN + 1) + wc + 8 * j + 2 * ft;
    r0[0] = acc[j][0];
    r0[1] = acc[j][1];
    r0[8 * (BN + 1)] = acc[j][2];
    r0[8 * (BN + 1) + 1] = acc[j][3];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < SBM * BN; idx += LS_THREADS) {
    int r = idx / BN, c = idx - r * BN;
    int gm = m0 + r, gn = n0 + c;
    if (gm >= M || gn >= N) continue;
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < LKG; ++q) v += red[(q * SBM + r) * (BN + 1) + c];
    if (bias) v += bias[gn];
    if (Cpre) Cpre[(int64_t)gm * N + gn] = v;
    C[(int64_t)gm * N + gn] = pvb::act_fwd(v, act);
  }
}

// second pass of a split-K forward: C = act(C + bias), optional pre-activation copy
__global__ void bias_act_kernel(float* __restrict__ C, float* __restrict__ Cpre,
                                const float* __restrict__ bias, int64_t M, int N, int act) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  float v = C[i];
  if (bias) v += bias[i % N];
  if (Cpre) Cpre[i] = v;
  C[i] = pvb::act_fwd(v, act);
}

template <int TA, int TB>
int launch_sgemm_small(const float* A, const float* B, float* C, float* Cpre, const float* bias,
                       int M, int N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int act,
                       int accumulate, cudaStream_t st) {
  const int tiles = ((M + SBM - 1) / SBM) * ((N + SBN - 1) / SBN);
  int64_t splits = (2 * 148 + tiles - 1) / tiles;
  int64_t max_splits = K / 64;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  // a forward layer with a fused epilogue: one launch beats split-K + memset + second pass once
  // a third of the SMs have a tile
  if (tiles >= 48 && !accumulate && (bias || Cpre || act != PVB_ACT_NONE)) splits = 1;
  int64_t k_chunk = K;
  if (splits > 1) {
    k_chunk = ((K + splits - 1) / splits + SBK - 1) / SBK * SBK;
    splits = (K + k_chunk - 1) / k_chunk;
  }
  dim3 grid((N + SBN - 1) / SBN, (M + SBM - 1) / SBM, (unsigned)splits);
  if (splits > 1) {
    PVB_CHECK_ARG(ldc == N, "sgemm: split-K needs a dense output");
    if (!accumulate) cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), st);
    sgemm_small_kernel<TA, TB><<<grid, SNT, 0, st>>>(A, B, C, nullptr, nullptr, M, N, K, lda, ldb,
                                                     ldc, PVB_ACT_NONE, 1, k_chunk);
    pvb::count_launch();
    if (bias || Cpre || act != PVB_ACT_NONE) {
      PVB_CHECK_ARG(!accumulate, "sgemm: split-K epilogue cannot accumulate");
      bias_act_kernel<<<pvb::cdiv((int64_t)M * N, 256), 256, 0, st>>>(C, Cpre, bias, M, N, act);
      pvb::count_launch();
    }
  } else {
    sgemm_small_kernel<TA, TB><<<grid, SNT, 0, st>>>(A, B, C, Cpre, bias, M, N, K, lda, ldb, ldc,
                                                     act, accumulate, k_chunk);
    pvb::count_launch();
  }
  return pvb::launch_status();
}

// dpre = dy * act'(y)   (elementwise; may run in place)
// (dy and dpre may alias: no __restrict__ on the pair)
__global__ void act_bwd_kernel(const float* dy, const float* __restrict__ y,
                               const float* __restrict__ pre, float* dpre, int64_t n, int act) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float p = pre ? pre[i] : 0.f;
    dpre[i] = dy[i] * pvb::act_grad(y[i], p, act);
  }
}

// db[n] += sum_m dpre[m,n]: each block sums a chunk of rows for 32 columns and
// adds its partial atomically.
constexpr int CS_ROWS = 2048;
__global__ void colsum_kernel(const float* __restrict__ a, float* __restrict__ out, int64_t M,
                              int N) {
  __shared__ float sm[32][33];
  int n = blockIdx.x * 32 + threadIdx.x;
  int64_t m_begin = (int64_t)blockIdx.y * CS_ROWS;
  int64_t m_end = m_begin + CS_ROWS < M ? m_begin + CS_ROWS : M;
  float s = 0.f;
  if (n < N)
    for (int64_t m = m_begin + threadIdx.y; m < m_end; m += 32) s += a[m * N + n];
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += sm[k][threadIdx.x];
    atomicAdd(out + n, t);
  }
}

// ---- skinny layers: N <= 8 outputs over a long K (VED's features2latent: 32768 -> 4) ----------------
// HBM-bound (one pass over x / dx), so no tiling: dot products per row forward, a broadcast outer
// product for dx, column-parallel sums over the rows for dW.
constexpr int SK_MAXN = 8;

__global__ void __launch_bounds__(256)
skinny_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ b,
                  float* __restrict__ y, float* __restrict__ pre, int N, int K, int act) {
  __shared__ float red[8][SK_MAXN];
  const int64_t row = blockIdx.x;
  const float4* xr = reinterpret_cast<const float4*>(x + row * K);
  float acc[SK_MAXN];
#pragma unroll
  for (int n = 0; n < SK_MAXN; ++n) acc[n] = 0.f;
  for (int k4 = threadIdx.x; k4 < K / 4; k4 += 256) {
    const float4 xv = __ldg(xr + k4);
#pragma unroll
    for (int n = 0; n < SK_MAXN; ++n)
      if (n < N) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(W + (int64_t)n * K) + k4);
        acc[n] = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, acc[n]))));
      }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int n = 0; n < SK_MAXN; ++n) {
    const float v = pvb::warp_sum(acc[n]);
    if (lane == 0) red[warp][n] = v;
  }
  __syncthreads();
  if (threadIdx.x < N) {
    float s = b ? b[threadIdx.x] : 0.f;
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    if (pre) pre[row * N + threadIdx.x] = s;
    y[row * N + threadIdx.x] = pvb::act_fwd(s, act);
  }
}

// same, SK_ROWS batch rows per block: the weight rows are fetched once per SK_ROWS rows of x instead
// of once per row (at M = 512, K = 32768 the one-row version moved 134 MB of weights through L2 for
// 67 MB of x)
constexpr int SK_ROWS = 4;
__global__ void __launch_bounds__(256)
skinny_fwd_rows_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ b,
                       float* __restrict__ y, float* __restrict__ pre, int64_t M, int N, int K, int act) {
  __shared__ float red[8][SK_ROWS][SK_MAXN];
  const int64_t row0 = (int64_t)blockIdx.x * SK_ROWS;
  float acc[SK_ROWS][SK_MAXN];
#pragma unroll
  for (int r = 0; r < SK_ROWS; ++r)
#pragma unroll
    for (int n = 0; n < SK_MAXN; ++n) acc[r][n] = 0.f;
  for (int k4 = threadIdx.x; k4 < K / 4; k4 += 256) {
    float4 xv[SK_ROWS];
#pragma unroll
    for (int r = 0; r < SK_ROWS; ++r)
      xv[r] = (row0 + r < M) ? __ldg(reinterpret_cast<const float4*>(x + (row0 + r) * K) + k4)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int n = 0; n < SK_MAXN; ++n)
      if (n < N) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(W + (int64_t)n * K) + k4);
#pragma unroll
        for (int r = 0; r < SK_ROWS; ++r)
          acc[r][n] = fmaf(xv[r].x, wv.x, fmaf(xv[r].y, wv.y, fmaf(xv[r].z, wv.z, fmaf(xv[r].w, wv.w, acc[r][n]))));
      }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < SK_ROWS; ++r)
#pragma unroll
    for (int n = 0; n < SK_MAXN; ++n) {
      const float v = pvb::warp_sum(acc[r][n]);
      if (lane == 0) red[warp][r][n] = v;
    }
  __syncthreads();
  if (threadIdx.x < SK_ROWS * N) {
    const int r = threadIdx.x / N, n = threadIdx.x - r * N;
    if (row0 + r < M) {
      float s2 = b ? b[n] : 0.f;
      for (int w = 0; w < 8; ++w) s2 += red[w][r][n];
      if (pre) pre[(row0 + r) * N + n] = s2;
      y[(row0 + r) * N + n] = pvb::act_fwd(s2, act);
    }
  }
}

// same rows-per-block kernel with the K range split over the SK_SPLIT CTAs of a thread-block cluster
// (blockIdx.y = cluster rank): 8x the CTAs, so the stream over x has ~250 KB in flight per SM instead of 32 KB
// (the one-CTA-per-4-rows version ran 128 CTAs of 8 warps, latency-bound at 1 TB/s: 68 us for 67 MB).  The partial
// [SK_ROWS][N] sums go to rank 0 through distributed shared memory and are added there in rank order
// (deterministic), then bias + activation.
constexpr int SK_SPLIT = 8;
__global__ void __launch_bounds__(256)
skinny_fwd_rows_cluster_kernel(const float* __restrict__ x, const float* __restrict__ W,
                               const float* __restrict__ b, float* __restrict__ y, float* __restrict__ pre,
                               int64_t M, int N, int K, int act) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float red[8][SK_ROWS][4];
  __shared__ float cpart[SK_SPLIT][SK_ROWS][4];     // rank 0's copy collects every rank's partial sums
  const unsigned rank = cluster.block_rank();
  // every CTA of the cluster must be running before another one writes into its shared memory: arrive now,
  // wait just before the remote store (the wait is then hidden behind this CTA's own stream over x)
  cluster.barrier_arrive();
  const int64_t row0 = (int64_t)blockIdx.x * SK_ROWS;
  const int K4 = K / 4, per = (K4 + SK_SPLIT - 1) / SK_SPLIT;
  const int k_lo = (int)rank * per, k_hi = min(K4, k_lo + per);
  float acc[SK_ROWS][4];
#pragma unroll
  for (int r = 0; r < SK_ROWS; ++r)
#pragma unroll
    for (int n = 0; n < 4; ++n) acc[r][n] = 0.f;
  for (int k4 = k_lo + (int)threadIdx.x; k4 < k_hi; k4 += 256) {
    float4 xv[SK_ROWS];
#pragma unroll
    for (int r = 0; r < SK_ROWS; ++r)
      xv[r] = (row0 + r < M) ? __ldg(reinterpret_cast<const float4*>(x + (row0 + r) * K) + k4)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int n = 0; n < 4; ++n)
      if (n < N) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(W + (int64_t)n * K) + k4);
#pragma unroll
        for (int r = 0; r < SK_ROWS; ++r)
          acc[r][n] = fmaf(xv[r].x, wv.x, fmaf(xv[r].y, wv.y, fmaf(xv[r].z, wv.z, fmaf(xv[r].w, wv.w, acc[r][n]))));
      }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < SK_ROWS; ++r)
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const float v = pvb::warp_sum(acc[r][n]);
      if (lane == 0) red[warp][r][n] = v;
    }
  __syncthreads();
  cluster.barrier_wait();
  if (threadIdx.x < SK_ROWS * 4) {
    const int r = threadIdx.x >> 2, n = threadIdx.x & 3;
    float s2 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s2 += red[w][r][n];
    float* dst = cluster.map_shared_rank(&cpart[0][0][0], 0);
    dst[(rank * SK_ROWS + r) * 4 + n] = s2;
  }
  cluster.sync();
  if (rank == 0 && threadIdx.x < SK_ROWS * 4) {
    const int r = threadIdx.x >> 2, n = threadIdx.x & 3;
    if (n < N && row0 + r < M) {
      float s2 = b ? b[n] : 0.f;
#pragma unroll
      for (int q = 0; q < SK_SPLIT; ++q) s2 += cpart[q][r][n];
      if (pre) pre[(row0 + r) * N + n] = s2;
      y[(row0 + r) * N + n] = pvb::act_fwd(s2, act);
    }
  }
}

// dx[m][k] (+)= sum_n g[m][n] W[n][k]
constexpr int SK_DXR = 8;
__global__ void __launch_bounds__(256)
skinny_dx_kernel(const float* __restrict__ g, const float* __restrict__ W, float* __restrict__ dx,
                 int64_t M, int N, int K, int accumulate) {
  // thread = one float4 column of SK_DXR consecutive rows: the weight columns are fetched once per SK_DXR
  // rows (per row, the N weight rows moved N x the bytes of dx through L2)
  const int K4 = K / 4;
  const int kb = (K4 + 255) / 256;
  const int k4 = (int)(blockIdx.x % kb) * 256 + threadIdx.x;
  const int64_t m0 = (int64_t)(blockIdx.x / kb) * SK_DXR;
  if (k4 >= K4) return;
  float4 wv[SK_MAXN];
#pragma unroll
  for (int n = 0; n < SK_MAXN; ++n)
    if (n < N) wv[n] = __ldg(reinterpret_cast<const float4*>(W + (int64_t)n * K) + k4);
#pragma unroll
  for (int r = 0; r < SK_DXR; ++r) {
    const int64_t m = m0 + r;
    if (m >= M) break;
    float4* op = reinterpret_cast<float4*>(dx + m * K) + k4;
    float4 o = accumulate ? *op : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int n = 0; n < SK_MAXN; ++n)
      if (n < N) {
        const float gv = __ldg(g + m * N + n);     // warp-uniform address: one broadcast load
        o.x = fmaf(gv, wv[n].x, o.x); o.y = fmaf(gv, wv[n].y, o.y);
        o.z = fmaf(gv, wv[n].z, o.z); o.w = fmaf(gv, wv[n].w, o.w);
      }
    *op = o;
  }
}

// dW[n][k] += sum_m g[m][n] x[m][k]  (rows split over blockIdx.y, atomics);  db[n] += sum_m g[m][n]

__global__ void __launch_bounds__(256)
skinny_dw_kernel(const float* __restrict__ g, const float* __restrict__ x, float* __restrict__ dW,
                 float* __restrict__ db, int64_t M, int N, int K, int rows_per_split) {
  const int K4 = K / 4;
  const int k4 = blockIdx.x * 256 + threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.y * rows_per_split;
  const int64_t m1 = m0 + rows_per_split < M ? m0 + rows_per_split : M;
  if (db && blockIdx.x == 0 && threadIdx.x < N) {
    float s = 0.f;
    for (int64_t m = m0; m < m1; ++m) s += g[m * N + threadIdx.x];
    atomicAdd(db + threadIdx.x, s);
  }
  if (k4 >= K4) return;
  float4 acc[SK_MAXN];
#pragma unroll
  for (int n = 0; n < SK_MAXN; ++n) acc[n] = make_float4(0.f, 0.f, 0.f, 0.f);
  // not unrolled on purpose: with 4 / 8 rows in flight per thread the kernel was measured SLOWER (44 / 48 us
  // against 36 at batch 512 x 32768): the CTAs of one row split then stream several rows at once
#pragma unroll 1
  for (int64_t m = m0; m < m1; ++m) {
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x + m * K) + k4);
#pragma unroll
    for (int n = 0; n < SK_MAXN; ++n)
      if (n < N) {
        const float gv = __ldg(g + m * N + n);     // warp-uniform address: one broadcast load
        acc[n].x = fmaf(gv, xv.x, acc[n].x); acc[n].y = fmaf(gv, xv.y, acc[n].y);
        acc[n].z = fmaf(gv, xv.z, acc[n].z); acc[n].w = fmaf(gv, xv.w, acc[n].w);
      }
  }
#pragma unroll
  for (int n = 0; n < SK_MAXN; ++n)
    if (n < N) {
      // one 16-byte reduction per weight row (dW rows are 16-byte aligned: skinny_ok, K % 4 == 0)
      float* o = dW + (int64_t)n * K + (int64_t)k4 * 4;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(acc[n].x), "f"(acc[n].y),
                   "f"(acc[n].z), "f"(acc[n].w)
                   : "memory");
    }
}

inline bool skinny_ok(int64_t M, int N, int K, const void* a, const void* b, const void* c) {
  return N <= SK_MAXN && K >= 2048 && (K % 4) == 0 && M > 0 && M * (int64_t)(K / 4) < (1ll << 40) &&
         (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) == 0;
}

}  // namespace

extern "C" int pvb_linear_fwd(const float* x, const float* W, const float* b, float* y, float* pre,
                              int64_t M, int N, int K, int act, void* stream) {
  PVB_CHECK_ARG(x && W && y && M >= 0 && N > 0 && K > 0, "pvb_linear_fwd: bad argument");
  PVB_CHECK_ARG(act >= 0 && act <= PVB_ACT_SIGMOID, "pvb_linear_fwd: unknown activation %d", act);
  if (skinny_ok(M, N, K, x, W, nullptr)) {
    if (M >= 4 * 148 / 2 && N <= 4 && K >= 4 * 256 * SK_SPLIT) {
      // SK_ROWS rows per cluster of SK_SPLIT CTAs along K
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)((M + SK_ROWS - 1) / SK_ROWS), SK_SPLIT, 1);
      cfg.blockDim = dim3(256, 1, 1);
      cfg.stream = (cudaStream_t)stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 1;
      at[0].val.clusterDim.y = SK_SPLIT;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, skinny_fwd_rows_cluster_kernel, x, W, b, y, pre, M, N, K, act);
    } else if (M >= 4 * 148 / 2 && N <= 4)     // enough rows to fill the SMs with SK_ROWS rows per block
      skinny_fwd_rows_kernel<<<(unsigned)((M + SK_ROWS - 1) / SK_ROWS), 256, 0, (cudaStream_t)stream>>>(
          x, W, b, y, pre, M, N, K, act);
    else
      skinny_fwd_kernel<<<(unsigned)M, 256, 0, (cudaStream_t)stream>>>(x, W, b, y, pre, N, K, act);
    pvb::count_launch();
    return pvb::launch_status();
  }
  // small batch, 16-byte aligned K-contiguous rows: pipelined single-launch kernel once a third
  // of the SMs get a tile
  if (M <= 8192 && (K % 4) == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)W % 16) == 0) {
    int tiles = (int)(((M + SBM - 1) / SBM) * ((N + SBN - 1) / SBN));
    // 16-wide tiles when the 32-wide ones cover half of the SMs or fewer (measured: 784 -> 128 at batch 512,
    // 64 -> 128 CTAs: 9.4 -> 6.8 us; at batch 1024, 128 -> 256 CTAs, the repeated A fragments cost more than the
    // extra SMs give: 10.5 -> 11.5 us, and 32.9 -> 39.9 us for the 4100 -> 128 layer)
    const bool narrow = tiles <= 74 && N % 16 == 0;
    if (narrow) tiles = (int)(((M + SBM - 1) / SBM) * ((N + 15) / 16));
    if (tiles >= 48 && tiles < 4 * 148) {
      if (M == 0) return 0;
      static bool attr = false;
      if (!attr) {
        cudaFuncSetAttribute(linear_small_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, ls_smem_bytes<32>());
        cudaFuncSetAttribute(linear_small_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, ls_smem_bytes<16>());
        attr = true;
      }
      if (narrow) {
        dim3 grid((N + 15) / 16, (unsigned)((M + SBM - 1) / SBM));
        linear_small_kernel<16><<<grid, LS_THREADS, ls_smem_bytes<16>(), (cudaStream_t)stream>>>(x, W, y, pre, b, (int)M,
                                                                                                 N, K, act);
      } else {
        dim3 grid((N + SBN - 1) / SBN, (unsigned)((M + SBM - 1) / SBM));
        linear_small_kernel<32><<<grid, LS_THREADS, ls_smem_bytes<32>(), (cudaStream_t)stream>>>(x, W, y, pre, b, (int)M,
                                                                                                 N, K, act);
      }
      pvb::count_launch();
      return pvb::launch_status();
    }
  }
  return launch_sgemm<0, 1>(x, W, y, pre, b, M, N, K, K, K, N, act, 0, 1, (cudaStream_t)stream);
}

extern "C" int pvb_linear_bwd(const float* x, const float* W, const float* y, const float* pre,
                              const float* dy, float* dpre_ws, float* dx, int dx_accumulate,
                              float* dW, float* db, int64_t M, int N, int K, int act,
                              void* stream) {
  PVB_CHECK_ARG(x && W && dy && dpre_ws && M >= 0 && N > 0 && K > 0, "pvb_linear_bwd: bad argument");
  PVB_CHECK_ARG(act == PVB_ACT_NONE || y, "pvb_linear_bwd: saved output required");
  PVB_CHECK_ARG(act != PVB_ACT_GELU || pre, "pvb_linear_bwd: gelu needs the pre-activation");
  cudaStream_t st = (cudaStream_t)stream;
  if (M == 0) return 0;
  const float* dpre = dy;
  if (act != PVB_ACT_NONE) {
    int64_t n = M * N;
    int blocks = (int)((n + 255) / 256 < 148 * 16 ? (n + 255) / 256 : 148 * 16);
    act_bwd_kernel<<<blocks, 256, 0, st>>>(dy, y, pre, dpre_ws, n, act); pvb::count_launch();
    dpre = dpre_ws;
  }
  if (skinny_ok(M, N, K, x, W, dx) && ((uintptr_t)dW & 15) == 0) {
    if (dx) {
      const int64_t blocks = ((K / 4 + 255) / 256) * ((M + SK_DXR - 1) / SK_DXR);
      skinny_dx_kernel<<<(unsigned)blocks, 256, 0, st>>>(dpre, W, dx, M, N, K, dx_accumulate);
      pvb::count_launch();
    }
    if (dW) {
      const int kblocks = (K / 4 + 255) / 256;
      int splits = (148 * 4 + kblocks - 1) / kblocks;
      if (splits > M) splits = (int)M;
      const int rows = (int)((M + splits - 1) / splits);
      dim3 grid(kblocks, (unsigned)((M + rows - 1) / rows));
      skinny_dw_kernel<<<grid, 256, 0, st>>>(dpre, x, dW, db, M, N, K, rows);
      pvb::count_launch();
    } else if (db) {
      dim3 blk(32, 32);
      dim3 grid((N + 31) / 32, (unsigned)((M + CS_ROWS - 1) / CS_ROWS));
      colsum_kernel<<<grid, blk, 0, st>>>(dpre, db, M, N); pvb::count_launch();
    }
    return pvb::launch_status();
  }
  int rc;
  if (dx) {
    // dx[M,K] = dpre[M,N] W[N,K]
    rc = launch_sgemm<0, 0>(dpre, W, dx, nullptr, nullptr, M, K, N, N, K, K, PVB_ACT_NONE,
                            dx_accumulate, 1, st);
    if (rc) return rc;
  }
  if (dW) {
    // dW[N,K] += dpre^T[N,M] x[M,K]   (reduction over M: split-K when long)
    int splits = 1;
    int64_t tiles = (int64_t)((N + BN - 1) / BN) * ((K + BM - 1) / BM);
    if (M >= 4096) {
      int64_t want = (148 * 4 + tiles - 1) / tiles;
      int64_t maxs = M / 512;
      splits = (int)(want < maxs ? want : maxs);
      if (splits < 1) splits = 1;
    }
    if (splits > 1) {
      rc = launch_sgemm<1, 0>(dpre, x, dW, nullptr, nullptr, N, K, M, N, K, K, PVB_ACT_NONE, 1,
                              splits, st);
    } else {
      rc = launch_sgemm<1, 0>(dpre, x, dW, nullptr, nullptr, N, K, M, N, K, K, PVB_ACT_NONE, 1, 1,
                              st);
    }
    if (rc) return rc;
  }
  if (db) {
    dim3 blk(32, 32);
    dim3 grid((N + 31) / 32, (unsigned)((M + CS_ROWS - 1) / CS_ROWS));
    colsum_kernel<<<grid, blk, 0, st>>>(dpre, db, M, N); pvb::count_launch();
  }
  return pvb::launch_status();
}
