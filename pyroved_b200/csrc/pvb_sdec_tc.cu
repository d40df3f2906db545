// Fused spatial-decoder step on tcgen05 tensor cores (sm_100a).
//
// One persistent CTA per SM walks 128-row tiles of the R = I*N pixel rows.
// For each tile everything stays on the SM:
//
//   h0  = tanh(U g + v)                 CUDA cores  -> smem A0 (fp16)
//   h1  = tanh(h0 W1^T + b1)            tcgen05 (TMEM acc) + epilogue -> smem A1
//   h2  = tanh(h1 W2^T + b2)            tcgen05 + epilogue (registers, smem Db)
//   l   = h2 . wo + bo ; log-lik ; dl   per-row epilogue (row == thread)
//   D2  = dl wo (1-h2^2)                -> smem Da
//   dh1 = D2 W2      ; dW2' += D2^T [h1|1] ; dwo += h2^T dl      tcgen05
//   D1  = dh1 (1-h1^2)                  -> smem Db
//   dh0 = D1 W1      ; dW1' += D1^T [h0|1]                       tcgen05
//   D0  = dh0 (1-h0^2)                  -> smem Da
//   dUv = D0^T [gx,gy,1 per sample slot]                          tcgen05
//
// Weight-gradient accumulators (dW1', dW2', dwo) live in TMEM for the whole
// kernel and are written once per CTA as partials; per-sample dUv goes out as
// per-tile partials (both reduced by tiny deterministic kernels).
// Operands are fp16 (values are tanh outputs / O(1) gradients), accumulation
// fp32.  Every operand tile uses the no-swizzle "row-chunk" layout of
// umma.cuh, which serves both as a K-major operand (K = columns) and as an
// MN-major operand (K = rows), so no transposed copies are ever made.
//
// Replaces sDecoderNet.forward / coord_latent.forward (nets/fc.py:189-237),
// the Bernoulli/Normal log_prob (utils/prob.py:25-29) and their autograd
// backward on the reference path.
#include "pvb_common.cuh"
#include "umma.cuh"

namespace {

constexpr int HD = 128;            // hidden width (fixed for this kernel)
constexpr int TILE = 128;          // rows per tile
constexpr int NTHREADS = 256;      // 8 warps: (lane quarter q = warp%4) x (column half = warp/4)
constexpr int MAX_SLOTS = 5;       // samples a tile can touch when N >= 32
constexpr int CHUNK = TILE * 16;   // bytes of one chunk-column (8 fp16 columns x 128 rows)

// ---- shared memory map (bytes) ------------------------------------------------
constexpr int SM_W1 = 0;                          // fp16 [128 out][128 in]
constexpr int SM_W2 = SM_W1 + 16 * CHUNK;         // fp16 [128 out][128 in]
constexpr int SM_A0 = SM_W2 + 16 * CHUNK;         // h0 : 16 chunk-columns + 2 ("ones", zeros)
constexpr int SM_A1 = SM_A0 + 18 * CHUNK;         // h1 : same
constexpr int SM_DA = SM_A1 + 18 * CHUNK;         // D2, later D0
constexpr int SM_DB = SM_DA + 16 * CHUNK;         // h2, later D1
constexpr int SM_G = SM_DB + 16 * CHUNK;          // [128][16] grid coords per sample slot
constexpr int SM_DL = SM_G + 2 * CHUNK;           // [128][16] dl in column 0
constexpr int SM_F32 = SM_DL + 2 * CHUNK;         // fp32 scratch, see below
constexpr int F_B1 = 0, F_B2 = 128, F_WO = 256;   // biases / out weights
constexpr int F_UV = 384;                         // [MAX_SLOTS][3][128]
constexpr int F_PART = F_UV + MAX_SLOTS * 3 * HD; // [2][128] partial dots
constexpr int F_RED = F_PART + 256;               // [8] block reduction
constexpr int F_END = F_RED + 8;
constexpr int SM_BAR = SM_F32 + F_END * 4;        // mbarrier (8 B) + tmem base (4 B)
constexpr int SMEM_BYTES = SM_BAR + 16;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");

// ---- tensor memory map (columns) -------------------------------------------------
constexpr uint32_t TM_ACC = 0;     // 128 cols : forward accumulators / dh
constexpr uint32_t TM_DW1 = 128;   // 144 cols : dW1 (128) | db1 (col 128)
constexpr uint32_t TM_DW2 = 272;   // 144 cols : dW2 | db2
constexpr uint32_t TM_DWO = 416;   // 16 cols  : dwo in column 0
constexpr uint32_t TM_DUV = 432;   // 16 cols  : per-tile dUv
constexpr int TM_COLS = 512;

__device__ __forceinline__ float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct Params {
  const float* Uv; const float* x; const float* w;
  const float* W1; const float* b1; const float* W2; const float* b2;
  const float* wo; const float* bo;
  float* rowll; float* loc; float* gUv_part; float* wgrad_part;
  int64_t R; int64_t B; int N; int H; int W; int ndim;
  int sampler; int sigmoid_d; float sig; int backward; int64_t tiles;
};

// fp32 [128][128] row-major global weights -> fp16 row-chunk tile in smem
__device__ __forceinline__ void stage_weight(const float* __restrict__ Wg, uint8_t* dst, int tid) {
  for (int idx = tid; idx < HD * (HD / 8); idx += NTHREADS) {
    int r = idx / (HD / 8), c8 = idx % (HD / 8);
    const float4* src = reinterpret_cast<const float4*>(Wg + r * HD + c8 * 8);
    float4 a = __ldg(src), b = __ldg(src + 1);
    __half2 h[4] = {__floats2half2_rn(a.x, a.y), __floats2half2_rn(a.z, a.w),
                    __floats2half2_rn(b.x, b.y), __floats2half2_rn(b.z, b.w)};
    *reinterpret_cast<uint4*>(dst + umma::tile_off(TILE, r, c8 * 8)) = *reinterpret_cast<uint4*>(h);
  }
}

// descriptors for a 128-row tile buffer at shared address `base`
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t base, int k16) {   // K = columns
  return umma::smem_desc(base + k16 * 2 * CHUNK, CHUNK, 128);
}
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t base, int k16) {  // K = rows
  return umma::smem_desc(base + k16 * 256, 128, CHUNK);
}

__global__ void __launch_bounds__(NTHREADS, 1) sdec_tc_kernel(Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* f32 = reinterpret_cast<float*>(smem + SM_F32);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, hf = warp >> 2;
  const int row = q * 32 + lane;        // tile row == TMEM lane owned by this thread
  const int col0 = hf * 64;             // this thread's 64 columns

  // ---- one-time setup ----------------------------------------------------------
  stage_weight(P.W1, smem + SM_W1, tid);
  stage_weight(P.W2, smem + SM_W2, tid);
  if (tid < HD) {
    f32[F_B1 + tid] = P.b1[tid];
    f32[F_B2 + tid] = P.b2[tid];
    f32[F_WO + tid] = P.wo[tid];
  }
  {
    // constant chunk-columns: [A0|A1] column 128 = 1 (bias column), 129..143 = 0;
    // DL columns 8..15 = 0
    uint4 ones = make_uint4(0x00003C00u, 0u, 0u, 0u);  // fp16 {1,0,0,0,0,0,0,0}
    uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    if (hf == 0) {
      *reinterpret_cast<uint4*>(smem + SM_A0 + umma::tile_off(TILE, row, 128)) = ones;
      *reinterpret_cast<uint4*>(smem + SM_A0 + umma::tile_off(TILE, row, 136)) = zero;
      *reinterpret_cast<uint4*>(smem + SM_DL + umma::tile_off(TILE, row, 8)) = zero;
    } else {
      *reinterpret_cast<uint4*>(smem + SM_A1 + umma::tile_off(TILE, row, 128)) = ones;
      *reinterpret_cast<uint4*>(smem + SM_A1 + umma::tile_off(TILE, row, 136)) = zero;
    }
  }
  if (warp == 0) umma::tmem_alloc<TM_COLS>(tmem_slot);
  if (tid == 0) {
    umma::mbar_init(bar, 1);
    umma::mbar_fence_init();
  }
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = *tmem_slot;
  const uint32_t tm_lane = tm + ((uint32_t)(q * 32) << 16);
  const float bo = P.bo[0];
  uint32_t phase = 0;
  bool first_tile = true;
  float dl_sum = 0.f;  // sum of dl over this thread's rows (column half 0 only) -> dbo

  const uint32_t sW1 = umma::smem_u32(smem + SM_W1), sW2 = umma::smem_u32(smem + SM_W2);
  const uint32_t sA0 = umma::smem_u32(smem + SM_A0), sA1 = umma::smem_u32(smem + SM_A1);
  const uint32_t sDA = umma::smem_u32(smem + SM_DA), sDB = umma::smem_u32(smem + SM_DB);
  const uint32_t sG = umma::smem_u32(smem + SM_G), sDL = umma::smem_u32(smem + SM_DL);
  constexpr uint32_t ID_FWD = umma::idesc_f16(128, 128, 0, 0);   // A K-major, B K-major
  constexpr uint32_t ID_DH = umma::idesc_f16(128, 128, 0, 1);    // A K-major, B MN-major
  constexpr uint32_t ID_DW = umma::idesc_f16(128, 144, 1, 1);    // both MN-major, N = 128+16
  constexpr uint32_t ID_N16 = umma::idesc_f16(128, 16, 1, 1);    // both MN-major, N = 16

  for (int64_t tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
    // ---- S0: rows of this tile, sample slots, first layer ------------------------
    const int64_t r_glob = tile * TILE + row;
    const bool valid = r_glob < P.R;
    const int64_t i_first = (tile * TILE) / P.N;
    const int64_t r_last = (tile * TILE + TILE - 1 < P.R - 1) ? tile * TILE + TILE - 1 : P.R - 1;
    const int n_slots = (int)(r_last / P.N - i_first) + 1;
    const int64_t inst = valid ? r_glob / P.N : i_first;
    const int pix = valid ? (int)(r_glob - inst * P.N) : 0;
    const int slot = (int)(inst - i_first);
    float gx = 0.f, gy = 0.f;
    pvb::grid_xy(pix, P.H, P.W, P.ndim, gx, gy);
    float xv = 0.f, wi = 1.f;
    if (valid) {
      if (P.x) xv = __ldg(P.x + (inst % P.B) * P.N + pix);
      if (P.w) wi = __ldg(P.w + inst);
    }
    for (int k = tid; k < n_slots * 3 * HD; k += NTHREADS)
      f32[F_UV + k] = __ldg(P.Uv + i_first * 3 * HD + k);
    __syncthreads();
    {
      const float* u = f32 + F_UV + slot * 3 * HD;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        __half2 hh[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int c = col0 + c8 * 8 + 2 * j;
          float a = fast_tanh(fmaf(u[c], gx, fmaf(u[HD + c], gy, u[2 * HD + c])));
          float b = fast_tanh(fmaf(u[c + 1], gx, fmaf(u[HD + c + 1], gy, u[2 * HD + c + 1])));
          hh[j] = valid ? __floats2half2_rn(a, b) : __floats2half2_rn(0.f, 0.f);
        }
        *reinterpret_cast<uint4*>(smem + SM_A0 + umma::tile_off(TILE, row, col0 + c8 * 8)) =
            *reinterpret_cast<uint4*>(hh);
      }
      if (P.backward) {
        // G[row][3*slot + {0,1,2}] = {gx, gy, 1}; this thread fills columns [8*hf, 8*hf+8)
        __half g8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          int n = hf * 8 + j;
          float v = 0.f;
          if (valid && n / 3 == slot) v = (n % 3 == 0) ? gx : (n % 3 == 1) ? gy : 1.f;
          g8[j] = __float2half_rn(v);
        }
        *reinterpret_cast<uint4*>(smem + SM_G + umma::tile_off(TILE, row, hf * 8)) =
            *reinterpret_cast<uint4*>(g8);
      }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    // ---- S1: ACC = h0 W1^T ---------------------------------------------------------
    if (tid == 0) {
      umma::fence_after_sync();
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma::mma_f16_ss(tm + TM_ACC, desc_kmajor(sA0, k), desc_kmajor(sW1, k), ID_FWD, k > 0);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, phase);
    phase ^= 1;
    umma::fence_after_sync();
    // ---- S2: h1 = tanh(ACC + b1) -> A1 ----------------------------------------------
#pragma unroll
    for (int cb = 0; cb < 2; ++cb) {
      float v[32];
      umma::tmem_ld32(tm_lane + TM_ACC + col0 + cb * 32, v);
      umma::tmem_ld_wait();
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        __half2 hh[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int c = col0 + cb * 32 + c8 * 8 + 2 * j;
          hh[j] = __floats2half2_rn(fast_tanh(v[c8 * 8 + 2 * j] + f32[F_B1 + c]),
                                    fast_tanh(v[c8 * 8 + 2 * j + 1] + f32[F_B1 + c + 1]));
        }
        *reinterpret_cast<uint4*>(smem + SM_A1 +
                                  umma::tile_off(TILE, row, col0 + cb * 32 + c8 * 8)) =
            *reinterpret_cast<uint4*>(hh);
      }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    // ---- S3: ACC = h1 W2^T -----------------------------------------------------------
    if (tid == 0) {
      umma::fence_after_sync();
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma::mma_f16_ss(tm + TM_ACC, desc_kmajor(sA1, k), desc_kmajor(sW2, k), ID_FWD, k > 0);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, phase);
    phase ^= 1;
    umma::fence_after_sync();
    // ---- S4: h2, logit, log-lik, dl, D2 --------------------------------------------------
    __half2 h2p[32];
    float pdot = 0.f;
#pragma unroll
    for (int cb = 0; cb < 2; ++cb) {
      float v[32];
      umma::tmem_ld32(tm_lane + TM_ACC + col0 + cb * 32, v);
      umma::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        int c = col0 + cb * 32 + 2 * j;
        __half2 h = __floats2half2_rn(fast_tanh(v[2 * j] + f32[F_B2 + c]),
                                      fast_tanh(v[2 * j + 1] + f32[F_B2 + c + 1]));
        h2p[cb * 16 + j] = h;
        float2 hf2 = __half22float2(h);
        pdot = fmaf(hf2.x, f32[F_WO + c], pdot);
        pdot = fmaf(hf2.y, f32[F_WO + c + 1], pdot);
      }
    }
    f32[F_PART + hf * TILE + row] = pdot;
    if (P.backward) {
      // h2 -> Db (A operand of the dwo GEMM)
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8)
        *reinterpret_cast<uint4*>(smem + SM_DB + umma::tile_off(TILE, row, col0 + c8 * 8)) =
            *reinterpret_cast<uint4*>(&h2p[c8 * 4]);
    }
    __syncthreads();
    const float logit = f32[F_PART + row] + f32[F_PART + TILE + row] + bo;
    float ll = 0.f, dnll = 0.f, locv;
    if (P.x) {
      pvb::obs_terms(logit, xv, P.sampler, P.sigmoid_d, P.sig, ll, dnll, locv);
    } else {
      locv = P.sigmoid_d ? pvb::sigmoid_f(logit) : logit;
    }
    const float dl = valid ? wi * dnll : 0.f;
    if (hf == 0 && valid) {
      if (P.rowll) P.rowll[r_glob] = ll;
      if (P.loc) P.loc[r_glob] = locv;
    }
    if (!P.backward) {
      // forward only: the next tile may reuse A0/A1 after this barrier
      umma::fence_before_sync();
      __syncthreads();
      continue;
    }
    if (hf == 0) {
      dl_sum += dl;
      __half d8[8];
      d8[0] = __float2half_rn(dl);
#pragma unroll
      for (int j = 1; j < 8; ++j) d8[j] = __float2half_rn(0.f);
      *reinterpret_cast<uint4*>(smem + SM_DL + umma::tile_off(TILE, row, 0)) =
          *reinterpret_cast<uint4*>(d8);
    }
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      __half2 dd[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int c = col0 + c8 * 8 + 2 * j;
        float2 h = __half22float2(h2p[c8 * 4 + j]);
        dd[j] = __floats2half2_rn(dl * f32[F_WO + c] * (1.f - h.x * h.x),
                                  dl * f32[F_WO + c + 1] * (1.f - h.y * h.y));
      }
      *reinterpret_cast<uint4*>(smem + SM_DA + umma::tile_off(TILE, row, col0 + c8 * 8)) =
          *reinterpret_cast<uint4*>(dd);
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    // ---- S5: dh1 = D2 W2 ; dW2' += D2^T [h1|1] ; dwo += h2^T dl ------------------------------
    if (tid == 0) {
      umma::fence_after_sync();
      const uint32_t acc = first_tile ? 0u : 1u;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma::mma_f16_ss(tm + TM_ACC, desc_kmajor(sDA, k), desc_mnmajor(sW2, k), ID_DH, k > 0);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma::mma_f16_ss(tm + TM_DW2, desc_mnmajor(sDA, k), desc_mnmajor(sA1, k), ID_DW,
                         (k > 0) ? 1u : acc);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma::mma_f16_ss(tm + TM_DWO, desc_mnmajor(sDB, k), desc_mnmajor(sDL, k), ID_N16,
                         (k > 0) ? 1u : acc);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, phase);
    phase ^= 1;
    umma::fence_after_sync();
    // ---- S6: D1 = dh1 (1 - h1^2) -> Db ----------------------------------------------------------
#pragma unroll
    for (int cb = 0; cb < 2; ++cb) {
      float v[32];
      umma::tmem_ld32(tm_lane + TM_ACC + col0 + cb * 32, v);
      umma::tmem_ld_wait();
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        const uint32_t off = umma::tile_off(TILE, row, col0 + cb * 32 + c8 * 8);
        uint4 hraw = *reinterpret_cast<const uint4*>(smem + SM_A1 + off);
        const __half2* hh = reinterpret_cast<const __half2*>(&hraw);
        __half2 dd[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 h = __half22float2(hh[j]);
          dd[j] = __floats2half2_rn(v[c8 * 8 + 2 * j] * (1.f - h.x * h.x),
                                    v[c8 * 8 + 2 * j + 1] * (1.f - h.y * h.y));
        }
        *reinterpret_cast<uint4*>(smem + SM_DB + off) = *reinterpret_cast<uint4*>(dd);
      }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    // ---- S7: dh0 = D1 W1 ; dW1' += D1^T [h0|1] -----------------------------------------------------
    if (tid == 0) {
      umma::fence_after_sync();
      const uint32_t acc = first_tile ? 0u : 1u;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma::mma_f16_ss(tm + TM_ACC, desc_kmajor(sDB, k), desc_mnmajor(sW1, k), ID_DH, k > 0);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma::mma_f16_ss(tm + TM_DW1, desc_mnmajor(sDB, k), desc_mnmajor(sA0, k), ID_DW,
                         (k > 0) ? 1u : acc);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, phase);
    phase ^= 1;
    umma::fence_after_sync();
    // ---- S8: D0 = dh0 (1 - h0^2) -> Da ---------------------------------------------------------------
#pragma unroll
    for (int cb = 0; cb < 2; ++cb) {
      float v[32];
      umma::tmem_ld32(tm_lane + TM_ACC + col0 + cb * 32, v);
      umma::tmem_ld_wait();
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        const uint32_t off = umma::tile_off(TILE, row, col0 + cb * 32 + c8 * 8);
        uint4 hraw = *reinterpret_cast<const uint4*>(smem + SM_A0 + off);
        const __half2* hh = reinterpret_cast<const __half2*>(&hraw);
        __half2 dd[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 h = __half22float2(hh[j]);
          dd[j] = __floats2half2_rn(v[c8 * 8 + 2 * j] * (1.f - h.x * h.x),
                                    v[c8 * 8 + 2 * j + 1] * (1.f - h.y * h.y));
        }
        *reinterpret_cast<uint4*>(smem + SM_DA + off) = *reinterpret_cast<uint4*>(dd);
      }
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    // ---- S9: dUv(tile) = D0^T G -------------------------------------------------------------------------
    if (tid == 0) {
      umma::fence_after_sync();
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma::mma_f16_ss(tm + TM_DUV, desc_mnmajor(sDA, k), desc_mnmajor(sG, k), ID_N16, k > 0);
      umma::commit(bar);
    }
    umma::mbar_wait(bar, phase);
    phase ^= 1;
    umma::fence_after_sync();
    // ---- S10: per-tile dUv partials: lane == hidden unit ---------------------------------------------------
    if (hf == 0) {
      float v[16];
      umma::tmem_ld16(tm_lane + TM_DUV, v);
      umma::tmem_ld_wait();
      float* dst = P.gUv_part + tile * (MAX_SLOTS * 3 * HD);
#pragma unroll
      for (int n = 0; n < MAX_SLOTS * 3; ++n)
        if (n < n_slots * 3) dst[n * HD + row] = v[n];   // unused slots are never read
    }
    first_tile = false;
    umma::fence_before_sync();
    __syncthreads();
  }

  // ---- weight-gradient partials of this CTA ---------------------------------------------------------------
  if (P.backward) {
    umma::fence_after_sync();
    float* out = P.wgrad_part + (size_t)blockIdx.x * PVB_TC_WGRAD_STRIDE;
    // layout: dW1[128][128] | db1[128] | dW2[128][128] | db2[128] | dwo[128] | dbo
    float* o_dW1 = out;
    float* o_db1 = out + HD * HD;
    float* o_dW2 = o_db1 + HD;
    float* o_db2 = o_dW2 + HD * HD;
    float* o_dwo = o_db2 + HD;
    float* o_dbo = o_dwo + HD;
    const bool any = !first_tile;  // false if this CTA processed no tile
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const uint32_t base = which == 0 ? TM_DW1 : TM_DW2;
      float* oW = which == 0 ? o_dW1 : o_dW2;
      float* ob = which == 0 ? o_db1 : o_db2;
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        float v[32];
        if (any) {
          umma::tmem_ld32(tm_lane + base + col0 + cb * 32, v);
          umma::tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        float4* dst = reinterpret_cast<float4*>(oW + row * HD + col0 + cb * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      if (hf == 1) {
        float v[16];
        if (any) {
          umma::tmem_ld16(tm_lane + base + 128, v);
          umma::tmem_ld_wait();
        } else {
          v[0] = 0.f;
        }
        ob[row] = v[0];
      }
    }
    if (hf == 0) {
      float v[16];
      if (any) {
        umma::tmem_ld16(tm_lane + TM_DWO, v);
        umma::tmem_ld_wait();
      } else {
        v[0] = 0.f;
      }
      o_dwo[row] = v[0];
    }
    float tot = pvb::block_sum(dl_sum, f32 + F_RED);
    if (tid == 0) o_dbo[0] = tot;
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<TM_COLS>(tm);
}

// gUv[i][c][h] = sum over the tiles touching instance i of its slot partial
__global__ void gather_gUv_kernel(const float* __restrict__ part, float* __restrict__ gUv,
                                  int64_t I, int N) {
  const int64_t i = blockIdx.x;
  const int64_t t0 = (i * N) / TILE, t1 = ((i + 1) * N - 1) / TILE;
  for (int k = threadIdx.x; k < 3 * HD; k += blockDim.x) {
    float s = 0.f;
    for (int64_t t = t0; t <= t1; ++t) {
      int slot = (int)(i - (t * TILE) / N);
      s += part[t * (MAX_SLOTS * 3 * HD) + slot * 3 * HD + k];
    }
    gUv[i * 3 * HD + k] = s;
  }
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

}  // namespace

extern "C" int pvb_has_tcgen05(void) { return 1; }

extern "C" int pvb_sdec_tc_sizes(int64_t I, int N, pvb_tc_sizes* out) {
  PVB_CHECK_ARG(out && I >= 0 && N >= 32, "pvb_sdec_tc_sizes: need N >= 32 pixels per instance");
  int64_t R = I * N;
  out->tiles = (R + TILE - 1) / TILE;
  int sms = sm_count();
  out->ctas = (int)(out->tiles < sms ? (out->tiles > 0 ? out->tiles : 1) : sms);
  out->gUv_part_floats = out->tiles * MAX_SLOTS * 3 * HD;
  out->wgrad_part_floats = (int64_t)out->ctas * PVB_TC_WGRAD_STRIDE;
  return 0;
}

extern "C" int pvb_sdec_tc_step(const float* Uv, const float* x, const float* w, const float* W1,
                                const float* b1, const float* W2, const float* b2, const float* wo,
                                const float* bo, float* rowll, float* loc, float* gUv_part,
                                float* wgrad_part, int64_t I, int64_t B, int H, int W, int ndim,
                                int sampler, int sigmoid_d, float decoder_sig, int backward,
                                void* stream) {
  PVB_CHECK_ARG(Uv && W1 && b1 && W2 && b2 && wo && bo, "pvb_sdec_tc_step: null weights");
  PVB_CHECK_ARG(ndim == 1 || ndim == 2, "pvb_sdec_tc_step: ndim must be 1 or 2");
  PVB_CHECK_ARG(I >= 0 && B > 0 && H > 0 && W > 0, "pvb_sdec_tc_step: bad dims");
  PVB_CHECK_ARG(sampler == PVB_SAMPLER_BERNOULLI || sampler == PVB_SAMPLER_GAUSSIAN,
                "pvb_sdec_tc_step: sampler %d not supported", sampler);
  PVB_CHECK_ARG(!backward || (x && gUv_part && wgrad_part), "pvb_sdec_tc_step: backward needs x and workspaces");
  PVB_CHECK_ARG(((uintptr_t)W1 % 16 == 0) && ((uintptr_t)W2 % 16 == 0), "pvb_sdec_tc_step: weights must be 16-byte aligned");
  PVB_CHECK_ARG(!backward || ((uintptr_t)wgrad_part % 16 == 0), "pvb_sdec_tc_step: wgrad_part must be 16-byte aligned");
  const int N = (ndim == 1) ? H : H * W;
  PVB_CHECK_ARG(N >= 32, "pvb_sdec_tc_step: need >= 32 pixels per instance");
  if (I == 0) return 0;
  pvb_tc_sizes s;
  pvb_sdec_tc_sizes(I, N, &s);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(sdec_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) { pvb::set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    attr = true;
  }
  Params P;
  P.Uv = Uv; P.x = x; P.w = w; P.W1 = W1; P.b1 = b1; P.W2 = W2; P.b2 = b2; P.wo = wo; P.bo = bo;
  P.rowll = rowll; P.loc = loc; P.gUv_part = gUv_part; P.wgrad_part = wgrad_part;
  P.R = I * N; P.B = B; P.N = N; P.H = H; P.W = (ndim == 1) ? 1 : W; P.ndim = ndim;
  P.sampler = sampler; P.sigmoid_d = sigmoid_d; P.sig = decoder_sig; P.backward = backward;
  P.tiles = s.tiles;
  sdec_tc_kernel<<<s.ctas, NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(P);
  pvb::count_launch();
  return pvb::launch_status();
}

extern "C" int pvb_sdec_tc_gather_gUv(const float* gUv_part, float* gUv, int64_t I, int N,
                                      void* stream) {
  PVB_CHECK_ARG(gUv_part && gUv && I >= 0 && N >= 32, "pvb_sdec_tc_gather_gUv: bad argument");
  if (I == 0) return 0;
  gather_gUv_kernel<<<(unsigned)I, 128, 0, (cudaStream_t)stream>>>(gUv_part, gUv, I, N);
  pvb::count_launch();
  return pvb::launch_status();
}
