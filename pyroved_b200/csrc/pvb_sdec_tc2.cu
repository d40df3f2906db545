// Fused spatial-decoder TRAINING step on tcgen05 tensor cores (sm_100a), two tiles in flight.
//
// Same arithmetic as pvb_sdec_tc.cu (see there for the per-tile algebra and the operand layout);
// what changes is the schedule.  In that kernel one 128-row tile walks the serial chain
//   S0 -> G1 -> S2 -> G2 -> S4 -> G3 -> S6 -> G4 -> S8 -> dUv
// (S = element-wise stage on the 16 epilogue warps, G = tensor-core GEMM) and the tensor pipe idles
// through every S while the MUFU pipe idles through every G (ncu r01b: tensor 27 %, XU 27 %).
// Here every CTA keeps TWO tiles in flight: the FORWARD half of tile i (S0 S2 S4: tanh-heavy) is
// interleaved with the BACKWARD half of tile i-1 (S6 S8: cheap element-wise work behind long GEMMs):
//
//   epilogue warps:  S0(i)  S6(i-1)  S2(i)  S8(i-1)  S4(i)   | S0(i+1) ...
//   tensor pipe   :  G3(i-1)+dW2'(i-1) | G1(i) | G4(i-1)+dW1'(i-1) | G2(i) | dUv(i-1) | G3(i)+...
//
// so each GEMM runs under an element-wise stage of the OTHER tile and each stage finds its
// accumulator ready.  Resources: one accumulator (a stage pulls its 32 columns into registers first
// and then releases it: ACCFREE), all chain operands in shared memory (SS form; the operand copy the
// weight-gradient GEMMs need anyway), four 32 KB operand buffers:
//   H0[2] (tile parity: h0, later D0 in place) | H1 (h1) | DA (D2, later D1)
// whose reuse is ordered by tcgen05.commit barriers.  h2 borrows the H0 buffer of the OTHER parity
// between dUv(i-1) and S0(i+1) (operand of dwo = h2^T dl).  Bias gradients and dwo come from N = 16
// MMAs against one small tile ODL whose column 0 is constant 1 and whose column 1 holds dl.
//
// TMEM (columns): ACC 128 | dW1 128 | dW2 128 | db1 16 | db2 16 | dUv 16 | dwo 16 = 448 of 512.
// SMEM: W1, W2 64 KB | operands 128 KB | G, ODL 8 KB | staging, biases, partials 24 KB = 224 KB.
//
// Replaces sDecoderNet.forward / coord_latent.forward (nets/fc.py:189-237), the Bernoulli / Normal
// log_prob (utils/prob.py:25-29) and their autograd backward on the reference path.
#include "pvb_common.cuh"
#include "pvb_sdec_tc.cuh"
#include "umma.cuh"

namespace {
using pvb_sdec::Params;

constexpr int HD = 128;
constexpr int TILE = 128;
constexpr int NEPI = 512;            // 16 epilogue warps: (lane quarter q = warp % 4) x (column group cg = warp / 4)
constexpr int NTHREADS = NEPI + 32;  // + one MMA-issuing warp
constexpr int MMA_WARP = NEPI / 32;
constexpr int MAX_SLOTS = 5;
static_assert(TILE == PVB_TC_TILE && MAX_SLOTS == PVB_TC_MAX_SLOTS, "pvb.h constants out of sync");
constexpr int CHUNK = TILE * 16;     // bytes of one chunk-column (8 fp16 columns x 128 rows)
constexpr int TILE_BYTES = 16 * CHUNK;

// ---- shared memory map (bytes) ------------------------------------------------
constexpr int SM_W1 = 0;
constexpr int SM_W2 = SM_W1 + TILE_BYTES;
constexpr int SM_H0 = SM_W2 + TILE_BYTES;        // two buffers (tile parity)
constexpr int SM_H1 = SM_H0 + 2 * TILE_BYTES;
constexpr int SM_DA = SM_H1 + TILE_BYTES;
constexpr int SM_G = SM_DA + TILE_BYTES;         // [128][16] grid coords per sample slot
constexpr int SM_ODL = SM_G + 2 * CHUNK;         // [128][16] column 0 = 1, column 1 = dl of the tile
constexpr int SM_F32 = SM_ODL + 2 * CHUNK;
constexpr int F_B1 = 0, F_B2 = 128, F_WO = 256;
constexpr int F_PART = 384;                      // [2][4][128] partial dots per column group, by tile parity
constexpr int F_RED = F_PART + 2 * 4 * TILE;     // [32] block reduction
constexpr int F_STG = F_RED + 32;                // two staging buffers
constexpr int UV_FLOATS = MAX_SLOTS * 3 * HD;
constexpr int G_UV = 0, G_X = UV_FLOATS, G_WI = G_X + TILE, G_GX = G_WI + TILE, G_GY = G_GX + TILE,
              G_GI = G_GY + TILE;
constexpr int STG_FLOATS = G_GI + TILE;
constexpr int F_END = F_STG + 2 * STG_FLOATS;
constexpr int SM_BAR = SM_F32 + F_END * 4;
constexpr int BAR_OP = 0;        // OP[5]: operand of S0, S6, S2, S8, S4 published (16 warps)
constexpr int OP_S0 = 0, OP_S6 = 1, OP_S2 = 2, OP_S8 = 3, OP_S4 = 4;
constexpr int BAR_ACCFREE = 5;   // accumulator pulled into registers by all 16 warps
constexpr int BAR_ACCDONE = 6;   // chain GEMM complete (commit)
constexpr int BAR_DW2 = 7;       // G3 + dW2' + db2 of a tile complete: DA (D2) and H1 reusable
constexpr int BAR_DW1 = 8;       // G4 + dW1' + db1 complete: DA (D1) and H0 (h0) reusable
constexpr int BAR_DUV = 9;       // dUv complete: H0 (D0), G and the dUv accumulator reusable
constexpr int BAR_DWO = 10;      // dwo complete: H0 (borrowed for h2) reusable
constexpr int BAR_W = 11;        // pre-packed weights landed (bulk copy, 64 KB)
constexpr int N_BARS = 12;
constexpr int SMEM_BYTES = SM_BAR + (N_BARS + 1) * 8;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");
static_assert(SM_F32 % 16 == 0 && SM_BAR % 8 == 0, "alignment");

// ---- tensor memory map (columns) -------------------------------------------------
constexpr uint32_t TM_ACC = 0;
constexpr uint32_t TM_DW1 = 128;
constexpr uint32_t TM_DW2 = 256;
constexpr uint32_t TM_DB1 = 384;
constexpr uint32_t TM_DB2 = 400;
constexpr uint32_t TM_DUV = 416;
constexpr uint32_t TM_DWO = 432;   // column 1 = dwo
constexpr int TM_COLS = 512;

#ifdef PVB_TC2_TRACE
// debug build only: per-stage timestamps of CTA 0 (epilogue warp 0 / MMA warp), tools/tc2_trace.py
__device__ long long g_trace2[2][64][16];
#define TRACE(role, ev)                                                                \
  do {                                                                                 \
    if (blockIdx.x == 0 && lane == 0 && (role == 1 || warp == 0) && i < 64)            \
      g_trace2[role][i][ev] = clock64();                                               \
  } while (0)
#else
#define TRACE(role, ev) do {} while (0)
#endif

// mbarrier wait with a watchdog: a protocol error traps (the launch fails with an error) after
// ~2 s instead of hanging the GPU.  The clock is read once per 4096 polls: no cost on the fast path.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = umma::smem_u32(bar);
  uint32_t done = 0, polls = 0;
  long long t0 = 0;
  for (;;) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if ((++polls & 4095u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) __trap();
    }
  }
}

__device__ __forceinline__ float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

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

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 512;\n" ::: "memory"); }

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(umma::smem_u32(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(umma::smem_u32(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// this thread's 32 accumulator columns: chunk-columns cg, 4+cg, 8+cg, 12+cg (8 columns each)
__device__ __forceinline__ void load_acc(uint32_t tm_lane, int cg, float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j) umma::tmem_ld8(tm_lane + TM_ACC + 8 * (4 * j + cg), v + 8 * j);
  umma::tmem_ld_wait();
}

__device__ __forceinline__ void lds8(const float* p, float* o) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}

__device__ __forceinline__ uint4 tanh8(const float* v, const float* bias) {
  float b[8];
  lds8(bias, b);
  __half2 hh[4];
#pragma unroll
  for (int e = 0; e < 4; ++e)
    hh[e] = __floats2half2_rn(fast_tanh(v[2 * e] + b[2 * e]), fast_tanh(v[2 * e + 1] + b[2 * e + 1]));
  return *reinterpret_cast<uint4*>(hh);
}

// 8 columns of  d = v * (1 - h^2)  (h: four fp16 pairs) -> four fp16 pairs
__device__ __forceinline__ uint4 dact8(const float* v, uint4 hraw) {
  const __half2* hh = reinterpret_cast<const __half2*>(&hraw);
  const __half2 one = __float2half2_rn(1.f);
  __half2 dd[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __half2 g = __hfma2(__hneg2(hh[j]), hh[j], one);   // exactly rounded 1 - h^2
    dd[j] = __hmul2(__floats2half2_rn(v[2 * j], v[2 * j + 1]), g);
  }
  return *reinterpret_cast<uint4*>(dd);
}

// ---- tile geometry without per-thread 64-bit divisions (as pvb_sdec_tc.cu) --------------------------
struct TileCursor {
  int64_t tile;
  int64_t i_first;   // instance containing the first row of the tile
  int ib_first;      // i_first % B  (row of the target image)
  int off;           // tile*TILE - i_first*N, in [0, N)
};
__device__ __forceinline__ void cursor_advance(TileCursor& c, const Params& P) {
  c.tile += gridDim.x;
  c.i_first += P.step_q;
  c.ib_first += P.step_qb;
  c.off += P.step_r;
  if (c.off >= P.N) { c.off -= P.N; ++c.i_first; ++c.ib_first; }
  if (c.ib_first >= (int)P.B) c.ib_first -= (int)P.B;
}
__device__ __forceinline__ void split_slot(int rem, int N, int& slot, int& pix) {
  slot = 0;
#pragma unroll
  for (int s = 0; s < MAX_SLOTS - 1; ++s)
    if (rem >= N) { rem -= N; ++slot; }
  pix = rem;
}

// asynchronous staging of the cursor's tile into `stg`: Uv rows (all threads), targets (column
// group 0), instance weights (group 1), row geometry (group 3)
__device__ __forceinline__ void stage_tile(const Params& P, float* stg, const TileCursor& c, int tid,
                                           int row, int cg) {
  const int64_t left = P.R - c.tile * TILE;
  const int last_row = left < TILE ? (int)left - 1 : TILE - 1;
  int last_slot, last_pix;
  split_slot(c.off + last_row, P.N, last_slot, last_pix);
  const int n_slots = last_slot + 1;
  if (tid < n_slots * (3 * HD / 4))
    cp_async16(stg + G_UV + tid * 4, P.Uv + c.i_first * 3 * HD + tid * 4);
  const bool valid = row <= last_row;
  int slot, pix;
  split_slot(c.off + (valid ? row : 0), P.N, slot, pix);
  if (cg == 0) {
    if (valid) {
      int ib = c.ib_first + slot;
      while (ib >= (int)P.B) ib -= (int)P.B;
      cp_async4(stg + G_X + row, P.x + (int64_t)ib * P.N + pix);
    }
  } else if (cg == 1) {
    if (valid && P.w) cp_async4(stg + G_WI + row, P.w + c.i_first + slot);
  } else if (cg == 3) {
    float gx = 0.f, gy = 0.f;
    pvb::grid_xy(pix, P.H, P.W, P.ndim, gx, gy);
    stg[G_GX + row] = gx;
    stg[G_GY + row] = gy;
    reinterpret_cast<int*>(stg)[G_GI + row] = slot | ((int)valid << 8) | (n_slots << 16);
  }
}

__device__ __forceinline__ void store_chunk(uint8_t* tile, int row, int cg, int j, uint4 v) {
  *reinterpret_cast<uint4*>(tile + umma::tile_off(TILE, row, 8 * (4 * j + cg))) = v;
}
// this warp's shared-memory operand stores -> visible to the tensor core, then one arrive per warp
__device__ __forceinline__ void publish_smem(uint64_t* bar) {
  umma::fence_proxy_async();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) umma::mbar_arrive(bar);
}
// this warp's tcgen05.ld of the accumulator are complete -> the MMA warp may overwrite it
__device__ __forceinline__ void release_acc(uint64_t* bars) {
  umma::fence_before_sync();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) umma::mbar_arrive(bars + BAR_ACCFREE);
}

// per-tile dUv partials: lane == hidden unit, 3 columns per sample slot
__device__ __forceinline__ void write_duv(const Params& P, uint32_t tm_lane, int row, int64_t tile,
                                          int n_slots) {
  float v[16];
  umma::tmem_ld16(tm_lane + TM_DUV, v);
  umma::tmem_ld_wait();
  float* dst = P.gUv_part + tile * (MAX_SLOTS * 3 * HD);
#pragma unroll
  for (int n = 0; n < MAX_SLOTS * 3; ++n)
    if (n < n_slots * 3) dst[n * HD + row] = v[n];   // unused slots are never read
}

__global__ void __launch_bounds__(NTHREADS, 1) sdec_tc2_kernel(Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* f32 = reinterpret_cast<float*>(smem + SM_F32);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + N_BARS * 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, cg = (warp >> 2) & 3;
  const int row = q * 32 + lane;        // tile row == TMEM lane owned by this thread
  // tiles of this CTA: blockIdx.x, + grid, ...  (grid <= tiles: at least one)
  const int n_local = (int)((P.tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);

  // ---- one-time setup ----------------------------------------------------------
  if (!P.Wp) {       // otherwise: two bulk copies through the TMA engine, issued below
    stage_weight(P.W1, smem + SM_W1, tid);
    stage_weight(P.W2, smem + SM_W2, tid);
  }
  if (tid < HD) {
    f32[F_B1 + tid] = P.b1[tid];
    f32[F_B2 + tid] = P.b2[tid];
    f32[F_WO + tid] = P.wo[tid];
  }
  if (tid < NEPI) {
    uint4 ones = make_uint4(0x00003C00u, 0u, 0u, 0u);  // fp16 {1,0,0,0,0,0,0,0}
    uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    if (cg == 0) {
      *reinterpret_cast<uint4*>(smem + SM_ODL + umma::tile_off(TILE, row, 0)) = ones;
    } else if (cg == 1) {
      *reinterpret_cast<uint4*>(smem + SM_ODL + umma::tile_off(TILE, row, 8)) = zero;
    } else if (cg == 2) {
      // default targets 0 / weights 1 in both staging buffers (rows the staging does not touch)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        f32[F_STG + b * STG_FLOATS + G_X + row] = 0.f;
        f32[F_STG + b * STG_FLOATS + G_WI + row] = 1.f;
      }
    }
  }
  if (warp == MMA_WARP) umma::tmem_alloc<TM_COLS>(tmem_slot);
  if (tid == 0) {
#pragma unroll
    for (int j = 0; j < 5; ++j) umma::mbar_init(bars + BAR_OP + j, NEPI / 32);
    umma::mbar_init(bars + BAR_ACCFREE, NEPI / 32);
    umma::mbar_init(bars + BAR_ACCDONE, 1);
    umma::mbar_init(bars + BAR_DW2, 1);
    umma::mbar_init(bars + BAR_DW1, 1);
    umma::mbar_init(bars + BAR_DUV, 1);
    umma::mbar_init(bars + BAR_DWO, 1);
    umma::mbar_init(bars + BAR_W, 1);
    umma::mbar_fence_init();
    if (P.Wp) pvb_sdec::bulk_load_weights(P.Wp, smem + SM_W1, smem + SM_W2, bars + BAR_W);
  }
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = *tmem_slot;
  const uint32_t tm_lane = tm + ((uint32_t)(q * 32) << 16);
  float dl_sum = 0.f;          // sum of dl over this thread's rows (column group 0 only) -> dbo

  if (warp == MMA_WARP) {
    // =========================== MMA issuer =====================================
    const uint32_t sW1 = umma::smem_u32(smem + SM_W1), sW2 = umma::smem_u32(smem + SM_W2);
    const uint32_t sH0 = umma::smem_u32(smem + SM_H0), sH1 = umma::smem_u32(smem + SM_H1);
    const uint32_t sDA = umma::smem_u32(smem + SM_DA), sG = umma::smem_u32(smem + SM_G);
    const uint32_t sODL = umma::smem_u32(smem + SM_ODL);
    constexpr uint32_t ID_FWD = umma::idesc_f16(128, 128, 0, 0);   // A K-major, B K-major
    constexpr uint32_t ID_DH = umma::idesc_f16(128, 128, 0, 1);    // A K-major, B MN-major
    constexpr uint32_t ID_DW = umma::idesc_f16(128, 128, 1, 1);    // both MN-major
    constexpr uint32_t ID_N16 = umma::idesc_f16(128, 16, 1, 1);    // both MN-major, N = 16
    uint32_t ph_op[5] = {0, 0, 0, 0, 0};
    uint32_t ph_free = 0;
    if (P.Wp) mbar_wait(bars + BAR_W, 0);   // both weight tiles have landed
    for (int i = 0; i <= n_local; ++i) {
      const bool F = i < n_local, Bk = i >= 1;
      const uint32_t sH0f = sH0 + (uint32_t)(i & 1) * TILE_BYTES;         // h0 of tile i
      const uint32_t sH0b = sH0 + (uint32_t)((i - 1) & 1) * TILE_BYTES;   // h0 / D0 of tile i-1
      if (F) {
        // ---- G1(i): ACC = h0 W1^T ----
        mbar_wait(bars + BAR_OP + OP_S0, ph_op[OP_S0]); ph_op[OP_S0] ^= 1;
        TRACE(1, 0);
        if (Bk) { mbar_wait(bars + BAR_ACCFREE, ph_free); ph_free ^= 1; }   // S6(i-1) holds G3's result
        TRACE(1, 1);
        umma::fence_after_sync();
        if (umma::elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma::mma_f16_ss(tm + TM_ACC, desc_kmajor(sH0f, k), desc_kmajor(sW1, k), ID_FWD, k > 0);
          umma::commit(bars + BAR_ACCDONE);
        }
        __syncwarp();
      }
      if (Bk) {
        // ---- G4(i-1): ACC = D1 W1 ; dW1' += D1^T h0 ; db1 += D1^T 1 ----
        mbar_wait(bars + BAR_OP + OP_S6, ph_op[OP_S6]); ph_op[OP_S6] ^= 1;
        TRACE(1, 2);
        mbar_wait(bars + BAR_ACCFREE, ph_free); ph_free ^= 1;     // S2(i) (drain: S6(i-1))
        TRACE(1, 3);
        umma::fence_after_sync();
        if (umma::elect_one()) {
          const uint32_t accw = (i - 1) > 0 ? 1u : 0u;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma::mma_f16_ss(tm + TM_ACC, desc_kmajor(sDA, k), desc_mnmajor(sW1, k), ID_DH, k > 0);
          umma::commit(bars + BAR_ACCDONE);
#ifndef PVB_TC2_NO_DW_MMA
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma::mma_f16_ss(tm + TM_DW1, desc_mnmajor(sDA, k), desc_mnmajor(sH0b, k), ID_DW,
                             (k > 0) ? 1u : accw);
#endif
#ifndef PVB_TC2_NO_BIAS_MMA     // (timing experiments only)
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma::mma_f16_ss(tm + TM_DB1, desc_mnmajor(sDA, k), desc_mnmajor(sODL, k), ID_N16,
                             (k > 0) ? 1u : accw);
#endif
          umma::commit(bars + BAR_DW1);
        }
        __syncwarp();
      }
      if (F) {
        // ---- G2(i): ACC = h1 W2^T ----
        mbar_wait(bars + BAR_OP + OP_S2, ph_op[OP_S2]); ph_op[OP_S2] ^= 1;
        TRACE(1, 4);
        mbar_wait(bars + BAR_ACCFREE, ph_free); ph_free ^= 1;     // S8(i-1) (i == 0: S2(0))
        TRACE(1, 5);
        umma::fence_after_sync();
        if (umma::elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma::mma_f16_ss(tm + TM_ACC, desc_kmajor(sH1, k), desc_kmajor(sW2, k), ID_FWD, k > 0);
          umma::commit(bars + BAR_ACCDONE);
        }
        __syncwarp();
      }
      if (Bk) {
        // ---- dUv(i-1) = D0^T G ----
        mbar_wait(bars + BAR_OP + OP_S8, ph_op[OP_S8]); ph_op[OP_S8] ^= 1;
        TRACE(1, 6);
        umma::fence_after_sync();
        if (umma::elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma::mma_f16_ss(tm + TM_DUV, desc_mnmajor(sH0b, k), desc_mnmajor(sG, k), ID_N16, k > 0);
          umma::commit(bars + BAR_DUV);
        }
        __syncwarp();
      }
      if (F) {
        // ---- dwo += h2^T dl ; G3(i): ACC = D2 W2 ; dW2' += D2^T h1 ; db2 += D2^T 1 ----
        // (S4 pulled the accumulator into registers before it published D2: no ACCFREE wait)
        mbar_wait(bars + BAR_OP + OP_S4, ph_op[OP_S4]); ph_op[OP_S4] ^= 1;
        TRACE(1, 7);
        umma::fence_after_sync();
        if (umma::elect_one()) {
          const uint32_t accw = i > 0 ? 1u : 0u;
          // h2 sits in the H0 buffer of the other parity, which S0(i+1) rewrites: first in line
#ifndef PVB_TC2_NO_DWO_MMA
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma::mma_f16_ss(tm + TM_DWO, desc_mnmajor(sH0b, k), desc_mnmajor(sODL, k), ID_N16,
                             (k > 0) ? 1u : accw);
#endif
          umma::commit(bars + BAR_DWO);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma::mma_f16_ss(tm + TM_ACC, desc_kmajor(sDA, k), desc_mnmajor(sW2, k), ID_DH, k > 0);
          umma::commit(bars + BAR_ACCDONE);
#ifndef PVB_TC2_NO_DW_MMA
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma::mma_f16_ss(tm + TM_DW2, desc_mnmajor(sDA, k), desc_mnmajor(sH1, k), ID_DW,
                             (k > 0) ? 1u : accw);
#endif
#ifndef PVB_TC2_NO_BIAS_MMA
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma::mma_f16_ss(tm + TM_DB2, desc_mnmajor(sDA, k), desc_mnmajor(sODL, k), ID_N16,
                             (k > 0) ? 1u : accw);
#endif
          umma::commit(bars + BAR_DW2);
        }
        __syncwarp();
      }
    }
  } else {
    // =========================== epilogue warps ====================================
    const float bo = P.bo[0];
    uint32_t ph_acc = 0, ph_dw2 = 0, ph_dw1 = 0, ph_duv = 0, ph_dwo = 0;
    TileCursor cur_c;
    cur_c.tile = blockIdx.x;
    cur_c.i_first = (cur_c.tile * TILE) / P.N;     // the only 64-bit divisions of the kernel
    cur_c.off = (int)(cur_c.tile * TILE - cur_c.i_first * P.N);
    cur_c.ib_first = (int)(cur_c.i_first % P.B);
    stage_tile(P, f32 + F_STG, cur_c, tid, row, cg);
    cp_async_wait_all();
    epi_bar();
    float p_gx = 0.f, p_gy = 0.f;                     // geometry of backward tile i-1
    int p_gi = 0;
    uint32_t slots_hist = 0;                          // n_slots of the last tiles, 4 bits each
    float v[32];
    for (int i = 0; i <= n_local; ++i) {
      const bool F = i < n_local, Bk = i >= 1;
      uint8_t* H0f = smem + SM_H0 + (i & 1) * TILE_BYTES;
      uint8_t* H0b = smem + SM_H0 + ((i - 1) & 1) * TILE_BYTES;
      if (i >= 2) {
        // dUv(i-2) complete: H0f (held D0(i-2)), G and the dUv accumulator are free again
        mbar_wait(bars + BAR_DUV, ph_duv);
        ph_duv ^= 1;
        umma::fence_after_sync();
        if (cg == 1)
          write_duv(P, tm_lane, row, (int64_t)blockIdx.x + (int64_t)(i - 2) * gridDim.x,
                    (int)((slots_hist >> 4) & 0xf));
      }
      TRACE(0, 0);
      // staging of tile i+1 (its buffer was last read in S0(i-1), before the S4(i-1) barrier)
      if (F) {
        TileCursor nxt_c = cur_c;
        cursor_advance(nxt_c, P);
        if (nxt_c.tile < P.tiles)
          stage_tile(P, f32 + F_STG + ((i + 1) & 1) * STG_FLOATS, nxt_c, tid, row, cg);
      }
      // ---- S0(i): h0 = tanh(U g + v) -> H0f ---------------------------------------------------------
      if (F) {
        const float* stg = f32 + F_STG + (i & 1) * STG_FLOATS;
        const int gi = reinterpret_cast<const int*>(stg)[G_GI + row];
        const float gx = stg[G_GX + row], gy = stg[G_GY + row];
        const bool valid = (gi >> 8) & 1;
        slots_hist = (slots_hist << 4) | (uint32_t)(gi >> 16);
        const float* u = stg + G_UV + (gi & 0xff) * 3 * HD;
        if (Bk) {
          // dwo(i-1) has read h2(i-1), which borrowed H0f (issued first behind OP_S4: long done)
          mbar_wait(bars + BAR_DWO, ph_dwo);
          ph_dwo ^= 1;
        }
        TRACE(0, 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c0 = 8 * (4 * j + cg);
          float ux[8], uy[8], uc[8];
          lds8(u + c0, ux);
          lds8(u + HD + c0, uy);
          lds8(u + 2 * HD + c0, uc);
          __half2 hh[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float a = fast_tanh(fmaf(ux[2 * e], gx, fmaf(uy[2 * e], gy, uc[2 * e])));
            float b = fast_tanh(fmaf(ux[2 * e + 1], gx, fmaf(uy[2 * e + 1], gy, uc[2 * e + 1])));
            hh[e] = valid ? __floats2half2_rn(a, b) : __floats2half2_rn(0.f, 0.f);
          }
          store_chunk(H0f, row, cg, j, *reinterpret_cast<uint4*>(hh));
        }
        publish_smem(bars + BAR_OP + OP_S0);
        TRACE(0, 2);
      }
      // ---- S6(i-1): D1 = dh1 (1 - h1^2) -> DA ---------------------------------------------------------
      if (Bk) {
        mbar_wait(bars + BAR_ACCDONE, ph_acc);       // G3(i-1)
        ph_acc ^= 1;
        TRACE(0, 3);
        umma::fence_after_sync();
        load_acc(tm_lane, cg, v);
        release_acc(bars);
        // dW2'(i-1) (issued right behind G3) has read D2 (DA) and h1 (H1): it ran under S0(i)
        mbar_wait(bars + BAR_DW2, ph_dw2);
        ph_dw2 ^= 1;
        TRACE(0, 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t off = umma::tile_off(TILE, row, 8 * (4 * j + cg));
          *reinterpret_cast<uint4*>(smem + SM_DA + off) =
              dact8(v + 8 * j, *reinterpret_cast<const uint4*>(smem + SM_H1 + off));
        }
        publish_smem(bars + BAR_OP + OP_S6);
        TRACE(0, 5);
      }
      // ---- S2(i): h1 = tanh(ACC + b1) -> H1 --------------------------------------------------------------
      if (F) {
        mbar_wait(bars + BAR_ACCDONE, ph_acc);       // G1(i)
        ph_acc ^= 1;
        TRACE(0, 6);
        umma::fence_after_sync();
        load_acc(tm_lane, cg, v);
        release_acc(bars);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          store_chunk(smem + SM_H1, row, cg, j, tanh8(v + 8 * j, f32 + F_B1 + 8 * (4 * j + cg)));
        publish_smem(bars + BAR_OP + OP_S2);
        TRACE(0, 7);
      }
      // ---- S8(i-1): D0 = dh0 (1 - h0^2) -> H0b (in place), G tile of tile i-1 ---------------------------------
      if (Bk) {
        mbar_wait(bars + BAR_ACCDONE, ph_acc);       // G4(i-1)
        ph_acc ^= 1;
        TRACE(0, 8);
        umma::fence_after_sync();
        load_acc(tm_lane, cg, v);
        release_acc(bars);
        // dW1'(i-1) (issued right behind G4) has read D1 (DA) and h0 (H0b): it ran under S2(i)
        mbar_wait(bars + BAR_DW1, ph_dw1);
        ph_dw1 ^= 1;
        TRACE(0, 9);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t off = umma::tile_off(TILE, row, 8 * (4 * j + cg));
          uint4* hp = reinterpret_cast<uint4*>(H0b + off);
          *hp = dact8(v + 8 * j, *hp);                    // D0 over h0, same thread, same address
        }
        if (cg >= 2) {
          // G[row][3*slot + {0,1,2}] = {gx, gy, 1}; column groups 2 and 3 fill 8 columns each
          const int hf = cg - 2;
          const bool pvalid = (p_gi >> 8) & 1;
          const int pslot = p_gi & 0xff;
          __half g8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            int n = hf * 8 + e;
            float gv = 0.f;
            if (pvalid && n / 3 == pslot) gv = (n % 3 == 0) ? p_gx : (n % 3 == 1) ? p_gy : 1.f;
            g8[e] = __float2half_rn(gv);
          }
          *reinterpret_cast<uint4*>(smem + SM_G + umma::tile_off(TILE, row, hf * 8)) =
              *reinterpret_cast<uint4*>(g8);
        }
        publish_smem(bars + BAR_OP + OP_S8);
        TRACE(0, 10);
      }
      // ---- S4(i): h2, logit, dl, dwo, D2 -> DA; log-lik / reconstruction out -----------------------------------
      if (F) {
        // this row's target / weight / geometry (the staging buffer of tile i is rewritten only
        // after the barrier below)
        const float* stg = f32 + F_STG + (i & 1) * STG_FLOATS;
        const int gi = reinterpret_cast<const int*>(stg)[G_GI + row];
        const float xv = stg[G_X + row], wi = stg[G_WI + row];
        const float gx = stg[G_GX + row], gy = stg[G_GY + row];
        const bool valid = (gi >> 8) & 1;
        mbar_wait(bars + BAR_ACCDONE, ph_acc);       // G2(i)
        ph_acc ^= 1;
        TRACE(0, 11);
        umma::fence_after_sync();
        float pdot = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c0 = 8 * (4 * j + cg);
          if ((j & 1) == 0) {
            // two chunks at a time: the fp32 copy of the accumulator is short-lived
            umma::tmem_ld8(tm_lane + TM_ACC + 8 * (4 * j + cg), v + 8 * j);
            umma::tmem_ld8(tm_lane + TM_ACC + 8 * (4 * (j + 1) + cg), v + 8 * (j + 1));
            umma::tmem_ld_wait();
          }
          const uint4 h2 = tanh8(v + 8 * j, f32 + F_B2 + c0);
          float wv[8];
          lds8(f32 + F_WO + c0, wv);
          const __half2* hh = reinterpret_cast<const __half2*>(&h2);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float2 hf2 = __half22float2(hh[e]);
            pdot = fmaf(hf2.x, wv[2 * e], pdot);
            pdot = fmaf(hf2.y, wv[2 * e + 1], pdot);
          }
          // dUv(i-1) (issued behind G2(i)) has read D0 from H0b; no phase toggle here: the top of
          // the next iteration waits on the same phase again
          if (Bk && j == 0) mbar_wait(bars + BAR_DUV, ph_duv);
          store_chunk(H0b, row, cg, j, h2);      // h2: operand of dwo, re-read below for D2
        }
        float* part = f32 + F_PART + (i & 1) * 4 * TILE;   // by parity: a fast warp's next tile
        part[cg * TILE + row] = pdot;                       // never overwrites a slow warp's reads
        cp_async_wait_all();   // this thread's share of the next tile's staging has landed
        TRACE(0, 12);
        epi_bar();             // partial dots exchanged; staging of tile i+1 published
        TRACE(0, 13);
        const float logit = ((part[row] + part[TILE + row]) +
                             (part[2 * TILE + row] + part[3 * TILE + row])) + bo;
        const float dnll = pvb::obs_dnll_fast(logit, xv, P.sampler, P.sigmoid_d, P.sig);
        const float dl = valid ? wi * dnll : 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float wv[8];
          lds8(f32 + F_WO + 8 * (4 * j + cg), wv);
          const __half2 one = __float2half2_rn(1.f);
          const uint4 h2 = *reinterpret_cast<const uint4*>(H0b + umma::tile_off(TILE, row, 8 * (4 * j + cg)));
          const __half2* hh = reinterpret_cast<const __half2*>(&h2);
          __half2 dd[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            __half2 g = __hfma2(__hneg2(hh[e]), hh[e], one);
            dd[e] = __hmul2(__floats2half2_rn(dl * wv[2 * e], dl * wv[2 * e + 1]), g);
          }
          // DA: D1(i-1) was read by G4(i-1) and dW1'(i-1), both seen complete in S8(i-1)
          store_chunk(smem + SM_DA, row, cg, j, *reinterpret_cast<uint4*>(dd));
        }
        if (cg == 1) {
          // ODL[row][0..7] = {1, dl, 0, ...}: column 0 sums rows (bias gradients), column 1 carries dl
          // (its previous readers db2 / dwo of tile i-1 and db1 of tile i-1 were seen complete in
          // S6(i-1) / S8(i-1))
          __half d8[8];
          d8[0] = __float2half_rn(1.f);
          d8[1] = __float2half_rn(dl);
#pragma unroll
          for (int e = 2; e < 8; ++e) d8[e] = __float2half_rn(0.f);
          *reinterpret_cast<uint4*>(smem + SM_ODL + umma::tile_off(TILE, row, 0)) =
              *reinterpret_cast<uint4*>(d8);
        }
        if (cg == 0) dl_sum += dl;
        umma::fence_before_sync();   // accumulator reads ordered before G3 (issued behind OP_S4)
        publish_smem(bars + BAR_OP + OP_S4);
        TRACE(0, 14);
        if (cg == 0 && valid) {
          // per-pixel log-likelihood and reconstruction (fast intrinsics, ~1e-6 relative)
          float ll, dn_unused, locv;
          pvb::obs_terms_fast(logit, xv, P.sampler, P.sigmoid_d, P.sig, ll, dn_unused, locv);
          const int64_t r_glob = cur_c.tile * TILE + row;
          if (P.rowll) P.rowll[r_glob] = ll;
          if (P.loc) P.loc[r_glob] = locv;
        }
        p_gx = gx;
        p_gy = gy;
        p_gi = gi;
        cursor_advance(cur_c, P);
      }
    }
    // last tile: wait for its dUv, write its partials
    mbar_wait(bars + BAR_DUV, ph_duv);
    umma::fence_after_sync();
    if (cg == 1)
      write_duv(P, tm_lane, row, (int64_t)blockIdx.x + (int64_t)(n_local - 1) * gridDim.x,
                (int)(slots_hist & 0xf));
  }

  // ---- weight-gradient partials of this CTA ---------------------------------------------------------------
  // (every weight-gradient MMA was seen complete by the epilogue warps: DW2 in S6, DW1 in S8 of the
  // last tile)
  {
    float* outp = P.wgrad_part + (size_t)blockIdx.x * PVB_TC_WGRAD_STRIDE;
    // layout: dW1[128][128] | db1[128] | dW2[128][128] | db2[128] | dwo[128] | dbo
    float* o_dW1 = outp;
    float* o_db1 = outp + HD * HD;
    float* o_dW2 = o_db1 + HD;
    float* o_db2 = o_dW2 + HD * HD;
    float* o_dwo = o_db2 + HD;
    float* o_dbo = o_dwo + HD;
    if (warp != MMA_WARP) {
      umma::fence_after_sync();
      const int col0 = cg * 32;             // this thread's 32 contiguous columns of row `row`
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        const uint32_t base = which == 0 ? TM_DW1 : TM_DW2;
        float* oW = which == 0 ? o_dW1 : o_dW2;
        float w32[32];
        umma::tmem_ld32(tm_lane + base + col0, w32);
        umma::tmem_ld_wait();
        float4* dst = reinterpret_cast<float4*>(oW + row * HD + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_float4(w32[4 * j], w32[4 * j + 1], w32[4 * j + 2], w32[4 * j + 3]);
      }
      if (cg >= 1) {
        // lane == hidden unit: db1 / db2 in column 0 of their accumulators, dwo in column 1
        float b[16];
        umma::tmem_ld16(tm_lane + (cg == 1 ? TM_DWO : cg == 2 ? TM_DB1 : TM_DB2), b);
        umma::tmem_ld_wait();
        if (cg == 1) o_dwo[row] = b[1];
        else (cg == 2 ? o_db1 : o_db2)[row] = b[0];
      }
    }
    float tot = pvb::block_sum(dl_sum, f32 + F_RED);
    if (tid == 0) o_dbo[0] = tot;
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == MMA_WARP) umma::tmem_dealloc<TM_COLS>(tm);
}

}  // namespace

namespace pvb_sdec {

int launch_v2(const Params& P, int ctas, cudaStream_t stream) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(sdec_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         SMEM_BYTES);
    if (e != cudaSuccess) {
      pvb::set_error("cudaFuncSetAttribute(sdec_tc2_kernel): %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr = true;
  }
  sdec_tc2_kernel<<<ctas, NTHREADS, SMEM_BYTES, stream>>>(P);
  return 0;
}

}  // namespace pvb_sdec

#ifdef PVB_TC2_TRACE
extern "C" int pvb_tc2_trace_read(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_trace2, sizeof(long long) * 2 * 64 * 16);
}
#endif
