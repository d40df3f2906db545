// Fused spatial-decoder step on tcgen05 tensor cores (sm_100a).
//
// One persistent CTA per SM walks 128-row tiles of the R = I*N pixel rows.
// For each tile everything stays on the SM:
//
//   h0  = tanh(U g + v)                 CUDA cores  -> smem A0 (fp16)
//   h1  = tanh(h0 W1^T + b1)            tcgen05 (TMEM acc) + epilogue -> smem A1
//   h2  = tanh(h1 W2^T + b2)            tcgen05 + epilogue (registers, smem Db)
//   l   = h2 . wo + bo ; log-lik ; dl   per-row epilogue
//   D2  = dl wo (1-h2^2)                -> smem Da
//   dh1 = D2 W2      ; dW2' += D2^T [h1|1] ; dwo += h2^T dl      tcgen05
//   D1  = dh1 (1-h1^2)                  -> smem Db
//   dh0 = D1 W1      ; dW1' += D1^T [h0|1]                       tcgen05
//   D0  = dh0 (1-h0^2)                  -> smem Da
//   dUv = D0^T [gx,gy,1 per sample slot]                          tcgen05
//
// Warp roles: 16 epilogue warps (thread = one tile row x 32 columns, taken as
// four 8-column chunks interleaved over the tile width) + 1 MMA-issuing warp.
// The epilogue of GEMM k writes the next operand 32 columns at a time and
// signals an mbarrier per 32-column group, so the MMA warp issues the K-steps
// of GEMM k+1 while the rest of the epilogue is still running; the
// weight-gradient MMAs are issued behind the dh MMAs and finish under the
// next epilogue.  (One accumulator suffices: every epilogue thread pulls its
// 32 accumulator columns into registers before it signals the first group.)
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
#include "pvb_sdec_tc.cuh"
#include "umma.cuh"

#include <cstdlib>

namespace {
using pvb_sdec::Params;

constexpr int HD = 128;            // hidden width (fixed for this kernel)
constexpr int TILE = 128;          // rows per tile
constexpr int NEPI = 512;          // 16 epilogue warps: (lane quarter q = warp%4) x (column group cg = warp/4)
constexpr int NTHREADS = NEPI + 32;  // + one MMA-issuing warp
constexpr int MMA_WARP = NEPI / 32;
constexpr int MAX_SLOTS = 5;       // samples a 128-row tile can touch when N >= MIN_PIX
constexpr int MIN_PIX = 32;        // floor(127/N) + 2 <= 5
static_assert(TILE == PVB_TC_TILE && MAX_SLOTS == PVB_TC_MAX_SLOTS, "pvb.h constants out of sync");
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
constexpr int UV_FLOATS = MAX_SLOTS * 3 * HD;
// per-tile staging (single-buffered: staged for tile t+1 once GEMM1 of tile t has completed,
// i.e. after every warp has consumed tile t's copy; published by the S4 barrier of tile t)
constexpr int F_UV = 384;                         // [MAX_SLOTS][3][128] first-layer coefficients
constexpr int F_X = F_UV + UV_FLOATS;             // [128] targets of the tile rows
constexpr int F_WI = F_X + TILE;                  // [128] instance weights of the tile rows
constexpr int F_GX = F_WI + TILE;                 // [128] grid x of the tile rows
constexpr int F_GY = F_GX + TILE;                 // [128] grid y
constexpr int F_GI = F_GY + TILE;                 // [128] int: slot | valid << 8 | n_slots << 16
constexpr int F_PART = F_GI + TILE;               // [4][128] partial dots per column group
constexpr int F_RED = F_PART + 4 * TILE;          // [32] block reduction
constexpr int F_END = F_RED + 32;
constexpr int SM_BAR = SM_F32 + F_END * 4;        // mbarriers + tmem base
constexpr int BAR_READY = 0;                      // ready[4]: TMEM A column group written (16 warps)
constexpr int BAR_SM = 4;                         // sm[5]: smem operands of S0,S2,S4,S6,S8 published
constexpr int BAR_ACC = 9;                        // accumulator of the current GEMM complete
constexpr int BAR_DUV = 10;                       // all MMAs of the tile complete (dUv last)
constexpr int BAR_DW = 11;                        // dW1' MMAs complete (A0, Db reusable)
constexpr int BAR_DWO = 12;                       // dwo MMAs complete (Db: h2 -> D1)
constexpr int BAR_W = 13;                         // pre-packed weights landed (bulk copy, 64 KB)
constexpr int BAR_DW2 = 14;                       // dW2' MMAs complete (Da: D2 -> D0)
constexpr int N_BARS = 15;
constexpr int SMEM_BYTES = SM_BAR + (N_BARS + 1) * 8;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget exceeded");

// ---- tensor memory map (columns) -------------------------------------------------
constexpr uint32_t TM_ACC = 0;     // 128 cols : forward accumulators / dh
constexpr uint32_t TM_DW1 = 128;   // 144 cols : dW1 (128) | db1 (col 128)
constexpr uint32_t TM_DW2 = 272;   // 144 cols : dW2 | db2
constexpr uint32_t TM_DWO = 416;   // 16 cols  : dwo in column 0
constexpr uint32_t TM_DUV = 432;   // 16 cols  : per-tile dUv
constexpr uint32_t TM_AT = 448;    // 64 cols  : A operand of the forward / dh GEMMs (128 x 128 fp16)
constexpr int TM_COLS = 512;

#ifdef PVB_TC_TRACE
// debug build only: per-stage timestamps of CTA 0 (epilogue warp 0 / MMA warp), tools/tc_trace.py
__device__ long long g_trace[2][64][32];
#define TRACE(role, ev)                                                              \
  do {                                                                               \
    if (blockIdx.x == 0 && lane == 0 && (role == 1 || warp == 0) && trace_it < 64)   \
      g_trace[role][trace_it][ev] = clock64();                                       \
  } while (0)
#define TRACE_NEXT() ++trace_it
__device__ long long g_wtrace[16][16];
#define WTRACE(ev)                                                                    \
  do {                                                                               \
    if (blockIdx.x == 0 && lane == 0 && trace_it == 3) g_wtrace[warp][ev] = clock64(); \
  } while (0)
// whole-kernel milestones of CTA 0 (entry, set-up done, first tile, loop done, partials written, exit)
__device__ long long g_ktrace[8];
#define KTRACE(ev, cond) do { if (blockIdx.x == 0 && (cond)) g_ktrace[ev] = clock64(); } while (0)
#else
#define WTRACE(ev) do {} while (0)
#define TRACE(role, ev) do {} while (0)
#define TRACE_NEXT() do {} while (0)
#define KTRACE(ev, cond) do {} while (0)
#endif

__device__ __forceinline__ float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}


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

// barrier among the 16 epilogue warps only (the MMA warp never joins)
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

// two tanh -> one packed fp16 pair.  (tanh.approx.f16x2 was tried: ptxas lowers it to TWO
// MUFU.TANH.F16 operations, one per half, so it saves nothing on the MUFU pipe that bounds the
// tanh stages, and it rounds the pre-activation to fp16 first.)
__device__ __forceinline__ __half2 tanh_h2(float a, float b) {
  return __floats2half2_rn(fast_tanh(a), fast_tanh(b));
}

// 8 columns of tanh(v + bias) -> four fp16 pairs
__device__ __forceinline__ uint4 tanh8(const float* v, const float* bias) {
  float b[8];
  lds8(bias, b);
  __half2 hh[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) hh[e] = tanh_h2(v[2 * e] + b[2 * e], v[2 * e + 1] + b[2 * e + 1]);
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

// ---- tile geometry without per-thread 64-bit divisions -----------------------------------------
// A CTA walks tiles blockIdx.x, +grid, +2 grid, ...; (i_first, off) = first instance of the tile
// and the tile's first row inside it, advanced incrementally with host-computed step constants.
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

// (offset inside the first instance) -> slot, pixel; N >= MIN_PIX bounds the quotient by 3
__device__ __forceinline__ void split_slot(int rem, int N, int& slot, int& pix) {
  slot = 0;
#pragma unroll
  for (int s = 0; s < MAX_SLOTS - 1; ++s)
    if (rem >= N) { rem -= N; ++slot; }
  pix = rem;
}

// asynchronous staging of the cursor's tile: Uv rows (all threads), targets (column group 0),
// instance weights (group 1), row geometry (group 3)
__device__ __forceinline__ void stage_tile(const Params& P, float* f32, const TileCursor& c,
                                           int tid, int row, int cg) {
  const int64_t left = P.R - c.tile * TILE;                 // rows from the tile start to R
  const int last_row = left < TILE ? (int)left - 1 : TILE - 1;
  int last_slot, last_pix;
  split_slot(c.off + last_row, P.N, last_slot, last_pix);
  const int n_slots = last_slot + 1;
  if (tid < n_slots * (3 * HD / 4))
    cp_async16(f32 + F_UV + tid * 4, P.Uv + c.i_first * 3 * HD + tid * 4);
  const bool valid = row <= last_row;
  int slot, pix;
  split_slot(c.off + (valid ? row : 0), P.N, slot, pix);
  if (cg == 0) {
    if (valid && P.x) {
      int ib = c.ib_first + slot;
      while (ib >= (int)P.B) ib -= (int)P.B;
      cp_async4(f32 + F_X + row, P.x + (int64_t)ib * P.N + pix);
    }
  } else if (cg == 1) {
    if (valid && P.w) cp_async4(f32 + F_WI + row, P.w + c.i_first + slot);
  } else if (cg == 3) {
    float gx = 0.f, gy = 0.f;
    pvb::grid_xy(pix, P.H, P.W, P.ndim, gx, gy);
    f32[F_GX + row] = gx;
    f32[F_GY + row] = gy;
    reinterpret_cast<int*>(f32)[F_GI + row] = slot | ((int)valid << 8) | (n_slots << 16);
  }
}

// chunk j of this thread -> tensor-memory A operand, then signal column group j.
// (tcgen05.st is ordered against the MMA by wait::st + fence::before_thread_sync; no
// generic->async proxy fence is involved, which is what makes a per-chunk signal cheap.)
__device__ __forceinline__ void publish_chunk(uint64_t* bars, uint32_t tm_lane, int cg, int j, uint4 out) {
  umma::tmem_st4(tm_lane + TM_AT + 4 * (4 * j + cg), out);
  umma::tmem_st_wait();
  umma::fence_before_sync();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) umma::mbar_arrive(bars + BAR_READY + j);
}
// chunks j0, j0 + 1 with ONE store-wait / fence / warp handshake (both column groups are signalled):
// for the stages whose arithmetic per chunk is a handful of packed multiplies, the handshake
// (~120 cycles) costs more than the earlier start of two K-steps gains
__device__ __forceinline__ void publish_pair(uint64_t* bars, uint32_t tm_lane, int cg, int j0, uint4 a,
                                             uint4 b) {
  umma::tmem_st4(tm_lane + TM_AT + 4 * (4 * j0 + cg), a);
  umma::tmem_st4(tm_lane + TM_AT + 4 * (4 * (j0 + 1) + cg), b);
  umma::tmem_st_wait();
  umma::fence_before_sync();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    umma::mbar_arrive(bars + BAR_READY + j0);
    umma::mbar_arrive(bars + BAR_READY + j0 + 1);
  }
}
// this thread's 4 chunks -> row-chunk tile in shared memory (operands of the weight-gradient
// MMAs and of later element-wise passes), then one proxy fence + signal per warp
__device__ __forceinline__ void store_tile4(uint8_t* tile, int row, int cg, const uint4* out) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<uint4*>(tile + umma::tile_off(TILE, row, 8 * (4 * j + cg))) = out[j];
}
__device__ __forceinline__ void store_chunk(uint8_t* tile, int row, int cg, int j, uint4 v) {
  *reinterpret_cast<uint4*>(tile + umma::tile_off(TILE, row, 8 * (4 * j + cg))) = v;
}
__device__ __forceinline__ void signal_smem(uint64_t* bars, int stage) {
  umma::fence_proxy_async();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) umma::mbar_arrive(bars + BAR_SM + stage);
}

__global__ void __launch_bounds__(NTHREADS, 1) sdec_tc_kernel(Params P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* f32 = reinterpret_cast<float*>(smem + SM_F32);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + N_BARS * 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, cg = (warp >> 2) & 3;
  const int row = q * 32 + lane;        // tile row == TMEM lane owned by this thread
  KTRACE(0, tid == 0);

  // ---- one-time setup ----------------------------------------------------------
  // weights: pre-packed fp16 operand tiles come in through the TMA engine (one thread, two bulk
  // copies, issued below once the barrier is initialised); otherwise every thread converts its
  // share of the fp32 weights
  if (!P.Wp) {
    stage_weight(P.W1, smem + SM_W1, tid);
    stage_weight(P.W2, smem + SM_W2, tid);
  }
  if (tid < HD) {
    f32[F_B1 + tid] = P.b1[tid];
    f32[F_B2 + tid] = P.b2[tid];
    f32[F_WO + tid] = P.wo[tid];
  }
  if (tid < NEPI) {
    // constant chunk-columns: [A0|A1] column 128 = 1 (bias column), 129..143 = 0;
    // DL columns 8..15 = 0; default targets 0 / weights 1
    uint4 ones = make_uint4(0x00003C00u, 0u, 0u, 0u);  // fp16 {1,0,0,0,0,0,0,0}
    uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    if (cg == 0) {
      *reinterpret_cast<uint4*>(smem + SM_A0 + umma::tile_off(TILE, row, 128)) = ones;
      *reinterpret_cast<uint4*>(smem + SM_A0 + umma::tile_off(TILE, row, 136)) = zero;
    } else if (cg == 1) {
      *reinterpret_cast<uint4*>(smem + SM_A1 + umma::tile_off(TILE, row, 128)) = ones;
      *reinterpret_cast<uint4*>(smem + SM_A1 + umma::tile_off(TILE, row, 136)) = zero;
    } else if (cg == 2) {
      *reinterpret_cast<uint4*>(smem + SM_DL + umma::tile_off(TILE, row, 8)) = zero;
    } else {
      f32[F_X + row] = 0.f;
      f32[F_WI + row] = 1.f;
    }
  }
  if (warp == MMA_WARP) umma::tmem_alloc<TM_COLS>(tmem_slot);
  if (tid == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) umma::mbar_init(bars + BAR_READY + j, NEPI / 32);
#pragma unroll
    for (int j = 0; j < 5; ++j) umma::mbar_init(bars + BAR_SM + j, NEPI / 32);
    umma::mbar_init(bars + BAR_ACC, 1);
    umma::mbar_init(bars + BAR_DUV, 1);
    umma::mbar_init(bars + BAR_DW, 1);
    umma::mbar_init(bars + BAR_DWO, 1);
    umma::mbar_init(bars + BAR_W, 1);
    umma::mbar_init(bars + BAR_DW2, 1);
    umma::mbar_fence_init();
    if (P.Wp) pvb_sdec::bulk_load_weights(P.Wp, smem + SM_W1, smem + SM_W2, bars + BAR_W);
  }
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = *tmem_slot;
  const uint32_t tm_lane = tm + ((uint32_t)(q * 32) << 16);
  KTRACE(1, tid == 0);
  float dl_sum = 0.f;  // sum of dl over this thread's rows (column group 0 only) -> dbo
  bool any_tile = false;
#ifdef PVB_TC_TRACE
  int trace_it = 0;
#endif

  if (warp == MMA_WARP) {
    // =========================== MMA issuer =====================================
    const uint32_t sW1 = umma::smem_u32(smem + SM_W1), sW2 = umma::smem_u32(smem + SM_W2);
    const uint32_t sA0 = umma::smem_u32(smem + SM_A0), sA1 = umma::smem_u32(smem + SM_A1);
    const uint32_t sDA = umma::smem_u32(smem + SM_DA), sDB = umma::smem_u32(smem + SM_DB);
    const uint32_t sG = umma::smem_u32(smem + SM_G), sDL = umma::smem_u32(smem + SM_DL);
    const uint32_t tA = tm + TM_AT;
    constexpr uint32_t ID_FWD = umma::idesc_f16(128, 128, 0, 0);   // A (TMEM), B K-major
    constexpr uint32_t ID_DH = umma::idesc_f16(128, 128, 0, 1);    // A (TMEM), B MN-major
    constexpr uint32_t ID_DW = umma::idesc_f16(128, 144, 1, 1);    // both MN-major, N = 128+16
    constexpr uint32_t ID_N16 = umma::idesc_f16(128, 16, 1, 1);    // both MN-major, N = 16
    uint32_t rph = 0, sph = 0;
    uint32_t acc = 0;  // 0 on this CTA's first tile: weight-gradient accumulators start fresh
    if (P.Wp) umma::mbar_wait(bars + BAR_W, 0);   // both weight tiles have landed
    for (int64_t tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
      // ---- GEMM1: ACC = h0 W1^T, K-steps issued as the h0 column groups land in TMEM ----
      TRACE(1, 0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        umma::mbar_wait(bars + BAR_READY + j, rph);
        TRACE(1, 1 + j);
        umma::fence_after_sync();
        if (umma::elect_one()) {
          umma::mma_f16_ts(tm + TM_ACC, tA + 16 * j, desc_kmajor(sW1, 2 * j), ID_FWD, j > 0);
          umma::mma_f16_ts(tm + TM_ACC, tA + 16 * j + 8, desc_kmajor(sW1, 2 * j + 1), ID_FWD, 1);
          if (j == 3) umma::commit(bars + BAR_ACC);
        }
        __syncwarp();
      }
      rph ^= 1;
      // ---- GEMM2: ACC = h1 W2^T ----
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        umma::mbar_wait(bars + BAR_READY + j, rph);
        TRACE(1, 5 + j);
        umma::fence_after_sync();
        if (umma::elect_one()) {
          umma::mma_f16_ts(tm + TM_ACC, tA + 16 * j, desc_kmajor(sW2, 2 * j), ID_FWD, j > 0);
          umma::mma_f16_ts(tm + TM_ACC, tA + 16 * j + 8, desc_kmajor(sW2, 2 * j + 1), ID_FWD, 1);
          if (j == 3) umma::commit(bars + BAR_ACC);
        }
        __syncwarp();
      }
      rph ^= 1;
      if (!P.backward) {
        TRACE_NEXT();
        continue;
      }
      // ---- GEMM3: ACC = D2 W2 ----
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        umma::mbar_wait(bars + BAR_READY + j, rph);
        TRACE(1, 9 + j);
        umma::fence_after_sync();
        if (umma::elect_one()) {
          umma::mma_f16_ts(tm + TM_ACC, tA + 16 * j, desc_mnmajor(sW2, 2 * j), ID_DH, j > 0);
          umma::mma_f16_ts(tm + TM_ACC, tA + 16 * j + 8, desc_mnmajor(sW2, 2 * j + 1), ID_DH, 1);
          if (j == 3) umma::commit(bars + BAR_ACC);
        }
        __syncwarp();
      }
      rph ^= 1;
      // shared-memory operands of S0 (h0), S2 (h1), S4 (h2, D2, dl): one publication, in S4
      umma::mbar_wait(bars + BAR_SM + 2, sph);
      umma::fence_after_sync();
      if (umma::elect_one()) {
        // dwo += h2^T dl ; Db (h2) may be overwritten with D1 once this completes
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma::mma_f16_ss(tm + TM_DWO, desc_mnmajor(sDB, k), desc_mnmajor(sDL, k), ID_N16,
                           (k > 0) ? 1u : acc);
        umma::commit(bars + BAR_DWO);
      }
      __syncwarp();
      // ---- GEMM4: ACC = D1 W1 ----
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        umma::mbar_wait(bars + BAR_READY + j, rph);
        TRACE(1, 13 + j);
        umma::fence_after_sync();
        if (umma::elect_one()) {
          umma::mma_f16_ts(tm + TM_ACC, tA + 16 * j, desc_mnmajor(sW1, 2 * j), ID_DH, j > 0);
          umma::mma_f16_ts(tm + TM_ACC, tA + 16 * j + 8, desc_mnmajor(sW1, 2 * j + 1), ID_DH, 1);
          if (j == 3) umma::commit(bars + BAR_ACC);
        }
        __syncwarp();
      }
      rph ^= 1;
      // ---- dW2' += D2^T [h1|1]: BEHIND GEMM4 in the (in-order) tensor pipe -- in front of it, its
      // 0.6 k cycles delayed GEMM4 and with it S8 (trace: 0.8 k of waiting per tile); S8 now waits
      // for it only before it overwrites Da (D2 -> D0), after its arithmetic ----
      if (umma::elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma::mma_f16_ss(tm + TM_DW2, desc_mnmajor(sDA, k), desc_mnmajor(sA1, k), ID_DW,
                           (k > 0) ? 1u : acc);
        umma::commit(bars + BAR_DW2);
      }
      __syncwarp();
      // ---- dW1' += D1^T [h0|1] (needs the shared-memory copy of D1) ----
      umma::mbar_wait(bars + BAR_SM + 3, sph);
      umma::fence_after_sync();
      if (umma::elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma::mma_f16_ss(tm + TM_DW1, desc_mnmajor(sDB, k), desc_mnmajor(sA0, k), ID_DW,
                           (k > 0) ? 1u : acc);
        umma::commit(bars + BAR_DW);
      }
      __syncwarp();
      // ---- dUv(tile) = D0^T G ----
      umma::mbar_wait(bars + BAR_SM + 4, sph);
      sph ^= 1;
      TRACE(1, 17);
      umma::fence_after_sync();
      if (umma::elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma::mma_f16_ss(tm + TM_DUV, desc_mnmajor(sDA, k), desc_mnmajor(sG, k), ID_N16, k > 0);
        umma::commit(bars + BAR_DUV);
      }
      __syncwarp();
      acc = 1;
      TRACE_NEXT();
    }
  } else {
    // =========================== epilogue warps ====================================
    const float bo = P.bo[0];
    uint32_t aph = 0, dph = 0, wph = 0, oph = 0, w2ph = 0;
    int64_t prev_tile = -1;
    int prev_slots = 0;
    TileCursor cur_c;
    cur_c.tile = blockIdx.x;                       // grid <= tiles: every CTA owns a tile
    cur_c.i_first = (cur_c.tile * TILE) / P.N;     // the only 64-bit divisions of the kernel
    cur_c.off = (int)(cur_c.tile * TILE - cur_c.i_first * P.N);
    cur_c.ib_first = (int)(cur_c.i_first % P.B);
    stage_tile(P, f32, cur_c, tid, row, cg);
    cp_async_wait_all();
    epi_bar();
    KTRACE(2, tid == 0);
    while (cur_c.tile < P.tiles) {
      const int64_t tile = cur_c.tile;
      any_tile = true;
      TRACE(0, 0);
      // this row's staged geometry / target / weight (published by the previous S4 barrier)
      const int gi = reinterpret_cast<const int*>(f32)[F_GI + row];
      const bool valid = (gi >> 8) & 1;
      const int slot = gi & 0xff, n_slots = gi >> 16;
      const float gx = f32[F_GX + row], gy = f32[F_GY + row];
      const int64_t r_glob = tile * TILE + row;
      const float xv = f32[F_X + row];
      const float wi = f32[F_WI + row];
      uint4 out[4];
      // ---- S0: first layer h0 = tanh(U g + v) -> TMEM A (+ A0) ------------------------------
      {
        const float* u = f32 + F_UV + slot * 3 * HD;
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
            const float a = fmaf(ux[2 * e], gx, fmaf(uy[2 * e], gy, uc[2 * e]));
            const float b = fmaf(ux[2 * e + 1], gx, fmaf(uy[2 * e + 1], gy, uc[2 * e + 1]));
            hh[e] = valid ? tanh_h2(a, b) : __floats2half2_rn(0.f, 0.f);
          }
          out[j] = *reinterpret_cast<uint4*>(hh);
          publish_chunk(bars, tm_lane, cg, j, out[j]);
          if (j == 0 && P.backward && prev_tile >= 0) {
            // dW1' of the previous tile (last but one in the tensor queue) has finished reading
            // A0 / Db: waited for here, behind the first chunk's arithmetic, not at the tile start
            umma::mbar_wait(bars + BAR_DW, wph);
            wph ^= 1;
          }
          // smem copy right behind the signal: by the end of the stage only the last store
          // is still in flight when the proxy fence drains them
          if (P.backward) store_chunk(smem + SM_A0, row, cg, j, out[j]);
          TRACE(0, 1 + j);
        }
      }
      if (P.backward) {
        if (prev_tile >= 0) {
          // every MMA of the previous tile (dUv was issued last) is complete: Da, G and the
          // dUv accumulator are free
          umma::mbar_wait(bars + BAR_DUV, dph);
          dph ^= 1;
          umma::fence_after_sync();
        }
        if (cg >= 2) {
          // G[row][3*slot + {0,1,2}] = {gx, gy, 1}; column groups 2 and 3 fill 8 columns each
          const int hf = cg - 2;
          __half g8[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            int n = hf * 8 + e;
            float v = 0.f;
            if (valid && n / 3 == slot) v = (n % 3 == 0) ? gx : (n % 3 == 1) ? gy : 1.f;
            g8[e] = __float2half_rn(v);
          }
          *reinterpret_cast<uint4*>(smem + SM_G + umma::tile_off(TILE, row, hf * 8)) =
              *reinterpret_cast<uint4*>(g8);
        }
        // (h0 and G are published by the fence of a later stage of this thread: S4 for the
        // weight-gradient operands, S8 for G -- a proxy fence orders ALL earlier writes)
        if (prev_tile >= 0 && cg == 1) {
          // per-tile dUv partials of the previous tile: lane == hidden unit (after the fence:
          // a MEMBAR would otherwise wait for these global stores)
          float v[16];
          umma::tmem_ld16(tm_lane + TM_DUV, v);
          umma::tmem_ld_wait();
          float* dst = P.gUv_part + prev_tile * (MAX_SLOTS * 3 * HD);
#pragma unroll
          for (int n = 0; n < MAX_SLOTS * 3; ++n)
            if (n < prev_slots * 3) dst[n * HD + row] = v[n];   // unused slots are never read
        }
      }
      TRACE(0, 5);
      prev_tile = tile;
      prev_slots = n_slots;
      float v[32];
      TRACE(0, 6);
      // ---- S2: h1 = tanh(ACC + b1) -> TMEM A (+ A1) -------------------------------------------
      umma::mbar_wait(bars + BAR_ACC, aph);
      aph ^= 1;
      TRACE(0, 7);
      umma::fence_after_sync();
      load_acc(tm_lane, cg, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        out[j] = tanh8(v + 8 * j, f32 + F_B1 + 8 * (4 * j + cg));
        publish_chunk(bars, tm_lane, cg, j, out[j]);
        if (P.backward) store_chunk(smem + SM_A1, row, cg, j, out[j]);
      }
      // next tile's staging: any time after GEMM1 of this tile (every warp has then consumed the
      // current copy).  Forward-only: here, published by the S4 barrier.  With backward: in the
      // shadow of GEMM4 + dW2' (tensor-bound phase), published by a barrier at the tile end.
      TileCursor nxt_c = cur_c;
      cursor_advance(nxt_c, P);
      if (!P.backward && nxt_c.tile < P.tiles) stage_tile(P, f32, nxt_c, tid, row, cg);
      TRACE(0, 8);
      // ---- S4: h2, logit, dl, D2 -> TMEM A (+ Db, Da, DL) ------------------------------------------
      umma::mbar_wait(bars + BAR_ACC, aph);
      aph ^= 1;
      TRACE(0, 9);
      umma::fence_after_sync();
      load_acc(tm_lane, cg, v);
      float pdot = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c0 = 8 * (4 * j + cg);
        out[j] = tanh8(v + 8 * j, f32 + F_B2 + c0);
        float wv[8];
        lds8(f32 + F_WO + c0, wv);
        const __half2* hh = reinterpret_cast<const __half2*>(&out[j]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 hf2 = __half22float2(hh[e]);
          pdot = fmaf(hf2.x, wv[2 * e], pdot);
          pdot = fmaf(hf2.y, wv[2 * e + 1], pdot);
        }
        if (P.backward) {
          store_chunk(smem + SM_DB, row, cg, j, out[j]);   // h2: dwo operand
          // everything of D2 = dl wo (1 - h2^2) that does not need dl is formed here, in the shadow
          // of the tanh stage: after the exchange below only one packed multiply per pair is left
          const __half2 one = __float2half2_rn(1.f);
          __half2 g2[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            g2[e] = __hmul2(__floats2half2_rn(wv[2 * e], wv[2 * e + 1]),
                            __hfma2(__hneg2(hh[e]), hh[e], one));
          out[j] = *reinterpret_cast<uint4*>(g2);
        }
      }
      f32[F_PART + cg * TILE + row] = pdot;
      cp_async_wait_all();   // this thread's share of the next tile's staging has landed
      TRACE(0, 10);
      // exchange of the partial dots: with a backward pass only the four warps that share these
      // 32 rows have to meet (the staging of the next tile is published at the tile end);
      // forward-only, this barrier also publishes the staging: all epilogue warps
      if (P.backward) asm volatile("bar.sync %0, 128;\n" ::"r"(2 + q) : "memory");
      else epi_bar();
      TRACE(0, 11);
      const float logit = ((f32[F_PART + row] + f32[F_PART + TILE + row]) +
                           (f32[F_PART + 2 * TILE + row] + f32[F_PART + 3 * TILE + row])) + bo;
      // observation terms (fast intrinsics, ~1e-6 relative): dl feeds the backward pass, the
      // per-pixel log-likelihood and reconstruction go out from column group 0
      if (P.backward) {
        // critical path first: dl by the shortest chain, D2 -> TMEM A
        const float dnll = P.x ? pvb::obs_dnll_fast(logit, xv, P.sampler, P.sigmoid_d, P.sig) : 0.f;
        const float dl = valid ? wi * dnll : 0.f;
        uint4 d2[4];
        const __half2 dl2 = __float2half2_rn(dl);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const __half2* g2 = reinterpret_cast<const __half2*>(&out[j]);
          __half2 dd[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) dd[e] = __hmul2(dl2, g2[e]);
          d2[j] = *reinterpret_cast<uint4*>(dd);
          if (j & 1) {
            publish_pair(bars, tm_lane, cg, j - 1, d2[j - 1], d2[j]);
            store_chunk(smem + SM_DA, row, cg, j - 1, d2[j - 1]);
            store_chunk(smem + SM_DA, row, cg, j, d2[j]);
          }
        }
        TRACE(0, 12);
        // dl -> DL (dwo operand)
        if (cg == 1) {
          __half d8[8];
          d8[0] = __float2half_rn(dl);
#pragma unroll
          for (int e = 1; e < 8; ++e) d8[e] = __float2half_rn(0.f);
          *reinterpret_cast<uint4*>(smem + SM_DL + umma::tile_off(TILE, row, 0)) =
              *reinterpret_cast<uint4*>(d8);
        }
        if (cg == 0) dl_sum += dl;
        signal_smem(bars, 2);
      }
      // (handing this block to the MMA warp was tried: its serial MUFU chains delayed the issue of
      // GEMM4 by ~1.5 k cycles per tile)
      if (cg == 0 && valid) {
        // per-pixel log-likelihood and reconstruction (fast intrinsics, ~1e-6 relative)
        float ll = 0.f, dn_unused, locv;
        if (P.x) {
          pvb::obs_terms_fast(logit, xv, P.sampler, P.sigmoid_d, P.sig, ll, dn_unused, locv);
        } else {
          locv = P.sigmoid_d ? __fdividef(1.f, 1.f + __expf(-logit)) : logit;
        }
        if (P.rowll) P.rowll[r_glob] = ll;
        if (P.loc) P.loc[r_glob] = locv;
      }
      TRACE(0, 13);
      if (!P.backward) {
        umma::fence_before_sync();   // accumulator reads done before the next tile's signals
        cur_c = nxt_c;
        TRACE_NEXT();
        continue;
      }
      // ---- S6: D1 = dh1 (1 - h1^2) -> TMEM A (+ Db) ---------------------------------------------------
      umma::mbar_wait(bars + BAR_ACC, aph);
      aph ^= 1;
      TRACE(0, 14);
      umma::fence_after_sync();
      load_acc(tm_lane, cg, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t off = umma::tile_off(TILE, row, 8 * (4 * j + cg));
        out[j] = dact8(v + 8 * j, *reinterpret_cast<const uint4*>(smem + SM_A1 + off));
        if (j & 1) publish_pair(bars, tm_lane, cg, j - 1, out[j - 1], out[j]);
      }
      if (nxt_c.tile < P.tiles) stage_tile(P, f32, nxt_c, tid, row, cg);
      umma::mbar_wait(bars + BAR_DWO, oph);   // dwo has finished reading h2 from Db
      oph ^= 1;
      store_tile4(smem + SM_DB, row, cg, out);
      signal_smem(bars, 3);
      TRACE(0, 15);
      // ---- S8: D0 = dh0 (1 - h0^2) -> Da ---------------------------------------------------------------
      umma::mbar_wait(bars + BAR_ACC, aph);
      aph ^= 1;
      TRACE(0, 16);
      umma::fence_after_sync();
      load_acc(tm_lane, cg, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t off = umma::tile_off(TILE, row, 8 * (4 * j + cg));
        out[j] = dact8(v + 8 * j, *reinterpret_cast<const uint4*>(smem + SM_A0 + off));
      }
      umma::mbar_wait(bars + BAR_DW2, w2ph);   // dW2' has read D2 from Da
      w2ph ^= 1;
      store_tile4(smem + SM_DA, row, cg, out);
      umma::fence_before_sync();   // accumulator reads done before the next tile's signals
      signal_smem(bars, 4);
      cp_async_wait_all();
      epi_bar();             // publishes the staging of the next tile
      TRACE(0, 17);
      cur_c = nxt_c;
      TRACE_NEXT();
    }
    KTRACE(3, tid == 0);
    // last tile: wait for its MMAs, write its dUv partials
    if (P.backward && prev_tile >= 0) {
      umma::mbar_wait(bars + BAR_DUV, dph);
      umma::fence_after_sync();
      if (cg == 1) {
        float v[16];
        umma::tmem_ld16(tm_lane + TM_DUV, v);
        umma::tmem_ld_wait();
        float* dst = P.gUv_part + prev_tile * (MAX_SLOTS * 3 * HD);
#pragma unroll
        for (int n = 0; n < MAX_SLOTS * 3; ++n)
          if (n < prev_slots * 3) dst[n * HD + row] = v[n];
      }
    }
  }

  KTRACE(6, tid == 0);
  // ---- weight-gradient partials of this CTA ---------------------------------------------------------------
  if (P.backward) {
    float* out = P.wgrad_part + (size_t)blockIdx.x * PVB_TC_WGRAD_STRIDE;
    // layout: dW1[128][128] | db1[128] | dW2[128][128] | db2[128] | dwo[128] | dbo
    float* o_dW1 = out;
    float* o_db1 = out + HD * HD;
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
        float* ob = which == 0 ? o_db1 : o_db2;
        float v[32];
        if (any_tile) {
          umma::tmem_ld32(tm_lane + base + col0, v);
          umma::tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        // A thread holds 32 columns of ONE row: stored directly, each warp instruction touched 32 different
        // lines (8 192 store transactions per CTA, 13 k cycles = 6.6 us at the end of every launch, trace).  The
        // warp's 32 x 32 block goes through a padded shared-memory block instead (the operand tiles are dead:
        // every MMA has completed), and leaves as 32 full-line rows.
        float* tr = reinterpret_cast<float*>(smem) + warp * (32 * 33);
#pragma unroll
        for (int j = 0; j < 32; ++j) tr[lane * 33 + j] = v[j];
        __syncwarp();
        float* og = oW + (q * 32) * HD + col0 + lane;
#pragma unroll
        for (int r = 0; r < 32; ++r) og[r * HD] = tr[r * 33 + lane];
        __syncwarp();
        if (cg == 3) {
          float b[16];
          if (any_tile) {
            umma::tmem_ld16(tm_lane + base + 128, b);
            umma::tmem_ld_wait();
          } else {
            b[0] = 0.f;
          }
          ob[row] = b[0];
        }
      }
      if (cg == 0) {
        float b[16];
        if (any_tile) {
          umma::tmem_ld16(tm_lane + TM_DWO, b);
          umma::tmem_ld_wait();
        } else {
          b[0] = 0.f;
        }
        o_dwo[row] = b[0];
      }
    }
    KTRACE(7, tid == 0);
    float tot = pvb::block_sum(dl_sum, f32 + F_RED);
    if (tid == 0) o_dbo[0] = tot;
  }
  KTRACE(4, tid == 0);
  umma::fence_before_sync();
  __syncthreads();
  if (warp == MMA_WARP) umma::tmem_dealloc<TM_COLS>(tm);
  KTRACE(5, tid == 0);
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

#ifdef PVB_TC_TRACE
extern "C" int pvb_tc_trace_read(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_trace, sizeof(long long) * 2 * 64 * 32);
}
extern "C" int pvb_tc_ktrace_read(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_ktrace, sizeof(long long) * 8);
}
extern "C" int pvb_tc_wtrace_read(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, g_wtrace, sizeof(long long) * 16 * 16);
}
#endif

extern "C" int pvb_sdec_tc_sizes(int64_t I, int N, pvb_tc_sizes* out) {
  PVB_CHECK_ARG(out && I >= 0 && N >= MIN_PIX, "pvb_sdec_tc_sizes: need N >= 32 pixels per instance");
  int64_t R = I * N;
  out->tiles = (R + TILE - 1) / TILE;
  int sms = sm_count();
  out->ctas = (int)(out->tiles < sms ? (out->tiles > 0 ? out->tiles : 1) : sms);
  out->gUv_part_floats = out->tiles * MAX_SLOTS * 3 * HD;
  out->wgrad_part_floats = (int64_t)out->ctas * PVB_TC_WGRAD_STRIDE;
  return 0;
}

namespace {
// fp32 [128][128] x 2 -> fp16 operand tiles (row-chunk layout), W1 at +0, W2 at +32 KB
__global__ void sdec_pack_weights_kernel(const float* __restrict__ W1, const float* __restrict__ W2,
                                         uint8_t* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;      // one 16-byte piece each
  if (idx >= 2 * HD * (HD / 8)) return;
  const int which = idx / (HD * (HD / 8)), rem = idx % (HD * (HD / 8));
  const int r = rem / (HD / 8), c8 = rem % (HD / 8);
  const float4* src = reinterpret_cast<const float4*>((which ? W2 : W1) + r * HD + c8 * 8);
  const float4 a = __ldg(src), b = __ldg(src + 1);
  __half2 h[4] = {__floats2half2_rn(a.x, a.y), __floats2half2_rn(a.z, a.w),
                  __floats2half2_rn(b.x, b.y), __floats2half2_rn(b.z, b.w)};
  *reinterpret_cast<uint4*>(out + which * 16 * CHUNK + umma::tile_off(TILE, r, c8 * 8)) =
      *reinterpret_cast<uint4*>(h);
}
}  // namespace

extern "C" int64_t pvb_sdec_tc_packed_weight_bytes(void) { return 2 * 16 * CHUNK; }

extern "C" int pvb_sdec_tc_pack_weights(const float* W1, const float* W2, void* packed, void* stream) {
  PVB_CHECK_ARG(W1 && W2 && packed, "pvb_sdec_tc_pack_weights: null pointer");
  PVB_CHECK_ARG(((uintptr_t)W1 % 16 == 0) && ((uintptr_t)W2 % 16 == 0) && ((uintptr_t)packed % 16 == 0),
                "pvb_sdec_tc_pack_weights: buffers must be 16-byte aligned");
  sdec_pack_weights_kernel<<<(2 * HD * (HD / 8) + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      W1, W2, reinterpret_cast<uint8_t*>(packed));
  pvb::count_launch();
  return pvb::launch_status();
}

extern "C" int pvb_sdec_tc_step(const float* Uv, const float* x, const float* w, const float* W1,
                                const float* b1, const float* W2, const float* b2, const float* wo,
                                const float* bo, float* rowll, float* loc, float* gUv_part,
                                float* wgrad_part, int64_t I, int64_t B, int H, int W, int ndim,
                                int sampler, int sigmoid_d, float decoder_sig, int backward,
                                const void* packed_w, void* stream) {
  PVB_CHECK_ARG(Uv && W1 && b1 && W2 && b2 && wo && bo, "pvb_sdec_tc_step: null weights");
  PVB_CHECK_ARG(ndim == 1 || ndim == 2, "pvb_sdec_tc_step: ndim must be 1 or 2");
  PVB_CHECK_ARG(I >= 0 && B > 0 && H > 0 && W > 0, "pvb_sdec_tc_step: bad dims");
  PVB_CHECK_ARG(sampler >= PVB_SAMPLER_BERNOULLI && sampler <= PVB_SAMPLER_CONT_BERNOULLI,
                "pvb_sdec_tc_step: unknown sampler %d", sampler);
  PVB_CHECK_ARG(!backward || (x && gUv_part && wgrad_part), "pvb_sdec_tc_step: backward needs x and workspaces");
  PVB_CHECK_ARG(((uintptr_t)W1 % 16 == 0) && ((uintptr_t)W2 % 16 == 0), "pvb_sdec_tc_step: weights must be 16-byte aligned");
  PVB_CHECK_ARG(!backward || ((uintptr_t)wgrad_part % 16 == 0), "pvb_sdec_tc_step: wgrad_part must be 16-byte aligned");
  PVB_CHECK_ARG((uintptr_t)packed_w % 16 == 0, "pvb_sdec_tc_step: packed weights must be 16-byte aligned");
  const int N = (ndim == 1) ? H : H * W;
  PVB_CHECK_ARG(N >= MIN_PIX, "pvb_sdec_tc_step: need >= 32 pixels per instance");
  PVB_CHECK_ARG(B < (1LL << 30) && I < (1LL << 40), "pvb_sdec_tc_step: batch too large");
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
  P.Wp = packed_w;
  P.rowll = rowll; P.loc = loc; P.gUv_part = gUv_part; P.wgrad_part = wgrad_part;
  P.R = I * N; P.B = B; P.N = N; P.H = H; P.W = (ndim == 1) ? 1 : W; P.ndim = ndim;
  P.sampler = sampler; P.sigmoid_d = sigmoid_d; P.sig = decoder_sig; P.backward = backward;
  P.tiles = s.tiles;
  const int64_t step = (int64_t)TILE * s.ctas;
  P.step_q = step / N;
  P.step_r = (int)(step % N);
  P.step_qb = (int)(P.step_q % B);
  // PVB_SDEC_V2=1: the experimental interleaved kernel (two tiles in flight, csrc/pvb_sdec_tc2.cu;
  // measured slower than this one on B200) for the training step; default: this file
  const char* v2_env = std::getenv("PVB_SDEC_V2");
  if (backward && v2_env && v2_env[0] == '1') {
    int e = pvb_sdec::launch_v2(P, s.ctas, (cudaStream_t)stream);
    if (e != 0) return e;
    pvb::count_launch();
    return pvb::launch_status();
  }
  sdec_tc_kernel<<<s.ctas, NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(P);
  pvb::count_launch();
  return pvb::launch_status();
}

extern "C" int pvb_sdec_tc_gather_gUv(const float* gUv_part, float* gUv, int64_t I, int N,
                                      void* stream) {
  PVB_CHECK_ARG(gUv_part && gUv && I >= 0 && N >= MIN_PIX, "pvb_sdec_tc_gather_gUv: bad argument");
  if (I == 0) return 0;
  gather_gUv_kernel<<<(unsigned)I, 128, 0, (cudaStream_t)stream>>>(gUv_part, gUv, I, N);
  pvb::count_launch();
  return pvb::launch_status();
}
