// Convolutions of the VED nets on tcgen05 tensor cores (sm_100a): implicit GEMM with the
// accumulator in tensor memory, for layers with >= 16 channels on both sides (the 3x3 / 1x1
// convolutions that carry ~99 % of VED's FLOPs; reference nets/conv.py:146-249).  The activations
// stay NCHW fp32 in HBM (same buffers as the fp32 kernels of pvb_conv.cu): operand tiles are
// gathered plane by plane (a warp reads 32 consecutive pixels of one channel plane = one 128-byte
// line), converted and packed on the fly into the no-swizzle "row-chunk" UMMA layout of umma.cuh,
// so there is no im2col buffer and no layout-conversion pass.
//
//   pixel GEMM  (forward, backward-data):  M = 128 pixels / CTA, N = output channels (<= 128),
//               K = taps x gathered channels, chunked as (tap, 64 channels);
//               4 producer/epilogue warps (thread = pixel row = TMEM lane) + 1 MMA warp,
//               3-stage smem ring, 2 CTAs per SM.
//   weight GEMM (backward-weight): dW[co][(tap,ci)] = sum_px dpre[px][co] x[px+tap][ci]:
//               M = Cout (padded to 128 lanes), N = Cin per tap, K = 128 pixels per step, both
//               operands MN-major (K = tile rows, exactly as the decoder kernel's dW GEMMs);
//               accumulators for a group of taps live in TMEM over the CTA's pixel range and are
//               added to the fp32 gradient with atomics at the end.
// Operands: fp16 (bf16 is a compile-time switch for the backward GEMMs); fp32 accumulate.
#include <cuda_bf16.h>
#include "pvb_common.cuh"
#include "umma.cuh"

namespace {

constexpr int TP = 128;                 // pixels (rows) per tile
constexpr int CC = 64;                  // gathered channels per K chunk
constexpr int ROWB = 16;                // bytes of one row of a chunk-column (8 x 16-bit)
constexpr int PIX_STAGES = 2;   // 64 KB per CTA, 2 CTAs per SM: leaves ~100 KB of L1 for the tap re-reads
constexpr int A_STAGE = TP * CC * 2;    // 16 KB
constexpr int B_STAGE = 128 * CC * 2;   // 16 KB (N <= 128)
constexpr int PIX_SMEM = PIX_STAGES * (A_STAGE + B_STAGE) + 64;
constexpr int PIX_THREADS = 160;        // 4 producer/epilogue warps + 1 MMA warp

struct TcDims {
  int B, Cg, Nout, H, W, kh, kw;        // Cg: gathered channels (K side), Nout: output channels
  int sign;                             // +1 forward taps, -1 mirrored (backward data)
  float in_scale, out_scale;            // gathered values * in_scale (clamped), results * out_scale
};

// Activation gradients span ~1e-6 .. 1 per element (deep layers of the decoder are small): below
// fp16's normal range (6e-5) they would lose mantissa bits, so the backward GEMMs run on
// GRAD_SCALE * dpre (a power of two: exact) and un-scale the fp32 result; the clamp keeps an
// outlier finite instead of poisoning the sum with inf.
constexpr float GRAD_SCALE = 1024.f;
constexpr float F16_MAX = 60000.f;
__device__ __forceinline__ float scl(float v, float s) { return fminf(fmaxf(v * s, -F16_MAX), F16_MAX); }

template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if (BF16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__host__ __device__ constexpr uint32_t idesc_16b(int M, int N, int a_mn, int b_mn, bool bf16) {
  return umma::idesc_f16(M, N, a_mn, b_mn) | (bf16 ? ((1u << 7) | (1u << 10)) : 0u);
}

// weights fp32 [Cout][Cin][taps] -> 16-bit [tap][n][k]:
//   mode 0 (forward):        n = co, k = ci      mode 1 (backward data): n = ci, k = co
template <bool BF16>
__global__ void conv_tc_prep_kernel(const float* __restrict__ W, uint16_t* __restrict__ Wp, int Cout,
                                    int Cin, int taps, int mode) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t total = (int64_t)Cout * Cin * taps;
  if (i >= total) return;
  int Nn = mode == 0 ? Cout : Cin, Kk = mode == 0 ? Cin : Cout;
  int k = (int)(i % Kk);
  int n = (int)((i / Kk) % Nn);
  int t = (int)(i / ((int64_t)Kk * Nn));
  int co = mode == 0 ? n : k, ci = mode == 0 ? k : n;
  float v = W[((int64_t)co * Cin + ci) * taps + t];
  uint32_t p = pack2<BF16>(v, 0.f);
  Wp[i] = (uint16_t)(p & 0xffffu);
}

// gather NCH (compile-time: 16 / 32 / 48 / 64) channels of one pixel (plane stride hw floats) into
// the row-chunk tile: one uniform bounds test, every load of the chunk issued before the first
// conversion (one memory latency per chunk), 32-bit offsets from one base pointer
template <bool BF16, bool SCALED, int NCH>
__device__ __forceinline__ void gather_n(const float* __restrict__ p, int hw, bool ok, float scale,
                                         uint8_t* dst) {
  // `p` always points at readable memory (callers pass the un-shifted pixel when !ok): the loads
  // are unconditional and the result is masked, so border lanes do not diverge
  float v[NCH];
  const float m = ok ? 1.f : 0.f;
#pragma unroll
  for (int j = 0; j < NCH; ++j) v[j] = __ldg(p + j * hw);
  if (!SCALED) {
#pragma unroll
    for (int j = 0; j < NCH; ++j) v[j] *= m;
  } else {
    scale *= m;
  }
#pragma unroll
  for (int c8 = 0; c8 < NCH / 8; ++c8) {
    float* w = v + c8 * 8;
    if (SCALED) {
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = scl(w[j], scale);
    }
    *reinterpret_cast<uint4*>(dst + c8 * (TP * ROWB)) =
        make_uint4(pack2<BF16>(w[0], w[1]), pack2<BF16>(w[2], w[3]), pack2<BF16>(w[4], w[5]),
                   pack2<BF16>(w[6], w[7]));
  }
}
template <bool BF16, bool SCALED>
__device__ __forceinline__ void gather_row(const float* __restrict__ p, int64_t HW, int cc, bool ok,
                                           float scale, uint8_t* dst) {
  const int hw = (int)HW;
  switch (cc) {   // warp-uniform
    case 64: gather_n<BF16, SCALED, 64>(p, hw, ok, scale, dst); break;
    case 48: gather_n<BF16, SCALED, 48>(p, hw, ok, scale, dst); break;
    case 32: gather_n<BF16, SCALED, 32>(p, hw, ok, scale, dst); break;
    case 16: gather_n<BF16, SCALED, 16>(p, hw, ok, scale, dst); break;
    default: break;   // channel chunks are multiples of 16 (checked on the host)
  }
}

// bias + activation on 16 accumulator columns, the activation switch hoisted out of the loop
__device__ __forceinline__ void bias_act16(float* v, const float* __restrict__ bias, int n0, float oscale,
                                           int act, float* __restrict__ pre_out, int64_t HW) {
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = fmaf(v[j], oscale, bias ? __ldg(bias + n0 + j) : 0.f);
  if (pre_out) {
#pragma unroll
    for (int j = 0; j < 16; ++j) pre_out[(int64_t)(n0 + j) * HW] = v[j];
  }
  switch (act) {
    case PVB_ACT_NONE: break;
    case PVB_ACT_LRELU:
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = v[j] > 0.f ? v[j] : 0.01f * v[j];
      break;
    case PVB_ACT_RELU:
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
      break;
    default:
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = pvb::act_fwd(v[j], act);
  }
}

#ifdef PVB_TC_TRACE
__device__ long long g_ctrace[2][64];
#define CTRACE(role, ev) do { if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 4) && (ev) < 64) g_ctrace[role][ev] = clock64(); } while (0)
__device__ long long g_wtrace[2][64];
#define WTRACE(role, ev) do { if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && (warp == 0 || warp == mma_warp) && (ev) < 64) g_wtrace[role][ev] = clock64(); } while (0)
#else
#define CTRACE(role, ev) do {} while (0)
#define WTRACE(role, ev) do {} while (0)
#endif

template <bool BF16, bool SCALED>
__global__ void __launch_bounds__(PIX_THREADS, 2)
conv_tc_pix_kernel(const float* __restrict__ src, const uint16_t* __restrict__ Wp,
                   const float* __restrict__ bias, float* __restrict__ dst, float* __restrict__ pre,
                   TcDims d, int act) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = smem + PIX_STAGES * A_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PIX_STAGES * (A_STAGE + B_STAGE));
  uint64_t* full = bars;                  // [3] producers -> MMA   (count 4: one arrive per warp)
  uint64_t* empty = bars + PIX_STAGES;    // [3] MMA -> producers   (tcgen05.commit)
  uint64_t* accb = bars + 2 * PIX_STAGES; // accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * PIX_STAGES + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int taps = d.kh * d.kw, ph = d.kh / 2, pw = d.kw / 2;
  const int HW = d.H * d.W;
  const int64_t Mtot = (int64_t)d.B * HW;
  const int cchunks = (d.Cg + CC - 1) / CC;       // channel chunks per tap
  const int n_chunks = taps * cchunks;
  const int Nout = d.Nout;                        // multiple of 16, <= 128
  if (warp == 4) umma::tmem_alloc<128>(tmem_slot);
  if (tid == 0) {
    for (int s = 0; s < PIX_STAGES; ++s) {
      umma::mbar_init(full + s, 4);
      umma::mbar_init(empty + s, 1);
    }
    umma::mbar_init(accb, 1);
    umma::mbar_fence_init();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = *tmem_slot;
  CTRACE(warp == 4 ? 1 : 0, 0);

  if (warp == 4) {
    // ================= MMA issuer =================
    const uint32_t idesc = idesc_16b(128, Nout, 0, 0, BF16);
    const uint32_t a0 = umma::smem_u32(sA), b0 = umma::smem_u32(sB);
    for (int c = 0; c < n_chunks; ++c) {
      const int s = c % PIX_STAGES;
      const uint32_t ph_full = (c / PIX_STAGES) & 1;
      umma::mbar_wait(full + s, ph_full);
      CTRACE(1, 1 + 2 * c);
      umma::fence_after_sync();
      if (lane == 0) {
        const int cc = min(CC, d.Cg - (c % cchunks) * CC);   // channels in this chunk (multiple of 16)
        for (int k16 = 0; k16 < cc / 16; ++k16) {
          uint64_t da = umma::smem_desc(a0 + s * A_STAGE + k16 * 2 * (TP * ROWB), TP * ROWB, 128);
          uint64_t db = umma::smem_desc(b0 + s * B_STAGE + k16 * 2 * (Nout * ROWB), Nout * ROWB, 128);
          umma::mma_f16_ss(tm, da, db, idesc, (c > 0 || k16 > 0) ? 1u : 0u);
        }
        umma::commit(empty + s);
        if (c == n_chunks - 1) umma::commit(accb);
      }
      __syncwarp();
      CTRACE(1, 2 + 2 * c);
    }
  } else {
    // ================= producers (then epilogue) =================
    const int row = tid;                              // pixel row of the tile == TMEM lane
    const int64_t gm = (int64_t)blockIdx.x * TP + row;
    const bool m_ok = gm < Mtot;
    const int gb = m_ok ? (int)(gm / HW) : 0;
    const int gr = m_ok ? (int)(gm - (int64_t)gb * HW) : 0;
    const int gh = gr / d.W, gw = gr - gh * d.W;
    const float* gbase = src + (int64_t)gb * d.Cg * HW + gr;
    for (int c = 0; c < n_chunks; ++c) {
      const int s = c % PIX_STAGES;
      if (c >= PIX_STAGES) {
        umma::mbar_wait(empty + s, ((c / PIX_STAGES) - 1) & 1);   // MMAs of chunk c - STAGES done
      }
      const int tap = c / cchunks, c0 = (c % cchunks) * CC;
      const int cc = min(CC, d.Cg - c0);
      const int dh = d.sign * (tap / d.kw - ph), dw = d.sign * (tap % d.kw - pw);
      const int hh = gh + dh, ww = gw + dw;
      const bool ok = m_ok && hh >= 0 && hh < d.H && ww >= 0 && ww < d.W;
      const float* p = gbase + (int64_t)c0 * HW + (ok ? dh * d.W + dw : 0);
      uint8_t* a_dst = sA + s * A_STAGE + row * ROWB;
      CTRACE(0, 1 + 3 * c);
      // weights of this (tap, channel chunk): Wp[tap][n][c0 .. c0+cc), row n = tid; loaded first so
      // that they are in flight together with the activation gather
      uint4 wq[8];
      const int wrow_n = row < Nout ? row : 0;
      const uint16_t* wrow = Wp + ((int64_t)tap * Nout + wrow_n) * d.Cg + c0;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8)
        if (c8 * 8 < cc) wq[c8] = __ldg(reinterpret_cast<const uint4*>(wrow + c8 * 8));
      gather_row<BF16, SCALED>(p, HW, cc, ok, d.in_scale, a_dst);
      CTRACE(0, 2 + 3 * c);
      if (row < Nout) {
        uint8_t* b_dst = sB + s * B_STAGE + row * ROWB;
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8)
          if (c8 * 8 < cc) *reinterpret_cast<uint4*>(b_dst + c8 * (Nout * ROWB)) = wq[c8];
      }
      umma::fence_proxy_async();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(full + s);
      CTRACE(0, 3 + 3 * c);
    }
    // ---- epilogue: TMEM -> bias + activation -> NCHW fp32 ----
    umma::mbar_wait(accb, 0);
    CTRACE(0, 60);
    umma::fence_after_sync();
    const uint32_t tm_lane = tm + ((uint32_t)(warp * 32) << 16);
    float* obase = dst + (int64_t)gb * Nout * HW + gr;
    float* pbase = pre ? pre + (int64_t)gb * Nout * HW + gr : nullptr;
    for (int n0 = 0; n0 < Nout; n0 += 16) {
      float v[16];
      umma::tmem_ld16(tm_lane + n0, v);
      umma::tmem_ld_wait();
      if (m_ok) {
        bias_act16(v, bias, n0, d.out_scale, act, pbase, HW);
        float* o = obase + (int64_t)n0 * HW;
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j * HW] = v[j];
      }
    }
    umma::fence_before_sync();
    CTRACE(0, 61);
  }
  __syncthreads();
  if (warp == 4) umma::tmem_dealloc<128>(tm);
}

// ---- backward weight ------------------------------------------------------------------------------
// CTA (tap group, pixel split): accumulators [128 lanes = co][taps_in_group x Cin columns] in TMEM.
// Per 128-pixel step: A = dpre tile [128 px][128 co (zero padded)] (MN-major), B_t = x tile shifted by
// tap t [128 px][Cin] (MN-major), one N = Cin MMA chain (8 K-steps of 16 pixels) per tap.
constexpr int WG_MAX_GROUPS = 4;                // producer groups of 4 warps (thread = pixel row)
#ifndef PVB_WG_PAD
#define PVB_WG_PAD 0
#endif
// chunk-column stride of the MN-major operand tiles: 128 rows x 16 bytes (+ optional padding; measured:
// a 16-byte pad, which spreads the 16-byte pieces of one K row over the banks, changes nothing)
constexpr int WCS = TP * ROWB + PVB_WG_PAD;
constexpr int WG_A = 16 * WCS;                   // dpre tile, 16 chunk-columns (128 output channels)
constexpr int WG_MAX_TAPS = 9;                   // taps per CTA (<= (512 - 16) / Cin)
constexpr int WG_STAGES = 2;
struct WgItem { int src_off, shift, dst; short dh, dw; int is_x; };
constexpr int WG_TBL_BYTES = (8 + 9 * 16) * (int)sizeof(WgItem);   // item table: <= 8 + 9 * 16 entries

// The gathers are latency-bound (strided 4-byte loads, 16 in flight per thread), so the work of one
// 128-pixel step is cut into 16-channel items -- 8 for the dpre tile, Cin/16 per tap for the shifted
// x tiles -- dealt round-robin to up to four producer groups of 128 threads: up to 4x the loads in
// flight per SM, same two-stage ring towards the MMA warp.
template <bool BF16>
__global__ void __launch_bounds__(WG_MAX_GROUPS * 128 + 32, 1)
conv_tc_wgrad_kernel(const float* __restrict__ dpre, const float* __restrict__ x, float* __restrict__ dW,
                     float* __restrict__ db, int B, int Cin_real, int Cout, int H, int W, int kh, int kw,
                     int taps_per_cta, int64_t tiles_per_split) {
  // fewer than 16 input channels (the first layer): the x tile is zero-padded to 16 columns
  const int Cin = Cin_real < 16 ? 16 : Cin_real;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int b_tile = (Cin / 8) * WCS;                    // one shifted x tile
  const int stage_bytes = WG_A + taps_per_cta * b_tile + 2 * WCS;   // + ones tile (16 columns)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + WG_STAGES;
  uint64_t* accb = bars + 2 * WG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_STAGES + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_groups = (int)(blockDim.x >> 7);             // producer groups; the MMA warp comes last
  const int mma_warp = n_groups * 4;
  const int taps = kh * kw, ph = kh / 2, pw = kw / 2;
  const int HW = H * W;
  const int64_t Mtot = (int64_t)B * HW;
  const int64_t n_tiles = (Mtot + TP - 1) / TP;
  const int tap0 = blockIdx.x * taps_per_cta;
  const int ntap = min(taps_per_cta, taps - tap0);
  // pixel tiles are dealt round-robin to the gridDim.y CTAs of a tap group (tile = blockIdx.y +
  // it * gridDim.y): at any moment the CTAs read one contiguous window of every channel plane, which
  // keeps DRAM pages open across CTAs (a contiguous range per CTA made each CTA a separate stream of
  // 512-byte pieces per plane)
  const int64_t my_tiles = blockIdx.y < n_tiles ? (n_tiles - blockIdx.y + gridDim.y - 1) / gridDim.y : 0;
  const bool do_bias = db != nullptr && blockIdx.x == 0;
  const uint32_t col_bias = (uint32_t)(ntap * Cin);        // TMEM column block of the bias sums
  if (warp == mma_warp) umma::tmem_alloc<512>(tmem_slot);
  if (tid == 0) {
    for (int s = 0; s < WG_STAGES; ++s) {
      umma::mbar_init(full + s, (uint32_t)mma_warp);       // one arrive per producer warp
      umma::mbar_init(empty + s, 1);
    }
    umma::mbar_init(accb, 1);
    umma::mbar_fence_init();
  }
  // gather items of one 128-pixel step (16 channels each): Cout/16 blocks of the dpre tile, then
  // Cin/16 blocks of each tap's shifted x tile; decoded once into shared memory
  WgItem* tbl = reinterpret_cast<WgItem*>(smem + WG_STAGES * stage_bytes + 64);
  const int a_items = Cout / 16, b_items = Cin / 16;
  const int n_items = a_items + ntap * b_items;
  for (int k = tid; k < n_items; k += blockDim.x) {
    WgItem e;
    if (k < a_items) {
      e.src_off = k * 16 * HW;
      e.shift = 0; e.dh = 0; e.dw = 0; e.is_x = 0;
      e.dst = k * 2 * WCS;
    } else {
      const int kk = k - a_items, t = kk / b_items, cb = kk - t * b_items;
      const int tap = tap0 + t;
      e.dh = (short)(tap / kw - ph);
      e.dw = (short)(tap % kw - pw);
      e.shift = e.dh * W + e.dw;
      e.src_off = cb * 16 * HW;
      e.is_x = 1;
      e.dst = WG_A + t * b_tile + cb * 2 * WCS;
    }
    tbl[k] = e;
  }
  if (warp < mma_warp) {
    // output channels beyond Cout never change: zero those dpre chunk-columns once, in every stage
    for (int s = 0; s < WG_STAGES; ++s)
      for (int c8 = Cout / 8 + (warp >> 2); c8 < 16; c8 += n_groups)
        *reinterpret_cast<uint4*>(smem + s * stage_bytes + c8 * WCS + (tid & 127) * ROWB) =
            make_uint4(0u, 0u, 0u, 0u);
    umma::fence_proxy_async();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = *tmem_slot;

  if (warp == mma_warp) {
    // The tap tiles (and the ones tile behind them) are contiguous chunk-columns in shared memory
    // and their accumulators contiguous TMEM columns, so one MMA spans several taps (N <= 256):
    // the dpre operand (128 x 16, 4 KB) is read once per K-step instead of once per tap -- these
    // small-N MN-major MMAs are bound by operand reads from shared memory, not by the math.
    const int n_total = ntap * Cin + (do_bias ? 16 : 0);
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int s = (int)(it % WG_STAGES);
      WTRACE(1, (int)(3 * it));
      umma::mbar_wait(full + s, (uint32_t)((it / WG_STAGES) & 1));
      WTRACE(1, (int)(3 * it + 1));
      umma::fence_after_sync();
      if (lane == 0) {
        const uint32_t base = umma::smem_u32(smem + s * stage_bytes);
        for (int n0 = 0; n0 < n_total; n0 += 256) {
          const int nn = min(256, n_total - n0);
          const uint32_t idesc = idesc_16b(128, nn, 1, 1, BF16);
          const uint32_t bt = base + WG_A + (n0 / 8) * WCS;
          for (int k = 0; k < 8; ++k)   // 8 K-steps of 16 pixel rows
            umma::mma_f16_ss(tm + n0, umma::smem_desc(base + k * 256, 128, WCS),
                             umma::smem_desc(bt + k * 256, 128, WCS), idesc,
                             (it > 0 || k > 0) ? 1u : 0u);
        }
        umma::commit(empty + s);
        if (it == my_tiles - 1) umma::commit(accb);
      }
      __syncwarp();
      WTRACE(1, (int)(3 * it + 2));
    }
  } else {
    const int row = tid & 127, grp = warp >> 2;
    const int n_mine = (n_items - grp + n_groups - 1) / n_groups;    // >= 1 (host: groups <= items)
    const int64_t total = my_tiles * n_mine;
    const bool small_c = Cin_real < 16;
    // ---- load cursor: tile geometry kept incrementally (no 64-bit divisions in the loop) ----
    const int64_t adv = (int64_t)gridDim.y * TP;            // pixels between two tiles of this CTA
    const int adv_b = (int)(adv / HW), adv_r = (int)(adv - (int64_t)adv_b * HW);
    int l_gb, l_gr, l_gh, l_gw, l_i = 0;
    bool l_ok;
    const float *l_dp, *l_xp;
    {
      const int64_t gm = (int64_t)blockIdx.y * TP + row;
      l_gb = (int)(gm / HW);
      l_gr = (int)(gm - (int64_t)l_gb * HW);
    }
    auto tile_geometry = [&]() {
      l_ok = l_gb < B;
      const int gb = l_ok ? l_gb : 0;
      l_gh = l_gr / W;
      l_gw = l_gr - l_gh * W;
      l_dp = dpre + (int64_t)gb * Cout * HW + l_gr;
      l_xp = x + (int64_t)gb * Cin_real * HW + l_gr;
    };
    tile_geometry();
    // One 16-channel item: 16 loads issued back to back; converted / stored one item later, so the
    // loads of the next item (possibly of the next tile) are in flight meanwhile.
    struct Pending { float v[16]; uint32_t dst; bool ok, is_x; };
    auto issue = [&](Pending& pd) {
      const WgItem e = tbl[grp + l_i * n_groups];
      pd.is_x = e.is_x != 0;
      pd.ok = l_ok && (!pd.is_x || ((unsigned)(l_gh + e.dh) < (unsigned)H &&
                                    (unsigned)(l_gw + e.dw) < (unsigned)W));
      pd.dst = (uint32_t)e.dst + (uint32_t)(row * ROWB);
      // `p` always points at readable memory (the un-shifted pixel when the tap falls outside)
      const float* p = (pd.is_x ? l_xp : l_dp) + e.src_off + (pd.ok ? e.shift : 0);
      if (small_c && pd.is_x) {
#pragma unroll
        for (int j = 0; j < 16; ++j) pd.v[j] = j < Cin_real ? __ldg(p + (int64_t)j * HW) : 0.f;
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) pd.v[j] = __ldg(p + (int64_t)j * HW);
      }
      if (++l_i == n_mine) {       // next item belongs to the next tile of this CTA
        l_i = 0;
        l_gr += adv_r;
        l_gb += adv_b;
        if (l_gr >= HW) { l_gr -= HW; ++l_gb; }
        tile_geometry();
      }
    };
    // ---- store cursor ----
    int64_t s_it = 0;
    int s_i = 0;
    int tr_ev = 0;      // trace event counter (debug builds only)
    const __half2 hmax = __floats2half2_rn(F16_MAX, F16_MAX), hmin = __floats2half2_rn(-F16_MAX, -F16_MAX);
    auto finish = [&](Pending& pd) {
      const int s = (int)(s_it % WG_STAGES);
      uint8_t* st = smem + s * stage_bytes;
      WTRACE(0, tr_ev++);
      if (s_i == 0 && s_it >= WG_STAGES)
        umma::mbar_wait(empty + s, (uint32_t)(((s_it / WG_STAGES) - 1) & 1));   // stage free again
      WTRACE(0, tr_ev++);
      uint32_t pk[8];
      if (pd.is_x) {
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = pack2<BF16>(pd.v[2 * j], pd.v[2 * j + 1]);
      } else if (BF16) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          pk[j] = pack2<true>(scl(pd.v[2 * j], GRAD_SCALE), scl(pd.v[2 * j + 1], GRAD_SCALE));
      } else {
        // scale (a power of two) in fp32, saturate after the conversion on packed pairs
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          __half2 h = __floats2half2_rn(pd.v[2 * j] * GRAD_SCALE, pd.v[2 * j + 1] * GRAD_SCALE);
          h = __hmin2(__hmax2(h, hmin), hmax);
          pk[j] = *reinterpret_cast<uint32_t*>(&h);
        }
      }
      if (!pd.ok) {
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = 0u;
      }
      *reinterpret_cast<uint4*>(st + pd.dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(st + pd.dst + WCS) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      WTRACE(0, tr_ev++);
      if (++s_i == n_mine) {
        if (do_bias && grp == n_groups - 1) {
          // ones tile [128 px][16]: column 0 = 1 for valid pixels -> bias sums in one N = 16 chain
          const int64_t gm = ((int64_t)blockIdx.y + s_it * gridDim.y) * TP + row;
          uint8_t* bo = st + WG_A + ntap * b_tile;
          uint32_t one = pack2<BF16>(gm < Mtot ? 1.f : 0.f, 0.f);
          *reinterpret_cast<uint4*>(bo + row * ROWB) = make_uint4(one, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(bo + WCS + row * ROWB) = make_uint4(0u, 0u, 0u, 0u);
        }
        umma::fence_proxy_async();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(full + s);
        s_i = 0;
        ++s_it;
      }
    };
    Pending pa, pb;
    if (total > 0) issue(pa);
    for (int64_t q = 0; q < total; q += 2) {
      if (q + 1 < total) issue(pb);
      finish(pa);
      if (q + 2 < total) issue(pa);
      if (q + 1 < total) finish(pb);
    }
    if (my_tiles > 0 && grp == 0) {
      umma::mbar_wait(accb, 0);
      umma::fence_after_sync();
      const uint32_t tm_lane = tm + ((uint32_t)(warp * 32) << 16);
      const int co = row;                                   // TMEM lane = output channel
      for (int t = 0; t < ntap; ++t) {
        const int tap = tap0 + t;
        for (int n0 = 0; n0 < Cin; n0 += 16) {
          float v[16];
          umma::tmem_ld16(tm_lane + t * Cin + n0, v);
          umma::tmem_ld_wait();
          if (co < Cout) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n0 + j < Cin_real)
                atomicAdd(dW + ((int64_t)co * Cin_real + n0 + j) * taps + tap, v[j] * (1.f / GRAD_SCALE));
          }
        }
      }
      if (do_bias) {
        float v[16];
        umma::tmem_ld16(tm_lane + col_bias, v);
        umma::tmem_ld_wait();
        if (co < Cout) atomicAdd(db + co, v[0] * (1.f / GRAD_SCALE));
      }
      umma::fence_before_sync();
    }
  }
  __syncthreads();
  if (warp == mma_warp) umma::tmem_dealloc<512>(tm);
}

// operand type of the backward GEMMs.  bf16 has the range of fp32 but 8 mantissa bits: measured
// 3 % error on the first-layer weight gradient after five backward convolutions; fp16 (11 bits)
// gives 0.4 %, and per-element activation gradients of this loss (a batch SUM of per-pixel
// log-likelihoods, |dlogit| <= 1) sit far inside fp16's range (subnormals down to 6e-8).
constexpr bool BWD_BF16 = false;

bool tc_ok(int Cg, int Nout, int kh, int kw) {
  return Cg >= 16 && Cg % 16 == 0 && Nout >= 16 && Nout % 16 == 0 && Nout <= 128 && Cg <= 256 &&
         (kh == 1 || kh == 3) && (kw == 1 || kw == 3);
}

}  // namespace

extern "C" int pvb_conv_tc_supported(int Cin, int Cout, int kh, int kw) {
  return tc_ok(Cin, Cout, kh, kw) && tc_ok(Cout, Cin, kh, kw) ? 1 : 0;
}
// the weight-gradient kernel alone also takes layers with fewer than 16 input channels
extern "C" int pvb_conv_tc_wgrad_supported(int Cin, int Cout, int kh, int kw) {
  return tc_ok(Cin < 16 ? 16 : Cin, Cout, kh, kw) && Cout <= 128 ? 1 : 0;
}

extern "C" int64_t pvb_conv_tc_workspace_bytes(int Cin, int Cout, int kh, int kw) {
  return (int64_t)Cin * Cout * kh * kw * 2;
}

// mode 0: y = act(conv(x, W) + b)   (fp16 operands)     src = x   [B, Cin, H, W]
// mode 1: dx = conv_transpose(dpre, W)  (bf16 operands)  src = dpre [B, Cout, H, W]
extern "C" int pvb_conv_tc_pix(const float* src, const float* W, const float* b, float* dst, float* pre,
                               void* workspace, int B, int Cin, int Cout, int H, int Wd, int kh, int kw,
                               int act, int mode, void* stream) {
  PVB_CHECK_ARG(src && W && dst && workspace, "pvb_conv_tc_pix: null pointer");
  PVB_CHECK_ARG(mode == 0 || mode == 1, "pvb_conv_tc_pix: mode must be 0 (forward) or 1 (backward data)");
  const int Cg = mode == 0 ? Cin : Cout, Nout = mode == 0 ? Cout : Cin;
  PVB_CHECK_ARG(tc_ok(Cg, Nout, kh, kw), "pvb_conv_tc_pix: unsupported shape (channels must be multiples of 16, N <= 128)");
  PVB_CHECK_ARG(((uintptr_t)workspace % 16) == 0, "pvb_conv_tc_pix: workspace must be 16-byte aligned");
  if (B == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int taps = kh * kw;
  const int64_t total = (int64_t)Cin * Cout * taps;
  uint16_t* Wp = reinterpret_cast<uint16_t*>(workspace);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(conv_tc_pix_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PIX_SMEM);
    cudaFuncSetAttribute(conv_tc_pix_kernel<BWD_BF16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PIX_SMEM);
    attr = true;
  }
  TcDims d{B, Cg, Nout, H, Wd, kh, kw, mode == 0 ? 1 : -1,
           mode == 0 ? 1.f : GRAD_SCALE, mode == 0 ? 1.f : 1.f / GRAD_SCALE};
  const int64_t M = (int64_t)B * H * Wd;
  const unsigned grid = (unsigned)((M + TP - 1) / TP);
  if (mode == 0) {
    conv_tc_prep_kernel<false><<<pvb::cdiv(total, 256), 256, 0, st>>>(W, Wp, Cout, Cin, taps, 0);
    pvb::count_launch();
    conv_tc_pix_kernel<false, false><<<grid, PIX_THREADS, PIX_SMEM, st>>>(src, Wp, b, dst, pre, d, act);
  } else {
    conv_tc_prep_kernel<BWD_BF16><<<pvb::cdiv(total, 256), 256, 0, st>>>(W, Wp, Cout, Cin, taps, 1);
    pvb::count_launch();
    conv_tc_pix_kernel<BWD_BF16, true><<<grid, PIX_THREADS, PIX_SMEM, st>>>(src, Wp, nullptr, dst, nullptr, d, 0);
  }
  pvb::count_launch();
  return pvb::launch_status();
}

extern "C" int pvb_conv_tc_wgrad(const float* dpre, const float* x, float* dW, float* db, int B, int Cin,
                                 int Cout, int H, int Wd, int kh, int kw, void* stream) {
  PVB_CHECK_ARG(dpre && x && dW, "pvb_conv_tc_wgrad: null pointer");
  const int Cin_real = Cin;
  if (Cin < 16) Cin = 16;                        // zero-padded x tile
  PVB_CHECK_ARG(tc_ok(Cin, Cout, kh, kw) && Cout <= 128, "pvb_conv_tc_wgrad: unsupported shape");
  if (B == 0) return 0;
  const int taps = kh * kw;
  int tpc = (512 - 16) / Cin;                    // taps per CTA: accumulators + 16 bias columns <= 512
  if (tpc > WG_MAX_TAPS) tpc = WG_MAX_TAPS;
  if (tpc > taps) tpc = taps;
  {
    // two smem stages of (dpre tile + tpc shifted x tiles + ones tile) must fit
    int by_smem = ((227 * 1024 - 64 - WG_TBL_BYTES) / WG_STAGES - WG_A - 2 * WCS) / ((Cin / 8) * WCS);
    if (tpc > by_smem) tpc = by_smem;
  }
  PVB_CHECK_ARG(tpc >= 1, "pvb_conv_tc_wgrad: Cin too large");
  const int groups = (taps + tpc - 1) / tpc;
  const int stage = WG_A + tpc * (Cin / 8) * WCS + 2 * WCS;
  const int smem = WG_STAGES * stage + 64 + WG_TBL_BYTES;
  PVB_CHECK_ARG((int64_t)(Cin > Cout ? Cin : Cout) * H * Wd < (1ll << 31),
                "pvb_conv_tc_wgrad: one image's activations must have fewer than 2^31 elements");
  PVB_CHECK_ARG(smem <= 227 * 1024, "pvb_conv_tc_wgrad: tile does not fit shared memory");
  const int64_t M = (int64_t)B * H * Wd;
  const int64_t n_tiles = (M + TP - 1) / TP;
  // one CTA per SM (512 TMEM columns, ~200 KB smem): never more CTAs than SMs, or the extras
  // run as a second wave and double the kernel time
  int64_t splits = 148 / groups;
  if (splits < 1) splits = 1;
  if (splits > n_tiles) splits = n_tiles;
  const int64_t per = (n_tiles + splits - 1) / splits;
  splits = (n_tiles + per - 1) / per;
  static int attr_smem = 0;
  if (smem > attr_smem) {
    cudaFuncSetAttribute(conv_tc_wgrad_kernel<BWD_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_smem = 227 * 1024;
  }
  dim3 grid(groups, (unsigned)splits);
  // 16-channel gather items per 128-pixel step -> producer groups (no more groups than items)
  const int items = Cout / 16 + tpc * (Cin / 16);
  const int ng = items < WG_MAX_GROUPS ? (items < 1 ? 1 : items) : WG_MAX_GROUPS;
  conv_tc_wgrad_kernel<BWD_BF16><<<grid, ng * 128 + 32, smem, (cudaStream_t)stream>>>(
      dpre, x, dW, db, B, Cin_real, Cout, H, Wd, kh, kw, tpc, per);
  pvb::count_launch();
  return pvb::launch_status();
}

#ifdef PVB_TC_TRACE
extern "C" int pvb_wgrad_trace_read(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_wtrace, sizeof(long long) * 2 * 64);
}
extern "C" int pvb_conv_trace_read(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_ctrace, sizeof(long long) * 2 * 64);
}
#endif

