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
#include <cstdlib>
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

// scaled pair -> 16-bit pair, saturated: fp16 in ONE instruction (F2FP.SATFINITE clamps to the largest finite
// value; the fmin/fmax pair per element it replaces was a fifth of the gather instructions)
template <bool BF16>
__device__ __forceinline__ uint32_t pack2s(float a, float b, float s) {
  if (BF16) return pack2<true>(scl(a, s), scl(b, s));
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b * s), "f"(a * s));   // {upper, lower}
  return r;
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
    *reinterpret_cast<uint4*>(dst + c8 * (TP * ROWB)) =
        SCALED ? make_uint4(pack2s<BF16>(w[0], w[1], scale), pack2s<BF16>(w[2], w[3], scale),
                            pack2s<BF16>(w[4], w[5], scale), pack2s<BF16>(w[6], w[7], scale))
               : make_uint4(pack2<BF16>(w[0], w[1]), pack2<BF16>(w[2], w[3]), pack2<BF16>(w[4], w[5]),
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

// v[j] *= act'(y[j]) on 16 columns, the activation switch hoisted out of the loop (per element it costs ~20
// instructions on the four epilogue warps: the backward-data kernel ran 51 M warp instructions against the
// forward's 32 M and took 2.9x its time)
__device__ __forceinline__ void act_grad16(float* v, const float* y, int act) {
  switch (act) {
    case PVB_ACT_NONE: break;
    case PVB_ACT_LRELU:
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = y[j] > 0.f ? v[j] : 0.01f * v[j];
      break;
    case PVB_ACT_RELU:
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = y[j] > 0.f ? v[j] : 0.f;
      break;
    case PVB_ACT_TANH:
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] *= 1.f - y[j] * y[j];
      break;
    default:
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] *= pvb::act_grad(y[j], 0.f, act);
  }
}

#ifdef PVB_TC_TRACE
__device__ long long g_ctrace[2][64];
#define CTRACE(role, ev) do { if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 4) && (ev) < 64) g_ctrace[role][ev] = clock64(); } while (0)
__device__ long long g_wtrace[2][64];
#define WTRACE(role, ev) do { if (blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && (warp == 0 || warp == mma_warp) && (ev) < 64) g_wtrace[role][ev] = clock64(); } while (0)
#define PTRACE(role, ev) do { if (blockIdx.x == 300 && lane == 0 && (warp == 0 || warp == 8) && (ev) < 64) g_wtrace[role][ev] = clock64(); } while (0)
__device__ long long g_p3trace[3][64];
#define P3TRACE(role, ev) do { if (blockIdx.x == 17 && lane == 0 && (ev) < 64) g_p3trace[role][ev] = clock64(); } while (0)
#else
#define CTRACE(role, ev) do {} while (0)
#define WTRACE(role, ev) do {} while (0)
#define PTRACE(role, ev) do {} while (0)
#define P3TRACE(role, ev) do {} while (0)
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
      if (umma::elect_one()) {
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
        if (SCALED) {
          // backward data: un-scale; with `pre` = the activations this gradient flows into (output of
          // the previous layer) also apply that layer's activation derivative, so dst is its dpre
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] *= d.out_scale;
          if (pbase) {
            float yv[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) yv[j] = __ldg(pbase + (int64_t)(n0 + j) * HW);
            act_grad16(v, yv, act);
          }
        } else {
          bias_act16(v, bias, n0, d.out_scale, act, pbase, HW);
        }
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

// ---- pixel GEMM with tap reuse ----------------------------------------------------------------------
// Same GEMM as conv_tc_pix_kernel, but the input is gathered ONCE per 64-channel chunk instead of once
// per tap.  Pixels are addressed in a padded position space, q = (b*Hs + h)*Wp + wp with Wp = W + kw/2
// and Hs = H + kh/2: one zero column per row and one zero row per image, shared by the borders on
// either side, so the neighbour (dh, dw) of ANY position is q + dh*Wp + dw.  A tile is 128 consecutive
// positions (the padding positions compute garbage that is not stored: W/(W+1) * H/(H+1) of the MMA
// rows are useful); the CTA stages rows q0 - halo .. q0 + 127 + halo (halo = kh/2 * Wp + kw/2) of a
// channel chunk in the row-chunk layout, and the A operand of tap (dh, dw) is that buffer entered
// halo + dh*Wp + dw rows further down: a 16-byte step per row in the shared-memory descriptor.
// 8 producer / epilogue warps + 1 MMA warp; weights of (tap, chunk) stream through a 2-slot ring.
constexpr int P2_PROD = 256;                     // producer threads
constexpr int P2_THREADS = P2_PROD + 32;
constexpr int P2_MAX_ROWS = 288;                 // staged rows per chunk (<= 128 + 2 * halo)

struct P2Dims {
  int B, Cg, Nout, H, W, kh, kw, sign;
  float in_scale, out_scale;
  int n_rows, x_stages;                          // staged rows per chunk; 1 or 2 chunk buffers
  int n_total, n_off;                            // persistent kernel: this launch computes output channels
                                                 // [n_off, n_off + Nout) of n_total (weights of a 128-wide
                                                 // layer fit shared memory one half at a time)
};

template <bool BF16, bool SCALED, int NCH>
__device__ __forceinline__ void gather_cs(const float* __restrict__ p, int hw, bool ok, float scale,
                                          uint8_t* dst, int cs) {
  float v[NCH];
#pragma unroll
  for (int j = 0; j < NCH; ++j) v[j] = __ldg(p + j * hw);
  const float f = ok ? scale : 0.f;
#pragma unroll
  for (int c8 = 0; c8 < NCH / 8; ++c8) {
    float* w = v + c8 * 8;
    if (!SCALED) {
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] *= f;
    }
    *reinterpret_cast<uint4*>(dst + c8 * cs) =
        SCALED ? make_uint4(pack2s<BF16>(w[0], w[1], f), pack2s<BF16>(w[2], w[3], f),
                            pack2s<BF16>(w[4], w[5], f), pack2s<BF16>(w[6], w[7], f))
               : make_uint4(pack2<BF16>(w[0], w[1]), pack2<BF16>(w[2], w[3]), pack2<BF16>(w[4], w[5]),
                            pack2<BF16>(w[6], w[7]));
  }
}

template <bool BF16, bool SCALED>
__global__ void __launch_bounds__(P2_THREADS, 2)
conv_tc_pix2_kernel(const float* __restrict__ src, const uint16_t* __restrict__ Wp_,
                    const float* __restrict__ bias, float* __restrict__ dst, float* __restrict__ pre,
                    P2Dims d, int act) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int XCS = d.n_rows * ROWB;                 // chunk-column stride of the staged input
  const int x_stage = 8 * XCS;
  uint8_t* sX = smem;
  uint8_t* sB = smem + d.x_stages * x_stage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + 2 * B_STAGE);
  uint64_t* x_full = bars;          // [2] count 8 (producer warps)
  uint64_t* x_empty = bars + 2;     // [2] tcgen05.commit
  uint64_t* b_full = bars + 4;      // [2] count 8
  uint64_t* b_empty = bars + 6;     // [2] tcgen05.commit
  uint64_t* accb = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int taps = d.kh * d.kw, ph = d.kh / 2, pw = d.kw / 2;
  const int Wp = d.W + pw, Hs = d.H + ph;
  const int halo = ph * Wp + pw;
  const int HW = d.H * d.W;
  const int cchunks = (d.Cg + CC - 1) / CC;
  const int Nout = d.Nout;
  const int q0 = (int)blockIdx.x * TP;
  if (warp == 8) umma::tmem_alloc<128>(tmem_slot);
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      umma::mbar_init(x_full + s, 8);
      umma::mbar_init(x_empty + s, 1);
      umma::mbar_init(b_full + s, 8);
      umma::mbar_init(b_empty + s, 1);
    }
    umma::mbar_init(accb, 1);
    umma::mbar_fence_init();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = *tmem_slot;
  PTRACE(warp == 8 ? 1 : 0, 0);

  if (warp == 8) {
    // ================= MMA issuer =================
    const uint32_t idesc = idesc_16b(128, Nout, 0, 0, BF16);
    const uint32_t x0 = umma::smem_u32(sX), b0 = umma::smem_u32(sB);
    int i = 0;
    for (int c = 0; c < cchunks; ++c) {
      const int sx = c % d.x_stages;
      umma::mbar_wait(x_full + sx, (uint32_t)((c / d.x_stages) & 1));
      PTRACE(1, 1);
      const int cc = min(CC, d.Cg - c * CC);
      for (int t = 0; t < taps; ++t, ++i) {
        const int sb = i & 1;
        umma::mbar_wait(b_full + sb, (uint32_t)((i >> 1) & 1));
        PTRACE(1, 2 + 2 * i);
        umma::fence_after_sync();
        if (umma::elect_one()) {
          const int shift = halo + d.sign * ((t / d.kw - ph) * Wp + (t % d.kw - pw));
          for (int k16 = 0; k16 < cc / 16; ++k16) {
            uint64_t da = umma::smem_desc(x0 + sx * x_stage + shift * ROWB + k16 * 2 * XCS, XCS, 128);
            uint64_t db = umma::smem_desc(b0 + sb * B_STAGE + k16 * 2 * (Nout * ROWB), Nout * ROWB, 128);
            umma::mma_f16_ss(tm, da, db, idesc, (i > 0 || k16 > 0) ? 1u : 0u);
          }
          umma::commit(b_empty + sb);
          if (t == taps - 1) umma::commit(x_empty + sx);
          if (c == cchunks - 1 && t == taps - 1) umma::commit(accb);
        }
        __syncwarp();
        PTRACE(1, 3 + 2 * i);
      }
    }
  } else {
    // ================= producers (then epilogue) =================
    int i = 0;
    for (int c = 0; c < cchunks; ++c) {
      const int sx = c % d.x_stages;
      if (c >= d.x_stages) umma::mbar_wait(x_empty + sx, (uint32_t)(((c / d.x_stages) - 1) & 1));
      const int c0 = c * CC, cc = min(CC, d.Cg - c0);
      // ---- stage the input rows of this chunk: row r <-> position q0 - halo + r ----
      for (int r = tid; r < d.n_rows; r += P2_PROD) {
        const int q = q0 - halo + r;
        bool ok = q >= 0;
        const unsigned ri = (unsigned)(ok ? q : 0) / (unsigned)Wp;
        const int wp = (ok ? q : 0) - (int)ri * Wp;
        const int b = (int)(ri / (unsigned)Hs), h = (int)ri - b * Hs;
        ok = ok && wp < d.W && h < d.H && b < d.B;
        const float* p = src + (ok ? ((int64_t)b * d.Cg + c0) * HW + h * d.W + wp : (int64_t)c0 * HW);
        uint8_t* xd = sX + sx * x_stage + r * ROWB;
        switch (cc) {   // uniform
          case 64: gather_cs<BF16, SCALED, 64>(p, HW, ok, d.in_scale, xd, XCS); break;
          case 48: gather_cs<BF16, SCALED, 48>(p, HW, ok, d.in_scale, xd, XCS); break;
          case 32: gather_cs<BF16, SCALED, 32>(p, HW, ok, d.in_scale, xd, XCS); break;
          case 16: gather_cs<BF16, SCALED, 16>(p, HW, ok, d.in_scale, xd, XCS); break;
          default: break;
        }
      }
      PTRACE(0, 1);
      umma::fence_proxy_async();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(x_full + sx);
      PTRACE(0, 2);
      // ---- weights of (tap, chunk): Wp[tap][n][c0 .. c0 + cc): thread -> row n, half of the chunk ----
      const int n = tid & 127, half = tid >> 7;
      for (int t = 0; t < taps; ++t, ++i) {
        const int sb = i & 1;
        if (i >= 2) umma::mbar_wait(b_empty + sb, (uint32_t)(((i >> 1) - 1) & 1));
        if (n < Nout) {
          const uint16_t* wrow = Wp_ + ((int64_t)t * Nout + n) * d.Cg + c0 + half * 32;
          uint8_t* bd = sB + sb * B_STAGE + n * ROWB + half * 4 * (Nout * ROWB);
          uint4 wq[4];
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8)
            if (half * 32 + c8 * 8 < cc) wq[c8] = __ldg(reinterpret_cast<const uint4*>(wrow + c8 * 8));
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8)
            if (half * 32 + c8 * 8 < cc) *reinterpret_cast<uint4*>(bd + c8 * (Nout * ROWB)) = wq[c8];
        }
        umma::fence_proxy_async();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(b_full + sb);
        PTRACE(0, 3 + i);
      }
    }
    // ---- epilogue: TMEM -> bias + activation -> NCHW fp32; warps 0-3 / 4-7 split the columns ----
    umma::mbar_wait(accb, 0);
    PTRACE(0, 40);
    umma::fence_after_sync();
    const int q = q0 + (warp & 3) * 32 + lane;
    const unsigned ri = (unsigned)q / (unsigned)Wp;
    const int wp = q - (int)ri * Wp;
    const int b = (int)(ri / (unsigned)Hs), h = (int)ri - b * Hs;
    const bool m_ok = wp < d.W && h < d.H && b < d.B;
    const int gr = h * d.W + wp;
    const uint32_t tm_lane = tm + ((uint32_t)((warp & 3) * 32) << 16);
    float* obase = dst + (int64_t)b * Nout * HW + gr;
    float* pbase = pre ? pre + (int64_t)b * Nout * HW + gr : nullptr;
    const int n_half = ((Nout / 16 + 1) / 2) * 16;          // columns of warps 0-3
    const int n_lo = warp < 4 ? 0 : n_half, n_hi = warp < 4 ? n_half : Nout;
    for (int n0 = n_lo; n0 < n_hi; n0 += 16) {
      float v[16];
      umma::tmem_ld16(tm_lane + n0, v);
      umma::tmem_ld_wait();
      if (m_ok) {
        if (SCALED) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] *= d.out_scale;
          if (pbase) {
            float yv[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) yv[j] = __ldg(pbase + (int64_t)(n0 + j) * HW);
            act_grad16(v, yv, act);
          }
        } else {
          bias_act16(v, bias, n0, d.out_scale, act, pbase, HW);
        }
        float* o = obase + (int64_t)n0 * HW;
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j * HW] = v[j];
      }
    }
    umma::fence_before_sync();
    PTRACE(0, 41);
  }
  __syncthreads();
  if (warp == 8) umma::tmem_dealloc<128>(tm);
}

// ---- persistent pixel GEMM: weights resident, tiles pipelined -----------------------------------------
// Same position space and tap addressing as conv_tc_pix2_kernel, for layers whose repacked weights
// (taps x Cg x Nout fp16) fit in shared memory next to two staged input tiles.  One CTA per SM loops
// over its tiles with three roles running concurrently on different tiles:
//   warps 4-11  producers : gather the input rows of tile i + 1 (all channel chunks)
//   warp  12    MMA       : taps x chunks x K-steps MMAs of tile i back to back (no per-tap hand-shake:
//                           the weights never move), accumulator buffer i % 2 in TMEM
//   warps 0-3   epilogue  : TMEM -> bias / activation (or activation derivative) -> NCHW of tile i - 1
constexpr int P3_THREADS = 13 * 32;

// MMAs of one tile, issued by one thread: kernel shape compile-time so the tap / K-step loops unroll and
// every descriptor is one 64-bit add away from a register-resident base (the issuing thread is a single
// dependent instruction stream: its instruction count per MMA bounds the tile rate)
template <int KH, int KW>
__device__ __forceinline__ void issue_tile(uint32_t tacc, uint64_t xa0, uint64_t db_base, const uint32_t (&toff)[9],
                                           int cchunks, int Cg, int x_chunk, int w_tile, uint64_t da_kstep,
                                           uint64_t db_kstep, uint32_t idesc) {
  uint32_t first = 0u;
  for (int c = 0; c < cchunks; ++c) {
    const int ksteps = min(CC, Cg - c * CC) / 16;
    const uint64_t xc = xa0 + (uint64_t)((c * x_chunk) >> 4);
    const uint64_t wc = db_base + (uint64_t)((c * w_tile) >> 4);
    const uint64_t w_tap = (uint64_t)((cchunks * w_tile) >> 4);
#pragma unroll
    for (int t = 0; t < KH * KW; ++t) {
      const uint64_t da = xc + (uint64_t)toff[t];
      const uint64_t db = wc + (uint64_t)t * w_tap;
#pragma unroll
      for (int k16 = 0; k16 < 4; ++k16)
        if (k16 < ksteps) {
          umma::mma_f16_ss(tacc, da + (uint64_t)k16 * da_kstep, db + (uint64_t)k16 * db_kstep, idesc, first);
          first = 1u;
        }
    }
  }
}

template <bool BF16, bool SCALED>
__global__ void __launch_bounds__(P3_THREADS, 1)
conv_tc_pix3_kernel(const float* __restrict__ src, const uint16_t* __restrict__ Wp_,
                    const float* __restrict__ bias, float* __restrict__ dst, float* __restrict__ pre,
                    P2Dims d, int act, int n_tiles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int taps = d.kh * d.kw, ph = d.kh / 2, pw = d.kw / 2;
  const int cchunks = (d.Cg + CC - 1) / CC;
  const int Nout = d.Nout;
  const int XCS = d.n_rows * ROWB;                 // chunk-column stride of a staged input chunk
  const int x_chunk = 8 * XCS, x_stage = cchunks * x_chunk;
  const int w_tile = Nout * CC * 2;                // weights of one (tap, chunk): [Nout][64] fp16
  uint8_t* sW = smem;
  uint8_t* sX = smem + taps * cchunks * w_tile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sX + 2 * x_stage);
  uint64_t* x_full = bars;          // [2] count 8 (producer warps)
  uint64_t* x_empty = bars + 2;     // [2] tcgen05.commit
  uint64_t* a_full = bars + 4;      // [2] tcgen05.commit: accumulator of a tile complete
  uint64_t* a_empty = bars + 6;     // [2] count 4 (epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Wp = d.W + pw, Hs = d.H + ph;
  const int halo = ph * Wp + pw;
  const int HW = d.H * d.W;
  const int my_tiles = (int)blockIdx.x < n_tiles ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  if (warp == 12) umma::tmem_alloc<256>(tmem_slot);
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      umma::mbar_init(x_full + s, 8);
      umma::mbar_init(x_empty + s, 1);
      umma::mbar_init(a_full + s, 1);
      umma::mbar_init(a_empty + s, 4);
    }
    umma::mbar_fence_init();
  }
  // resident weights: Wp_[tap][n][k] -> per (tap, chunk) a K-major row-chunk tile [Nout][64]
  {
    const int k8s = d.Cg / 8;                      // 16-byte pieces per weight row
    const int total = taps * Nout * k8s;
    for (int i = tid; i < total; i += P3_THREADS) {
      const int k8 = i % k8s, n = (i / k8s) % Nout, t = i / (k8s * Nout);
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(Wp_ + ((int64_t)t * d.n_total + d.n_off + n) * d.Cg) + k8);
      const int c = k8 / 8, c8 = k8 % 8;
      *reinterpret_cast<uint4*>(sW + (t * cchunks + c) * w_tile + c8 * (Nout * ROWB) + n * ROWB) = v;
    }
    umma::fence_proxy_async();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = *tmem_slot;

  if (warp == 12) {
    // ================= MMA issuer =================
    const uint32_t idesc = idesc_16b(128, Nout, 0, 0, BF16);
    const uint64_t da_base = umma::smem_desc(umma::smem_u32(sX), XCS, 128);
    const uint64_t db_base = umma::smem_desc(umma::smem_u32(sW), Nout * ROWB, 128);
    const uint64_t da_kstep = (uint64_t)((2 * XCS) >> 4), db_kstep = (uint64_t)((2 * Nout * ROWB) >> 4);
    // row offset of every tap inside the staged tile (16-byte units), kept in registers
    uint32_t toff[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int th = t < taps ? t / d.kw : 0, tw = t < taps ? t % d.kw : 0;
      toff[t] = (uint32_t)(halo + d.sign * ((th - ph) * Wp + (tw - pw)));
    }
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it & 1;
      P3TRACE(1, 3 * it);
      umma::mbar_wait(x_full + s, (uint32_t)((it >> 1) & 1));
      if (it >= 2) umma::mbar_wait(a_empty + s, (uint32_t)(((it >> 1) - 1) & 1));
      P3TRACE(1, 3 * it + 1);
      umma::fence_after_sync();
      if (umma::elect_one()) {
        const uint64_t xa0 = da_base + (uint64_t)((s * x_stage) >> 4);
        const uint32_t tacc = tm + s * 128;
        if (d.kh == 3) issue_tile<3, 3>(tacc, xa0, db_base, toff, cchunks, d.Cg, x_chunk, w_tile, da_kstep, db_kstep, idesc);
        else if (d.kw == 3) issue_tile<1, 3>(tacc, xa0, db_base, toff, cchunks, d.Cg, x_chunk, w_tile, da_kstep, db_kstep, idesc);
        else issue_tile<1, 1>(tacc, xa0, db_base, toff, cchunks, d.Cg, x_chunk, w_tile, da_kstep, db_kstep, idesc);
        umma::commit(x_empty + s);
        umma::commit(a_full + s);
      }
      __syncwarp();
      P3TRACE(1, 3 * it + 2);
    }
  } else if (warp >= 4) {
    // ================= producers =================
    const int ptid = tid - 128;
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it & 1;
      if (warp == 4) P3TRACE(0, 3 * it);
      if (it >= 2) umma::mbar_wait(x_empty + s, (uint32_t)(((it >> 1) - 1) & 1));
      if (warp == 4) P3TRACE(0, 3 * it + 1);
      const int q0 = ((int)blockIdx.x + it * (int)gridDim.x) * TP;
      for (int r = ptid; r < d.n_rows; r += 256) {
        const int q = q0 - halo + r;
        bool ok = q >= 0;
        const unsigned ri = (unsigned)(ok ? q : 0) / (unsigned)Wp;
        const int wp = (ok ? q : 0) - (int)ri * Wp;
        const int b = (int)(ri / (unsigned)Hs), h = (int)ri - b * Hs;
        ok = ok && wp < d.W && h < d.H && b < d.B;
        const float* p = src + (ok ? (int64_t)b * d.Cg * HW + h * d.W + wp : 0);
        uint8_t* xd = sX + s * x_stage + r * ROWB;
        for (int c = 0; c < cchunks; ++c) {
          const int cc = min(CC, d.Cg - c * CC);
          const float* pc = p + (int64_t)c * CC * HW;
          switch (cc) {   // uniform
            case 64: gather_cs<BF16, SCALED, 64>(pc, HW, ok, d.in_scale, xd + c * x_chunk, XCS); break;
            case 48: gather_cs<BF16, SCALED, 48>(pc, HW, ok, d.in_scale, xd + c * x_chunk, XCS); break;
            case 32: gather_cs<BF16, SCALED, 32>(pc, HW, ok, d.in_scale, xd + c * x_chunk, XCS); break;
            case 16: gather_cs<BF16, SCALED, 16>(pc, HW, ok, d.in_scale, xd + c * x_chunk, XCS); break;
            default: break;
          }
        }
      }
      umma::fence_proxy_async();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(x_full + s);
      if (warp == 4) P3TRACE(0, 3 * it + 2);
    }
  } else {
    // ================= epilogue =================
    for (int it = 0; it < my_tiles; ++it) {
      const int s = it & 1;
      if (warp == 0) P3TRACE(2, 3 * it);
      umma::mbar_wait(a_full + s, (uint32_t)((it >> 1) & 1));
      if (warp == 0) P3TRACE(2, 3 * it + 1);
      umma::fence_after_sync();
      const int q = ((int)blockIdx.x + it * (int)gridDim.x) * TP + warp * 32 + lane;
      const unsigned ri = (unsigned)q / (unsigned)Wp;
      const int wp = q - (int)ri * Wp;
      const int b = (int)(ri / (unsigned)Hs), h = (int)ri - b * Hs;
      const bool m_ok = wp < d.W && h < d.H && b < d.B;
      const int gr = h * d.W + wp;
      const uint32_t tm_lane = tm + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * 128);
      const int64_t ooff = ((int64_t)b * d.n_total + d.n_off) * HW + gr;
      float* obase = dst + ooff;
      float* pbase = pre ? pre + ooff : nullptr;
      for (int n0 = 0; n0 < Nout; n0 += 16) {
        float v[16];
        umma::tmem_ld16(tm_lane + n0, v);
        umma::tmem_ld_wait();
        if (m_ok) {
          if (SCALED) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] *= d.out_scale;
            if (pbase) {
              float yv[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) yv[j] = __ldg(pbase + (int64_t)(n0 + j) * HW);
              act_grad16(v, yv, act);
            }
          } else {
            bias_act16(v, bias, n0, d.out_scale, act, pbase, HW);
          }
          float* o = obase + (int64_t)n0 * HW;
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j * HW] = v[j];
        }
      }
      umma::fence_before_sync();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(a_empty + s);
      if (warp == 0) P3TRACE(2, 3 * it + 2);
    }
  }
  __syncthreads();
  if (warp == 12) umma::tmem_dealloc<256>(tm);
}

// floor(n / d) for n < 2^31 by one widening multiply and a shift: mul = floor(2^p / d) + 1, p = 31 + ceil(log2 d)
// (the error term n * (mul * d - 2^p) <= n * 2^(p - 31) stays below 2^p).  Built once per thread; the hardware
// has no integer divider and the two divisions of the per-tile geometry were ~50 instructions per thread.
struct FastDiv {
  uint32_t mul, p;
  __device__ __forceinline__ explicit FastDiv(uint32_t d) {
    const uint32_t sh = d > 1 ? 32u - (uint32_t)__clz(d - 1) : 0u;
    p = 31u + sh;
    mul = (uint32_t)((1ull << p) / d) + 1u;
  }
  __device__ __forceinline__ uint32_t div(uint32_t n) const { return (uint32_t)(((uint64_t)n * mul) >> p); }
};

// ---- backward weight ------------------------------------------------------------------------------
// dW[co][ci][dh][dw] = sum over pixels of dpre[co][h][w] x[ci][h+dh][w+dw] as GEMMs with K = pixels:
// accumulators [128 lanes = co][kw x Cin columns] in TMEM, both operands MN-major (K = tile rows,
// exactly as the decoder kernel's dW GEMMs).
//
// Pixels are addressed through W-padded rows, q = (b*H + h)*Wp + wp with Wp = W + 1 for 3-wide kernels:
// the extra column is the zero padding shared by the right border of one row and the left border of
// the next, so "one pixel to the left / right" is q -/+ 1 everywhere.  CTA (g, y) owns kernel row
// dh = g - kh/2 and the tiles y, y + gridDim.y, ...; per tile (ADV = 128 - 2*(kw/2) positions):
//   A = dpre tile [128 rows: position q0 + r; rows >= ADV and padding positions are zero][128 co]
//   X = x tile    [128 rows: position q0 - kw/2 + r of image row h + dh][Cin], staged ONCE;
//   the B operand of tap (dh, dw) is X shifted by kw/2 + dw rows = +16 bytes per row in the smem
//   descriptor (row-chunk layout), so the kw taps of a kernel row share one gather.
// Gathers: 16-channel items (Cout/16 for A, Cin/16 for X) dealt round-robin to up to four producer
// groups of 128 threads (thread = tile row), software-pipelined one item ahead across tiles.
constexpr int WG_MAX_GROUPS = 4;
constexpr int WCS = TP * ROWB;                   // chunk-column stride: 128 rows x 16 bytes
constexpr int WG_A = 16 * WCS;                   // dpre tile, 16 chunk-columns (128 output channels)
constexpr int WG_ONES = 2 * WCS;                 // [128][16] tile, column 0 = 1: bias sums; sits directly behind
                                                 // the x tile of every stage, so the first tap's MMA spans
                                                 // N = Cin + 16 and no separate bias MMAs are needed
constexpr int WG_XPAD = 32;                      // two finite rows behind the last chunk-column
constexpr int WG_MAX_STAGES = 3;

template <bool BF16>
__global__ void __launch_bounds__(WG_MAX_GROUPS * 128 + 32, 1)
conv_tc_wgrad_kernel(const float* __restrict__ dpre, const float* __restrict__ x, float* __restrict__ dW,
                     float* __restrict__ db, float* __restrict__ scratch, int B, int Cin_real, int Cout, int H,
                     int W, int kh, int kw, int n_stages, int n_tiles) {
  // fewer than 16 input channels (the first layer): the x tile is zero-padded to 16 columns
  const int Cin = Cin_real < 16 ? 16 : Cin_real;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int x_tile = (Cin / 8) * WCS + WG_ONES + WG_XPAD;
  const int stage_bytes = WG_A + x_tile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + n_stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + WG_MAX_STAGES;
  uint64_t* accb = bars + 2 * WG_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_MAX_STAGES + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_groups = (int)(blockDim.x >> 7);             // producer groups; the MMA warp comes last
  const int mma_warp = n_groups * 4;
  const int ph = kh / 2, pw = kw / 2;
  const int dh = (int)blockIdx.x - ph;                     // kernel row of this CTA
  const int Wp = W + pw, ADV = TP - 2 * pw;
  const int HW = H * W;
  const int my_tiles = (int)blockIdx.y < n_tiles ? (n_tiles - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y : 0;
  const bool do_bias = db != nullptr && blockIdx.x == 0;
  // TMEM columns: tap 0 | bias sums (16, only where do_bias) | tap 1 | tap 2
  const uint32_t col_bias = (uint32_t)Cin;
  const uint32_t tap_gap = do_bias ? 16u : 0u;
  if (warp == mma_warp) umma::tmem_alloc<512>(tmem_slot);
  if (tid == 0) {
    for (int s = 0; s < n_stages; ++s) {
      umma::mbar_init(full + s, (uint32_t)mma_warp);       // one arrive per producer warp
      umma::mbar_init(empty + s, 1);
    }
    umma::mbar_init(accb, 1);
    umma::mbar_fence_init();
  }
  if (warp < mma_warp) {
    const int r = tid & 127, g = warp >> 2;
    for (int s = 0; s < n_stages; ++s) {
      // output channels beyond Cout never change: zero those dpre chunk-columns once
      for (int c8 = Cout / 8 + g; c8 < 16; c8 += n_groups)
        *reinterpret_cast<uint4*>(smem + s * stage_bytes + c8 * WCS + r * ROWB) = make_uint4(0u, 0u, 0u, 0u);
      uint8_t* ones = smem + s * stage_bytes + WG_A + (Cin / 8) * WCS;
      if (g == 0) {
        *reinterpret_cast<uint4*>(ones + r * ROWB) = make_uint4(pack2<BF16>(1.f, 0.f), 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(ones + WCS + r * ROWB) = make_uint4(0u, 0u, 0u, 0u);
      }
      if (tid < 2) *reinterpret_cast<uint4*>(ones + WG_ONES + tid * ROWB) = make_uint4(0u, 0u, 0u, 0u);
    }
    umma::fence_proxy_async();
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = *tmem_slot;

  if (warp == mma_warp) {
    const uint32_t id_w = idesc_16b(128, Cin, 1, 1, BF16);
    const uint32_t id_wb = idesc_16b(128, Cin + 16, 1, 1, BF16);   // first tap + the ones columns
    // descriptors are advanced by additions on their start-address field (16-byte units): the issuing
    // thread is one dependent instruction stream, re-encoding a descriptor per MMA costs more than the MMA
    const uint64_t da0 = umma::smem_desc(umma::smem_u32(smem), 128, WCS);              // dpre tile, stage 0
    const uint64_t dx0 = umma::smem_desc(umma::smem_u32(smem) + WG_A, 128, WCS);       // x tile, stage 0
    const uint64_t stage_step = (uint64_t)(stage_bytes >> 4);
    int s = 0;
    uint32_t fpar = 0;
    for (int it = 0; it < my_tiles; ++it, s = (s + 1 == n_stages ? 0 : s + 1), fpar ^= (s == 0 ? 1u : 0u)) {
      WTRACE(1, 3 * it);
      umma::mbar_wait(full + s, fpar);
      WTRACE(1, 3 * it + 1);
      umma::fence_after_sync();
      if (umma::elect_one()) {
        const uint64_t da = da0 + (uint64_t)s * stage_step, dx = dx0 + (uint64_t)s * stage_step;
        const uint32_t acc = it > 0 ? 1u : 0u;
        for (int j = 0; j < kw; ++j) {     // tap (dh, j - pw): X shifted by j rows (16 bytes each)
          const uint32_t tmj = tm + j * Cin + (j > 0 ? tap_gap : 0u);
          const uint32_t idj = (j == 0 && do_bias) ? id_wb : id_w;
#pragma unroll
          for (int k = 0; k < 8; ++k)      // 8 K-steps of 16 tile rows (256 bytes)
            umma::mma_f16_ss(tmj, da + (uint64_t)(k * 16), dx + (uint64_t)(j + k * 16), idj, k > 0 ? 1u : acc);
        }
        umma::commit(empty + s);
        if (it == my_tiles - 1) umma::commit(accb);
      }
      __syncwarp();
      WTRACE(1, 3 * it + 2);
    }
  } else {
    const int row = tid & 127, grp = warp >> 2;
    const int a_items = Cout / 16, b_items = Cin / 16;
    const int n_items = a_items + b_items;
    const int n_mine = (n_items - grp + n_groups - 1) / n_groups;    // >= 1 (host: groups <= items)
    const int total = my_tiles * n_mine;
    const bool small_c = Cin_real < 16;
    // ---- load cursor: position of this thread's tile row, decoded once per tile ----
    int l_tile = (int)blockIdx.y, l_i = 0;
    bool l_aok, l_xok;
    const float *l_dp, *l_xp;
    const FastDiv div_wp((uint32_t)Wp), div_h((uint32_t)H);
    auto tile_geometry = [&]() {
      const bool live = l_tile < n_tiles;     // the cursor runs one tile past the end (nothing is loaded there)
      const unsigned qa = (unsigned)(live ? l_tile : 0) * (unsigned)ADV + (unsigned)row;   // A position (< 2^31: host)
      const unsigned ri = div_wp.div(qa);
      const int wp = (int)(qa - ri * (unsigned)Wp);
      const int b = (int)div_h.div(ri), h = (int)(ri - (unsigned)b * (unsigned)H);
      const bool in_b = live && b < B;
      l_aok = in_b && row < ADV && wp < W;
      const int wx = wp - pw, hx = h + dh;                 // X position: pw to the left, row h + dh
      l_xok = in_b && wx >= 0 && wx < W && hx >= 0 && hx < H;
      // the pointers always address readable memory (pixel 0 of image 0 when the position is void)
      l_dp = dpre + (l_aok ? ((int64_t)b * Cout * H + h) * W + wp : 0);
      l_xp = x + (l_xok ? ((int64_t)b * Cin_real * H + hx) * W + wx : 0);
    };
    tile_geometry();
    struct Pending { float v[16]; uint32_t dst; bool ok, is_x; };
    auto issue = [&](Pending& pd) {
      const int k = grp + l_i * n_groups;
      pd.is_x = k >= a_items;
      const int cb = pd.is_x ? k - a_items : k;            // 16-channel block
      pd.ok = pd.is_x ? l_xok : l_aok;
      pd.dst = (uint32_t)((pd.is_x ? WG_A : 0) + cb * 2 * WCS + row * ROWB);
      const float* p = (pd.is_x ? l_xp : l_dp) + (int64_t)cb * 16 * HW;
      // plane offsets in 32-bit arithmetic (16 planes of one image: far below 2^31): one IMAD.WIDE per
      // load; as 64-bit products the address arithmetic was 7 instructions per load and the kernel's
      // largest cost (ncu: 107 M warp instructions for the 64 -> 64 layer at 32 x 32)
      if (small_c && pd.is_x) {
#pragma unroll
        for (int j = 0; j < 16; ++j) pd.v[j] = j < Cin_real ? __ldg(p + j * HW) : 0.f;
      } else {
        // the item's base pointer is materialised (opaque to the optimiser) so that each load address
        // is ONE IMAD.WIDE of a 32-bit plane offset; left to itself the compiler carried a 64-bit element
        // index and rebuilt base + 4 * (index + j * HW) with five instructions per load
        unsigned long long pb = reinterpret_cast<unsigned long long>(p);
        asm volatile("" : "+l"(pb));
        const float* pq = reinterpret_cast<const float*>(pb);
#pragma unroll
        for (int j = 0; j < 16; ++j) pd.v[j] = __ldg(pq + j * HW);
      }
      if (++l_i == n_mine) {       // next item belongs to the next tile of this CTA
        l_i = 0;
        l_tile += (int)gridDim.y;
        tile_geometry();
      }
    };
    // ---- store cursor ----
    int s_it = 0, s_i = 0, s_stage = 0;          // tile, item within it, stage = s_it % n_stages
    uint32_t s_par = 1;                           // parity of the `empty` phase that frees the stage: ((s_it / n_stages) - 1) & 1
    int tr_ev = 0;      // trace event counter (debug builds only)
    (void)tr_ev;
    auto finish = [&](Pending& pd) {
      const int s = s_stage;
      uint8_t* st = smem + s * stage_bytes;
      WTRACE(0, tr_ev++);
      if (s_i == 0 && s_it >= n_stages) umma::mbar_wait(empty + s, s_par);   // stage free again
      WTRACE(0, tr_ev++);
      uint32_t pk[8];
      if (pd.is_x) {
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = pack2<BF16>(pd.v[2 * j], pd.v[2 * j + 1]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = pack2s<BF16>(pd.v[2 * j], pd.v[2 * j + 1], GRAD_SCALE);
      }
      if (!pd.ok) {
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = 0u;
      }
      *reinterpret_cast<uint4*>(st + pd.dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(st + pd.dst + WCS) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      WTRACE(0, tr_ev++);
      if (++s_i == n_mine) {
        umma::fence_proxy_async();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(full + s);
        s_i = 0;
        ++s_it;
        if (++s_stage == n_stages) {
          s_stage = 0;
          s_par ^= 1u;
        }
      }
    };
    Pending pa, pb;
    if (total > 0) issue(pa);
    for (int q = 0; q < total; q += 2) {
      if (q + 1 < total) issue(pb);
      finish(pa);
      if (q + 2 < total) issue(pa);
      if (q + 1 < total) finish(pb);
    }
    if (my_tiles > 0 && scratch != nullptr) {
      // Accumulator read-out, coalesced: a TMEM lane is an output channel, so the 32 lanes of a warp add 32
      // CONSECUTIVE floats of the transposed scratch gradient S[tap][ci][co] (one 128-byte reduction per
      // instruction); wgrad_finish_kernel folds S into dW[co][ci][tap].  Every producer warp takes part: warp w
      // reads lane quarter w % 4, the column blocks are dealt to the n_groups warps of a quarter.  (Adding into
      // dW directly put the lanes Cin * taps floats apart -- 32 sectors per instruction, issued by four warps:
      // ~25 us per launch at 128 channels, most of the time of the small 1-D layers.)
      umma::mbar_wait(accb, 0);
      umma::fence_after_sync();
      const int quarter = warp & 3;
      const int co = quarter * 32 + lane;
      if (quarter * 32 < Cout) {
        const uint32_t tm_lane = tm + ((uint32_t)(quarter * 32) << 16);
        const int cblk = Cin / 16, nblk = kw * cblk;
        for (int blk = grp; blk < nblk; blk += n_groups) {
          const int j = blk / cblk, n0 = (blk - j * cblk) * 16;
          const int tap = (int)blockIdx.x * kw + j;
          float v[16];
          umma::tmem_ld16(tm_lane + j * Cin + (j > 0 ? tap_gap : 0u) + n0, v);
          umma::tmem_ld_wait();
          if (co < Cout) {
            float* sp = scratch + ((int64_t)tap * Cin_real + n0) * Cout + co;
#pragma unroll
            for (int jj = 0; jj < 16; ++jj)
              if (n0 + jj < Cin_real) atomicAdd(sp + jj * Cout, v[jj]);
          }
        }
        if (do_bias && grp == 0) {
          float v[16];
          umma::tmem_ld16(tm_lane + col_bias, v);
          umma::tmem_ld_wait();
          if (co < Cout) atomicAdd(scratch + (int64_t)kh * kw * Cin_real * Cout + co, v[0]);
        }
      }
      umma::fence_before_sync();
    } else if (my_tiles > 0 && grp == 0) {
      umma::mbar_wait(accb, 0);
      umma::fence_after_sync();
      const uint32_t tm_lane = tm + ((uint32_t)(warp * 32) << 16);
      const int co = row;                                   // TMEM lane = output channel
      const int taps = kh * kw;
      for (int j = 0; j < kw; ++j) {
        const int tap = (int)blockIdx.x * kw + j;
        for (int n0 = 0; n0 < Cin; n0 += 16) {
          float v[16];
          umma::tmem_ld16(tm_lane + j * Cin + (j > 0 ? tap_gap : 0u) + n0, v);
          umma::tmem_ld_wait();
          if (co < Cout) {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj)
              if (n0 + jj < Cin_real)
                atomicAdd(dW + ((int64_t)co * Cin_real + n0 + jj) * taps + tap, v[jj] * (1.f / GRAD_SCALE));
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

// dW[co][ci][tap] += S[tap][ci][co] / GRAD_SCALE, db[co] += S_b[co] / GRAD_SCALE, and S is left zero for the
// next call.  A CTA transposes one (32 co x 32 ci x all taps) block through shared memory, so that both the
// scratch rows (co contiguous) and the weight rows (ci, tap contiguous) move in full lines -- element-wise,
// one side is always 4 bytes per 32-byte sector and the fold of a whole backward pass cost ~50 us.
// blockIdx.y = layer: one launch folds every layer of a backward pass.
struct FoldArgs {
  pvb_wgrad_fold p[PVB_WGRAD_FOLD_MAX];
};
constexpr int FOLD_MAX_TAPS = 9;
__global__ void __launch_bounds__(256) wgrad_finish_kernel(FoldArgs a) {
  __shared__ float tile[FOLD_MAX_TAPS][32][33];
  const pvb_wgrad_fold pr = a.p[blockIdx.y];
  const int cob = (pr.Cout + 31) / 32, cib = (pr.Cin + 31) / 32;
  if ((int)blockIdx.x >= cob * cib) return;
  const int co0 = ((int)blockIdx.x % cob) * 32, ci0 = ((int)blockIdx.x / cob) * 32;
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  // scratch -> tile (and clear): row (t, ci) of 32 consecutive co per warp.  All of a warp's loads are issued
  // before the first use (as a load / store / use loop each row paid a full memory latency: 28 us per launch)
  constexpr int RPW = FOLD_MAX_TAPS * 32 / 8;      // rows per warp, upper bound
  float vals[RPW];
#pragma unroll
  for (int k = 0; k < RPW; ++k) {
    const int r = wy + 8 * k;
    const int t = r >> 5, ci = ci0 + (r & 31), co = co0 + lane;
    vals[k] = 0.f;
    if (t < pr.taps && ci < pr.Cin && co < pr.Cout)
      vals[k] = __ldcg(pr.scratch + ((int64_t)t * pr.Cin + ci) * pr.Cout + co);
  }
#pragma unroll
  for (int k = 0; k < RPW; ++k) {
    const int r = wy + 8 * k;
    const int t = r >> 5, ci = ci0 + (r & 31), co = co0 + lane;
    if (t < pr.taps) {
      if (ci < pr.Cin && co < pr.Cout) pr.scratch[((int64_t)t * pr.Cin + ci) * pr.Cout + co] = 0.f;
      tile[t][r & 31][lane] = vals[k];
    }
  }
  __syncthreads();
  // tile -> dW: for one co, the (ci, tap) block is nci * taps consecutive floats; again loads first
  const int nci = min(32, pr.Cin - ci0), run = nci * pr.taps;
#pragma unroll
  for (int cc = 0; cc < 4; ++cc) {
    const int c = wy + 8 * cc, co = co0 + c;
    if (co >= pr.Cout) break;
    float* wp = pr.dW + ((int64_t)co * pr.Cin + ci0) * pr.taps;
    float old[FOLD_MAX_TAPS];
#pragma unroll
    for (int u = 0; u < FOLD_MAX_TAPS; ++u) {
      const int e = lane + 32 * u;
      old[u] = e < run ? wp[e] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < FOLD_MAX_TAPS; ++u) {
      const int e = lane + 32 * u;
      if (e < run) {
        const int ci = e / pr.taps, t = e - ci * pr.taps;
        wp[e] = old[u] + tile[t][ci][c] * (1.f / GRAD_SCALE);
      }
    }
  }
  if (blockIdx.x == 0) {
    float* sb = pr.scratch + (int64_t)pr.taps * pr.Cin * pr.Cout;
    for (int co = threadIdx.x; co < pr.Cout; co += 256) {
      const float v = sb[co];
      sb[co] = 0.f;
      if (pr.db) pr.db[co] += v * (1.f / GRAD_SCALE);
    }
  }
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
  if (Cin == 1) return 0;      // HBM-bound: the direct fp32 kernel of pvb_conv_bwd_weight is faster
  const int C = Cin < 16 ? 16 : Cin;
  return tc_ok(C, Cout, kh, kw) && Cout <= 128 && kw * C + 16 <= 512 && C + 16 <= 256 ? 1 : 0;
}

extern "C" int64_t pvb_conv_tc_workspace_bytes(int Cin, int Cout, int kh, int kw) {
  return (int64_t)Cin * Cout * kh * kw * 2;
}

// mode 0: y = act(conv(x, W) + b)   (fp16 operands)     src = x   [B, Cin, H, W]
// mode 1: dx = conv_transpose(dpre, W)  (bf16 operands)  src = dpre [B, Cout, H, W]
extern "C" int pvb_conv_tc_prep(const float* W, void* workspace, int Cin, int Cout, int kh, int kw, int mode,
                                void* stream) {
  PVB_CHECK_ARG(W && workspace && (mode == 0 || mode == 1), "pvb_conv_tc_prep: bad argument");
  PVB_CHECK_ARG(((uintptr_t)workspace % 16) == 0, "pvb_conv_tc_prep: workspace must be 16-byte aligned");
  const int taps = kh * kw;
  const int64_t total = (int64_t)Cin * Cout * taps;
  uint16_t* Wp = reinterpret_cast<uint16_t*>(workspace);
  if (mode == 0)
    conv_tc_prep_kernel<false><<<pvb::cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(W, Wp, Cout, Cin, taps, 0);
  else
    conv_tc_prep_kernel<BWD_BF16><<<pvb::cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(W, Wp, Cout, Cin, taps, 1);
  pvb::count_launch();
  return pvb::launch_status();
}

extern "C" int pvb_conv_tc_pix(const float* src, const float* W, const float* b, float* dst, float* pre,
                               void* workspace, int B, int Cin, int Cout, int H, int Wd, int kh, int kw,
                               int act, int mode_flags, void* stream) {
  PVB_CHECK_ARG(src && W && dst && workspace, "pvb_conv_tc_pix: null pointer");
  // bit 1: the workspace already holds this mode's repacked weights (pvb_conv_tc_prep, once per step)
  const bool prepped = (mode_flags & 2) != 0;
  const int mode = mode_flags & ~2;
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
  // tap-reuse kernel whenever the staged rows (128 + 2 * halo) fit its buffers
  const int ph2 = kh / 2, pw2 = kw / 2;
  const int n_rows = TP + 2 * (ph2 * (Wd + pw2) + pw2);
  const int64_t positions = (int64_t)B * (H + ph2) * (Wd + pw2);
  static const bool use_pix2 = getenv("PVB_CONV_PIX1") == nullptr;
  if (use_pix2 && n_rows <= P2_MAX_ROWS && positions + TP < (1ll << 31)) {
    const int cch = (Cg + CC - 1) / CC;
    P2Dims p2{B, Cg, Nout, H, Wd, kh, kw, mode == 0 ? 1 : -1,
              mode == 0 ? 1.f : GRAD_SCALE, mode == 0 ? 1.f : 1.f / GRAD_SCALE, n_rows, cch > 1 ? 2 : 1,
              Nout, 0};
    const int smem2 = p2.x_stages * 8 * n_rows * ROWB + 2 * B_STAGE + 128;
    static bool attr2 = false;
    if (!attr2) {
      cudaFuncSetAttribute(conv_tc_pix2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
      cudaFuncSetAttribute(conv_tc_pix2_kernel<BWD_BF16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
      attr2 = true;
    }
    const unsigned grid2 = (unsigned)((positions + TP - 1) / TP);
    // persistent variant when all repacked weights + two staged input tiles fit one CTA's shared memory
    // (a 128-wide layer whose weights are too large runs as two launches of 64 output channels each)
    int n_launch = Nout;
    int64_t smem3 = (int64_t)taps * cch * n_launch * CC * 2 + 2ll * cch * 8 * n_rows * ROWB + 128;
    if (smem3 > 227 * 1024 && Nout == 128) {
      n_launch = 64;
      smem3 = (int64_t)taps * cch * n_launch * CC * 2 + 2ll * cch * 8 * n_rows * ROWB + 128;
    }
    static const bool use_pix3 = getenv("PVB_CONV_PIX2") == nullptr;
    if (use_pix3 && smem3 <= 227 * 1024 && Cg % 8 == 0 && grid2 >= 148) {
      static bool attr3 = false;
      if (!attr3) {
        cudaFuncSetAttribute(conv_tc_pix3_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(conv_tc_pix3_kernel<BWD_BF16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr3 = true;
      }
      if (mode == 0) {
        if (!prepped) conv_tc_prep_kernel<false><<<pvb::cdiv(total, 256), 256, 0, st>>>(W, Wp, Cout, Cin, taps, 0);
      } else {
        PVB_CHECK_ARG(!pre || (act != PVB_ACT_GELU), "pvb_conv_tc_pix: gelu's derivative needs the pre-activation");
        if (!prepped) conv_tc_prep_kernel<BWD_BF16><<<pvb::cdiv(total, 256), 256, 0, st>>>(W, Wp, Cout, Cin, taps, 1);
      }
      if (!prepped) pvb::count_launch();
      for (int n_off = 0; n_off < Nout; n_off += n_launch) {
        P2Dims p3 = p2;
        p3.Nout = n_launch;
        p3.n_total = Nout;
        p3.n_off = n_off;
        if (mode == 0)
          conv_tc_pix3_kernel<false, false><<<148, P3_THREADS, (size_t)smem3, st>>>(
              src, Wp, b ? b + n_off : nullptr, dst, pre, p3, act, (int)grid2);
        else
          conv_tc_pix3_kernel<BWD_BF16, true><<<148, P3_THREADS, (size_t)smem3, st>>>(
              src, Wp, nullptr, dst, pre, p3, pre ? act : 0, (int)grid2);
        pvb::count_launch();
      }
      return pvb::launch_status();
    }
    if (mode == 0) {
      if (!prepped) conv_tc_prep_kernel<false><<<pvb::cdiv(total, 256), 256, 0, st>>>(W, Wp, Cout, Cin, taps, 0);
      if (!prepped) pvb::count_launch();
      conv_tc_pix2_kernel<false, false><<<grid2, P2_THREADS, smem2, st>>>(src, Wp, b, dst, pre, p2, act);
    } else {
      PVB_CHECK_ARG(!pre || (act != PVB_ACT_GELU), "pvb_conv_tc_pix: gelu's derivative needs the pre-activation");
      if (!prepped) conv_tc_prep_kernel<BWD_BF16><<<pvb::cdiv(total, 256), 256, 0, st>>>(W, Wp, Cout, Cin, taps, 1);
      if (!prepped) pvb::count_launch();
      conv_tc_pix2_kernel<BWD_BF16, true><<<grid2, P2_THREADS, smem2, st>>>(src, Wp, nullptr, dst, pre, p2,
                                                                           pre ? act : 0);
    }
    pvb::count_launch();
    return pvb::launch_status();
  }
  TcDims d{B, Cg, Nout, H, Wd, kh, kw, mode == 0 ? 1 : -1,
           mode == 0 ? 1.f : GRAD_SCALE, mode == 0 ? 1.f : 1.f / GRAD_SCALE};
  const int64_t M = (int64_t)B * H * Wd;
  const unsigned grid = (unsigned)((M + TP - 1) / TP);
  if (mode == 0) {
    if (!prepped) conv_tc_prep_kernel<false><<<pvb::cdiv(total, 256), 256, 0, st>>>(W, Wp, Cout, Cin, taps, 0);
    if (!prepped) pvb::count_launch();
    conv_tc_pix_kernel<false, false><<<grid, PIX_THREADS, PIX_SMEM, st>>>(src, Wp, b, dst, pre, d, act);
  } else {
    if (!prepped) conv_tc_prep_kernel<BWD_BF16><<<pvb::cdiv(total, 256), 256, 0, st>>>(W, Wp, Cout, Cin, taps, 1);
    if (!prepped) pvb::count_launch();
    // `pre` (optional) = output of the layer below, `act` its activation: dst = dx * act'(pre)
    PVB_CHECK_ARG(!pre || (act != PVB_ACT_GELU), "pvb_conv_tc_pix: gelu's derivative needs the pre-activation");
    conv_tc_pix_kernel<BWD_BF16, true><<<grid, PIX_THREADS, PIX_SMEM, st>>>(src, Wp, nullptr, dst, pre, d,
                                                                           pre ? act : 0);
  }
  pvb::count_launch();
  return pvb::launch_status();
}

extern "C" int64_t pvb_conv_tc_wgrad_scratch_bytes(int Cin, int Cout, int kh, int kw) {
  return ((int64_t)kh * kw * Cin * Cout + Cout) * (int64_t)sizeof(float);
}

extern "C" int pvb_conv_tc_wgrad_fold(const pvb_wgrad_fold* probs, int n, void* stream) {
  PVB_CHECK_ARG(probs && n >= 0, "pvb_conv_tc_wgrad_fold: bad argument");
  for (int i0 = 0; i0 < n; i0 += PVB_WGRAD_FOLD_MAX) {
    FoldArgs a;
    const int cnt = n - i0 < PVB_WGRAD_FOLD_MAX ? n - i0 : PVB_WGRAD_FOLD_MAX;
    int bx = 1;
    for (int i = 0; i < cnt; ++i) {
      a.p[i] = probs[i0 + i];
      PVB_CHECK_ARG(a.p[i].scratch && a.p[i].dW && a.p[i].Cin > 0 && a.p[i].Cout > 0 && a.p[i].taps > 0 &&
                        a.p[i].taps <= FOLD_MAX_TAPS,
                    "pvb_conv_tc_wgrad_fold: bad problem %d", i0 + i);
      const int tiles = ((a.p[i].Cout + 31) / 32) * ((a.p[i].Cin + 31) / 32);
      bx = tiles > bx ? tiles : bx;
    }
    wgrad_finish_kernel<<<dim3((unsigned)bx, (unsigned)cnt), 256, 0, (cudaStream_t)stream>>>(a);
    pvb::count_launch();
  }
  return pvb::launch_status();
}

extern "C" int pvb_conv_tc_wgrad(const float* dpre, const float* x, float* dW, float* db, int B, int Cin,
                                 int Cout, int H, int Wd, int kh, int kw, void* scratch, int fold,
                                 void* stream) {
  PVB_CHECK_ARG(dpre && x && dW, "pvb_conv_tc_wgrad: null pointer");
  PVB_CHECK_ARG(((uintptr_t)scratch % 4) == 0, "pvb_conv_tc_wgrad: scratch must be float-aligned");
  const int Cin_real = Cin;
  if (Cin < 16) Cin = 16;                        // zero-padded x tile
  PVB_CHECK_ARG(tc_ok(Cin, Cout, kh, kw) && Cout <= 128, "pvb_conv_tc_wgrad: unsupported shape");
  PVB_CHECK_ARG(kw * Cin + 16 <= 512 && Cin + 16 <= 256,
                "pvb_conv_tc_wgrad: Cin too large (kw * Cin + 16 TMEM columns, Cin + 16 <= 256 per MMA)");
  if (B == 0) return 0;
  const int pw = kw / 2, Wp = Wd + pw, adv = TP - 2 * pw;
  const int64_t positions = (int64_t)B * H * Wp;          // W-padded pixel positions
  PVB_CHECK_ARG(positions + TP < (1ll << 31) && (int64_t)(Cin > Cout ? Cin : Cout) * H * Wd < (1ll << 31),
                "pvb_conv_tc_wgrad: tensor too large for 32-bit position arithmetic");
  const int n_tiles = (int)((positions + adv - 1) / adv);
  const int stage = WG_A + (Cin / 8) * WCS + WG_ONES + WG_XPAD;
  int n_stages = (227 * 1024 - 128) / stage;
  if (n_stages > WG_MAX_STAGES) n_stages = WG_MAX_STAGES;
  PVB_CHECK_ARG(n_stages >= 2, "pvb_conv_tc_wgrad: tile does not fit shared memory");
  const int smem = n_stages * stage + 128;
  // one CTA per SM (512 TMEM columns): never more CTAs than SMs, or the extras run as a second wave
  int splits = 148 / kh;
  if (splits > n_tiles) splits = n_tiles;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(conv_tc_wgrad_kernel<BWD_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr = true;
  }
  dim3 grid(kh, (unsigned)splits);
  // 16-channel gather items per tile -> producer groups (no more groups than items)
  const int items = Cout / 16 + Cin / 16;
  const int ng = items < WG_MAX_GROUPS ? items : WG_MAX_GROUPS;
  conv_tc_wgrad_kernel<BWD_BF16><<<grid, ng * 128 + 32, smem, (cudaStream_t)stream>>>(
      dpre, x, dW, db, reinterpret_cast<float*>(scratch), B, Cin_real, Cout, H, Wd, kh, kw, n_stages, n_tiles);
  pvb::count_launch();
  if (scratch && fold) {
    const pvb_wgrad_fold pr{reinterpret_cast<float*>(scratch), dW, db, Cin_real, Cout, kh * kw};
    return pvb_conv_tc_wgrad_fold(&pr, 1, stream);
  }
  return pvb::launch_status();
}

#ifdef PVB_TC_TRACE
extern "C" int pvb_pix3_trace_read(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_p3trace, sizeof(long long) * 3 * 64);
}
extern "C" int pvb_wgrad_trace_read(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_wtrace, sizeof(long long) * 2 * 64);
}
extern "C" int pvb_conv_trace_read(long long* out) {
  return (int)cudaMemcpyFromSymbol(out, g_ctrace, sizeof(long long) * 2 * 64);
}
#endif

