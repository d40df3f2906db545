// Shared between the two variants of the fused spatial-decoder kernel (pvb_sdec_tc.cu: one tile in
// flight, also the forward-only / inference kernel; pvb_sdec_tc2.cu: forward of tile i interleaved
// with the backward of tile i-1, the training step).
#pragma once
#include "pvb_common.cuh"

namespace pvb_sdec {

struct Params {
  const float* Uv; const float* x; const float* w;
  const float* W1; const float* b1; const float* W2; const float* b2;
  const float* wo; const float* bo;
  const void* Wp;   // optional: W1 | W2 pre-packed as fp16 operand tiles (pvb_sdec_tc_pack_weights)
  float* rowll; float* loc; float* gUv_part; float* wgrad_part;
  int64_t R; int64_t B; int N; int H; int W; int ndim;
  int sampler; int sigmoid_d; float sig; int backward; int64_t tiles;
  int64_t step_q; int step_r; int step_qb;   // (TILE*grid) / N, % N, and step_q % B
};

// one elected thread: both 32 KB weight tiles, already in operand layout, HBM -> shared memory
// through the TMA engine (cp.async.bulk), completion counted on `bar` (expect_tx = 64 KB)
__device__ __forceinline__ void bulk_load_weights(const void* Wp, void* smem_w1, void* smem_w2,
                                                  uint64_t* bar) {
  const uint32_t b = static_cast<uint32_t>(__cvta_generic_to_shared(bar));
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b), "r"(65536u) : "memory");
  const char* src = reinterpret_cast<const char*>(Wp);
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
          static_cast<uint32_t>(__cvta_generic_to_shared(smem_w1))),
      "l"(src), "r"(32768u), "r"(b)
      : "memory");
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
          static_cast<uint32_t>(__cvta_generic_to_shared(smem_w2))),
      "l"(src + 32768), "r"(32768u), "r"(b)
      : "memory");
}

// launches the interleaved (v2) training kernel on `ctas` CTAs; returns a CUDA error code (0 = ok)
int launch_v2(const Params& P, int ctas, cudaStream_t stream);

}  // namespace pvb_sdec
