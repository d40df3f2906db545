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
  float* rowll; float* loc; float* gUv_part; float* wgrad_part;
  int64_t R; int64_t B; int N; int H; int W; int ndim;
  int sampler; int sigmoid_d; float sig; int backward; int64_t tiles;
  int64_t step_q; int step_r; int step_qb;   // (TILE*grid) / N, % N, and step_q % B
};

// launches the interleaved (v2) training kernel on `ctas` CTAs; returns a CUDA error code (0 = ok)
int launch_v2(const Params& P, int ctas, cudaStream_t stream);

}  // namespace pvb_sdec
