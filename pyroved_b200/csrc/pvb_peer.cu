// Data-parallel exchange step of the SVI update: the SUM all-reduce of the flat
// [gradients | loss] buffer fused with the Adam update, in ONE kernel over NVLink / NVSwitch peer
// memory.  Every rank's gradient buffer lives in symmetric memory mapped into all ranks (one
// process per GPU); each rank loads every peer's buffer directly (one-shot: (world-1) x 0.6 MB
// inbound per GPU for cfg2), adds the values in rank order 0..world-1 -- the same order on every
// rank, so the replicas stay bit-identical -- and applies Adam to its own replica of the
// parameters.  Replaces ncclAllReduce + the Adam kernel + the two graph boundaries between them;
// being a plain kernel it is captured in the step's CUDA graph.
//
// Cross-GPU synchronisation: two sets of epoch flags per rank in symmetric memory.
//   ready[r] (written by rank r): rank r's gradients of this epoch are complete
//   done[r]  (written by rank r): rank r has finished reading every peer's gradients
// A kernel starts reading when all ready flags carry its epoch and retires when all done flags do
// (its own gradient buffer may be zeroed by the next step only after every peer has read it).
// No CTA waits on another CTA of the same grid except through the final ticket, and remote ranks
// only wait on flags written by CTA 0 at its start / the last CTA at its end, so there is no
// circular wait as long as every rank launches the kernel (the step is SPMD).
#include "pvb_common.cuh"

namespace {

constexpr int MAX_WORLD = 16;
constexpr int NT = 256;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer1(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

// Wait until a peer's flag reaches this epoch.  Peers arrive within microseconds in steady state and
// within seconds at start-up skew; a peer that never arrives (its process died) would hang this GPU,
// so after ~2 minutes the kernel traps and the failure surfaces as a CUDA error instead.
__device__ __forceinline__ void spin_until(const uint32_t* flag, uint32_t e) {
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(flag) - e) < 0) {
    if (clock64() - t0 > 240000000000ll) __trap();
  }
}

// state: [0] epoch of the last finished invocation, [1] ticket, [2] loss (float bits)
__global__ void __launch_bounds__(NT)
peer_allreduce_adam_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                           float* own_g, int64_t n, const float* const* __restrict__ peer_g,
                           uint32_t* const* __restrict__ peer_flags, int32_t* state, int rank, int world,
                           float lr, float b1, float b2, float eps, int32_t* step_counter,
                           const int32_t* __restrict__ first_step, float* loss_ring) {
  __shared__ const float* gp[MAX_WORLD];
  __shared__ int s_last;
  const uint32_t e = (uint32_t)(*reinterpret_cast<volatile int32_t*>(state)) + 1u;
  const int step = *reinterpret_cast<volatile int32_t*>(step_counter) + 1;
  uint32_t* mine = peer_flags[rank];
  if (threadIdx.x < world) {
    gp[threadIdx.x] = peer_g[threadIdx.x];
    if (blockIdx.x == 0) {
      __threadfence_system();
      st_release_sys(peer_flags[threadIdx.x] + rank, e);            // ready[rank] on every peer
    }
    spin_until(mine + threadIdx.x, e);                              // all peers ready
  }
  __syncthreads();
  const int64_t n4 = (n + 3) / 4;
  for (int64_t i4 = (int64_t)blockIdx.x * NT + threadIdx.x; i4 < n4; i4 += (int64_t)gridDim.x * NT) {
    const int64_t j0 = i4 * 4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < world; ++r) {            // fixed order: identical sums on every rank
      const float4 a = ld_peer4(gp[r] + j0);
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    }
    const float g4[4] = {s.x, s.y, s.z, s.w};
    int last_t = -1;
    float step_size = 0.f, bc2s = 1.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t j = j0 + k;
      if (j >= n) break;
      const int t = first_step ? (first_step[j] < 0 ? 0 : step - first_step[j]) : step;
      if (t <= 0) continue;                      // parameter has never carried a gradient
      if (t != last_t) {
        step_size = lr / (1.f - powf(b1, (float)t));
        bc2s = sqrtf(1.f - powf(b2, (float)t));
        last_t = t;
      }
      const float gj = g4[k];
      const float mj = b1 * m[j] + (1.f - b1) * gj;
      const float vj = b2 * v[j] + (1.f - b2) * gj * gj;
      m[j] = mj;
      v[j] = vj;
      p[j] -= step_size * mj / (sqrtf(vj) / bc2s + eps);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {     // the loss slot follows the n gradients
    float L = 0.f;
    for (int r = 0; r < world; ++r) L += ld_peer1(gp[r] + n);
    state[2] = __float_as_int(L);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(state + 1, 1) == (int)gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  // last CTA of this rank: every peer buffer has been read
  if (threadIdx.x < world) {
    st_release_sys(peer_flags[threadIdx.x] + MAX_WORLD + rank, e);   // done[rank] on every peer
    spin_until(mine + MAX_WORLD + threadIdx.x, e);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const float L = __int_as_float(*reinterpret_cast<volatile int32_t*>(state + 2));
    own_g[n] = L;                                   // global loss
    if (loss_ring) loss_ring[step & (PVB_LOSS_RING - 1)] = L;         // ... and straight to mapped host memory
    *step_counter = step;
    state[1] = 0;
    state[0] = (int32_t)e;
  }
}

}  // namespace

extern "C" int pvb_peer_flag_words(void) { return 2 * MAX_WORLD; }

extern "C" int pvb_peer_allreduce_adam(float* p, float* m, float* v, float* own_g, int64_t n,
                                       const void* peer_g, const void* peer_flags, int32_t* state,
                                       int rank, int world, float lr, float beta1, float beta2, float eps,
                                       int32_t* step_counter, const int32_t* first_step, float* loss_ring,
                                       void* stream) {
  PVB_CHECK_ARG(p && m && v && own_g && peer_g && peer_flags && state && step_counter,
                "pvb_peer_allreduce_adam: null pointer");
  PVB_CHECK_ARG(world >= 1 && world <= MAX_WORLD && rank >= 0 && rank < world,
                "pvb_peer_allreduce_adam: bad rank / world (<= 16 ranks)");
  PVB_CHECK_ARG(n > 0 && n % 4 == 0, "pvb_peer_allreduce_adam: n must be a positive multiple of 4");
  PVB_CHECK_ARG(((uintptr_t)p % 16 == 0) && ((uintptr_t)own_g % 16 == 0) && ((uintptr_t)m % 16 == 0) &&
                    ((uintptr_t)v % 16 == 0),
                "pvb_peer_allreduce_adam: buffers must be 16-byte aligned");
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + NT - 1) / NT;
  if (blocks > 148 * 4) blocks = 148 * 4;       // all CTAs co-resident
  peer_allreduce_adam_kernel<<<(unsigned)blocks, NT, 0, (cudaStream_t)stream>>>(
      p, m, v, own_g, n, reinterpret_cast<const float* const*>(peer_g),
      reinterpret_cast<uint32_t* const*>(peer_flags), state, rank, world, lr, beta1, beta2, eps,
      step_counter, first_step, loss_ring);
  pvb::count_launch();
  return pvb::launch_status();
}
