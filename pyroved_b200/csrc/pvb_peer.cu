// Data-parallel exchange step of the SVI update: the SUM all-reduce of the flat
// [gradients | loss] buffer fused with the Adam update, in ONE kernel over NVLink / NVSwitch peer
// memory (one process per GPU).  Replaces ncclAllReduce + the Adam kernel + the two graph
// boundaries between them; being a plain kernel it is captured in the step's CUDA graph.
//
// Every rank owns two STAGING buffers in symmetric memory (mapped into all ranks), used by epoch
// parity.  One invocation (epoch e, parity e & 1):
//   1. copy the local gradients into the own staging buffer of this parity; the last CTA to finish
//      publishes ready[rank] = e to every peer (st.release.sys);
//   2. wait until every peer's ready flag carries e;
//   3. one-shot (small worlds): every rank loads every peer's staging buffer ((world-1) x size
//      inbound), adds in rank order 0..world-1 -- the same order on every rank, so the replicas stay
//      bit-identical -- and applies Adam to its replica;
//      two-shot (world >= 4): rank r first reduces only slice r of every peer's buffer, in rank
//      order, IN PLACE into its own staging buffer, publishes reduced[rank] = e, waits for the
//      peers' flags and then gathers every reduced slice from its owner (inbound 2 (world-1)/world
//      x size instead of (world-1) x size); every rank reads the same reduced values, so the
//      replicas stay bit-identical here too;
//   4. the gradient buffer is zeroed for the next step as it is consumed (no separate memset), the
//      loss slots are summed, the step counter advances.
// There is NO "done reading" round trip: a staging buffer of parity p is rewritten at epoch e + 2,
// which this rank reaches only after it has seen ready[e + 1] from every peer -- and a peer
// publishes ready[e + 1] from the kernel that follows, in stream order, the one in which it read
// epoch e.  No CTA waits on another CTA of the same grid except through tickets whose last
// arriver publishes the flag everybody then spins on, so there is no circular wait as long as
// every rank launches the kernel (the step is SPMD) and the grid is co-resident (<= 4 CTAs / SM).
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

// state (int32): [0] epoch of the last finished invocation, [1] [2] [3] tickets, [4] loss bits
__device__ __forceinline__ bool grid_arrive_last(int32_t* ticket, int* s_last) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();          // this CTA's writes (staging / reduced slice) before the ticket
    *s_last = atomicAdd(ticket, 1) == (int)gridDim.x - 1;
    if (*s_last) __threadfence_system();
  }
  __syncthreads();
  return *s_last != 0;
}

// Adam on one quad (n % 4 == 0, 16-byte aligned buffers: one 16-byte access per buffer), in two halves: the
// optimizer state is fetched BEFORE the (volatile) peer loads are issued, so its latency runs under theirs
struct AdamQuad {
  float4 m4, v4, p4;
  int4 f4;
};
__device__ __forceinline__ AdamQuad adam_load(const float* __restrict__ p, const float* __restrict__ m,
                                              const float* __restrict__ v,
                                              const int32_t* __restrict__ first_step, int64_t j0) {
  AdamQuad q;
  q.m4 = *reinterpret_cast<const float4*>(m + j0);
  q.v4 = *reinterpret_cast<const float4*>(v + j0);
  q.p4 = *reinterpret_cast<const float4*>(p + j0);
  q.f4 = first_step ? *reinterpret_cast<const int4*>(first_step + j0) : make_int4(0, 0, 0, 0);
  return q;
}
__device__ __forceinline__ void adam_apply(const AdamQuad& q, float* __restrict__ p, float* __restrict__ m,
                                           float* __restrict__ v, const float* g4, int64_t j0, float lr,
                                           float b1, float b2, float eps, int step, bool per_param) {
  float mq[4] = {q.m4.x, q.m4.y, q.m4.z, q.m4.w}, vq[4] = {q.v4.x, q.v4.y, q.v4.z, q.v4.w};
  float pq[4] = {q.p4.x, q.p4.y, q.p4.z, q.p4.w};
  const int fq[4] = {q.f4.x, q.f4.y, q.f4.z, q.f4.w};
  int last_t = -1;
  float step_size = 0.f, bc2s = 1.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int t = per_param ? (fq[k] < 0 ? 0 : step - fq[k]) : step;
    if (t <= 0) continue;                      // parameter has never carried a gradient
    if (t != last_t) {
      step_size = lr / (1.f - powf(b1, (float)t));
      bc2s = sqrtf(1.f - powf(b2, (float)t));
      last_t = t;
    }
    const float gj = g4[k];
    mq[k] = b1 * mq[k] + (1.f - b1) * gj;
    vq[k] = b2 * vq[k] + (1.f - b2) * gj * gj;
    pq[k] -= step_size * mq[k] / (sqrtf(vq[k]) / bc2s + eps);
  }
  *reinterpret_cast<float4*>(m + j0) = make_float4(mq[0], mq[1], mq[2], mq[3]);
  *reinterpret_cast<float4*>(v + j0) = make_float4(vq[0], vq[1], vq[2], vq[3]);
  *reinterpret_cast<float4*>(p + j0) = make_float4(pq[0], pq[1], pq[2], pq[3]);
}

// quad i4 of every rank's buffer, summed in rank order 0 ... world-1 (the same order on every rank: the replicas
// stay bit-identical).  ALL peer loads are issued before the first add: as a load-add loop the seven NVLink round
// trips of an 8-rank step ran one after the other (each thread owns a single quad at these sizes).
// (eight ranks per batch: 32 registers of loads in flight; the grid must stay co-resident, see the host side)
__device__ __forceinline__ float4 sum_ranks(float* const* sp, const float4* g4p, int64_t i4, int rank, int world) {
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r0 = 0; r0 < world; r0 += 8) {
    float4 a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int r = r0 + u;
      if (r < world) a[u] = (r == rank) ? g4p[i4] : ld_peer4(sp[r] + 4 * i4);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
      if (r0 + u < world) {
        s.x += a[u].x; s.y += a[u].y; s.z += a[u].z; s.w += a[u].w;
      }
  }
  return s;
}

__global__ void __launch_bounds__(NT)
peer_exchange_adam_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                          float* g, int64_t n, float* const* __restrict__ stage_ptrs,
                          uint32_t* const* __restrict__ peer_flags, int32_t* state, int rank, int world,
                          int two_shot, float lr, float b1, float b2, float eps, int32_t* step_counter,
                          const int32_t* __restrict__ first_step, float* loss_ring) {
  __shared__ float* sp[MAX_WORLD];      // staging buffers of this epoch's parity, by rank
  __shared__ int s_last;
  const uint32_t e = (uint32_t)(*reinterpret_cast<volatile int32_t*>(state)) + 1u;
  const int step = *reinterpret_cast<volatile int32_t*>(step_counter) + 1;
  uint32_t* mine = peer_flags[rank];
  if (threadIdx.x < world) sp[threadIdx.x] = stage_ptrs[(e & 1u) * world + threadIdx.x];
  __syncthreads();
  float4* my_stage = reinterpret_cast<float4*>(sp[rank]);
  float4* g4p = reinterpret_cast<float4*>(g);
  const int64_t n4 = n / 4;                   // n % 4 == 0; quad n4 holds [loss, last loss, pad, pad]
  const int64_t tid0 = (int64_t)blockIdx.x * NT + threadIdx.x, stride = (int64_t)gridDim.x * NT;

  // ---- 1. local gradients (+ loss quad) -> own staging buffer; the last CTA publishes `ready` ----
  for (int64_t i4 = tid0; i4 <= n4; i4 += stride) my_stage[i4] = g4p[i4];
  if (grid_arrive_last(state + 1, &s_last) && threadIdx.x < world)
    st_release_sys(peer_flags[threadIdx.x] + rank, e);
  // ---- 2. all peers ready ----
  if (threadIdx.x < world) spin_until(mine + threadIdx.x, e);
  __syncthreads();

  if (blockIdx.x == 0) {                          // global loss: slot n of every buffer, one lane per rank
    __shared__ float s_loss[MAX_WORLD];
    if (threadIdx.x < world) s_loss[threadIdx.x] = (threadIdx.x == rank) ? g[n] : ld_peer1(sp[threadIdx.x] + n);
    __syncthreads();
    if (threadIdx.x == 0) {
      float L = 0.f;
      for (int r = 0; r < world; ++r) L += s_loss[r];
      state[4] = __float_as_int(L);
    }
  }

  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!two_shot) {
    // ---- 3. one-shot: every rank sums every buffer, fixed order ----
    for (int64_t i4 = tid0; i4 < n4; i4 += stride) {
      const AdamQuad q = adam_load(p, m, v, first_step, 4 * i4);
      const float4 s = sum_ranks(sp, g4p, i4, rank, world);
      const float gs[4] = {s.x, s.y, s.z, s.w};
      adam_apply(q, p, m, v, gs, 4 * i4, lr, b1, b2, eps, step, first_step != nullptr);
      g4p[i4] = zero4;                           // consumed: clean for the next step
    }
  } else {
    // ---- 3a. reduce-scatter: this rank sums slice `rank` of every buffer, in place ----
    const int64_t per = (n4 + world - 1) / world;
    const int64_t lo = (int64_t)rank * per, hi = (lo + per < n4) ? lo + per : n4;
    for (int64_t i4 = lo + tid0; i4 < hi; i4 += stride) my_stage[i4] = sum_ranks(sp, g4p, i4, rank, world);
    if (grid_arrive_last(state + 2, &s_last) && threadIdx.x < world)
      st_release_sys(peer_flags[threadIdx.x] + MAX_WORLD + rank, e);
    if (threadIdx.x < world) spin_until(mine + MAX_WORLD + threadIdx.x, e);
    __syncthreads();
    // ---- 3b. all-gather: every reduced slice from its owner ----
    for (int64_t i4 = tid0; i4 < n4; i4 += stride) {
      const int owner = (int)(i4 / per);
      const AdamQuad q = adam_load(p, m, v, first_step, 4 * i4);
      const float4 s = ld_peer4(sp[owner] + 4 * i4);  // own slice too: written by other CTAs of this grid
      const float gs[4] = {s.x, s.y, s.z, s.w};
      adam_apply(q, p, m, v, gs, 4 * i4, lr, b1, b2, eps, step, first_step != nullptr);
      g4p[i4] = zero4;
    }
  }
  // ---- 4. last CTA of this rank: loss, counters, tickets ----
  if (!grid_arrive_last(state + 3, &s_last)) return;
  if (threadIdx.x == 0) {
    const float L = __int_as_float(*reinterpret_cast<volatile int32_t*>(state + 4));
    g[n] = 0.f;                                     // loss accumulator of the next step
    g[n + 1] = L;                                   // global loss of this step (read by the host side)
    if (loss_ring) loss_ring[step & (PVB_LOSS_RING - 1)] = L;   // ... and straight to mapped host memory
    *step_counter = step;
    state[1] = 0;
    state[2] = 0;
    state[3] = 0;
    __threadfence();
    state[0] = (int32_t)e;
  }
}

}  // namespace

extern "C" int pvb_peer_flag_words(void) { return 2 * MAX_WORLD; }
extern "C" int pvb_peer_state_words(void) { return 8; }

extern "C" int pvb_peer_allreduce_adam(float* p, float* m, float* v, float* g, int64_t n,
                                       const void* stage_ptrs, const void* peer_flags, int32_t* state,
                                       int rank, int world, int two_shot, float lr, float beta1,
                                       float beta2, float eps, int32_t* step_counter,
                                       const int32_t* first_step, float* loss_ring, void* stream) {
  PVB_CHECK_ARG(p && m && v && g && stage_ptrs && peer_flags && state && step_counter,
                "pvb_peer_allreduce_adam: null pointer");
  PVB_CHECK_ARG(world >= 1 && world <= MAX_WORLD && rank >= 0 && rank < world,
                "pvb_peer_allreduce_adam: bad rank / world (<= 16 ranks)");
  PVB_CHECK_ARG(n > 0 && n % 4 == 0, "pvb_peer_allreduce_adam: n must be a positive multiple of 4");
  PVB_CHECK_ARG(((uintptr_t)p % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)m % 16 == 0) &&
                    ((uintptr_t)v % 16 == 0),
                "pvb_peer_allreduce_adam: buffers must be 16-byte aligned");
  PVB_CHECK_ARG((uintptr_t)first_step % 16 == 0, "pvb_peer_allreduce_adam: first_step must be 16-byte aligned");
  const int64_t n4 = n / 4;
  int64_t blocks = (n4 + NT - 1) / NT;
  // the kernel waits grid-wide (tickets, peer flags), so every CTA must be resident at once: the cap comes from
  // the occupancy the driver reports for this build of the kernel, not from an assumed register count
  static int max_resident = 0;
  if (max_resident == 0) {
    int per_sm = 0, dev = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, peer_exchange_adam_kernel, NT, 0) != cudaSuccess ||
        cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || per_sm < 1 || sms < 1) {
      per_sm = 1;
      sms = 148;
    }
    max_resident = per_sm * sms;
  }
  if (blocks > max_resident) blocks = max_resident;
  peer_exchange_adam_kernel<<<(unsigned)blocks, NT, 0, (cudaStream_t)stream>>>(
      p, m, v, g, n, reinterpret_cast<float* const*>(stage_ptrs),
      reinterpret_cast<uint32_t* const*>(peer_flags), state, rank, world, two_shot ? 1 : 0, lr, beta1,
      beta2, eps, step_counter, first_step, loss_ring);
  pvb::count_launch();
  return pvb::launch_status();
}
