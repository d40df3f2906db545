// libpvb: error state, version, small utility kernels (reduce, counter, Adam).
#include <atomic>
#include <cstdarg>
#include "pvb_common.cuh"

namespace pvb {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace pvb

extern "C" long long pvb_launch_count(void) { return pvb::g_launches.load(); }
extern "C" int pvb_version(void) { return 100; }
extern "C" const char* pvb_last_error_string(void) { return pvb::g_err; }

// ---------------------------------------------------------------------------
// 32 outputs per CTA; the G partials are split over 8 thread groups whose sums are combined in
// a fixed order -> bitwise reproducible, and short dependent chains even for large G
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ part, float* __restrict__ out, int G, int64_t n,
                       int64_t stride, int accumulate) {
  __shared__ float sm[8][33];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + lane;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (i < n) {
    int g = grp;
    for (; g + 24 < G; g += 32) {
      s0 += part[(int64_t)g * stride + i];
      s1 += part[(int64_t)(g + 8) * stride + i];
      s2 += part[(int64_t)(g + 16) * stride + i];
      s3 += part[(int64_t)(g + 24) * stride + i];
    }
    for (; g < G; g += 8) s0 += part[(int64_t)g * stride + i];
  }
  sm[grp][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (grp == 0 && i < n) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += sm[k][lane];
    out[i] = accumulate ? out[i] + s : s;
  }
}
extern "C" int pvb_reduce_partials(const float* part, float* out, int G, int64_t n,
                                   int64_t part_stride, int accumulate, void* stream) {
  PVB_CHECK_ARG(part && out && G > 0 && n >= 0 && part_stride >= n, "pvb_reduce_partials: bad argument");
  if (n == 0) return 0;
  reduce_partials_kernel<<<pvb::cdiv(n, 32), 256, 0, (cudaStream_t)stream>>>(part, out, G, n,
                                                                              part_stride, accumulate); pvb::count_launch();
  return pvb::launch_status();
}

__global__ void counter_add_kernel(int32_t* c, int32_t v) { *c += v; }
extern "C" int pvb_counter_add(int32_t* counter, int32_t v, void* stream) {
  PVB_CHECK_ARG(counter, "pvb_counter_add: null counter");
  counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter, v); pvb::count_launch();
  return pvb::launch_status();
}

// torch.optim.Adam (defaults; no amsgrad / weight decay):
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2
//   p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
template <bool ZERO_G>
__device__ __forceinline__ void adam_update4(float* __restrict__ p, float* __restrict__ g,
                                             float* __restrict__ m, float* __restrict__ v, int64_t n,
                                             float lr, float b1, float b2, float eps, int step,
                                             const int32_t* __restrict__ first_step) {
  // one 16-byte access per thread and buffer (the buffers are 16-byte aligned and padded to a
  // multiple of four floats by their owners; the last quad of an odd-sized buffer goes scalar)
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i0 >= n) return;
  const bool full = i0 + 4 <= n;
  float gq[4], mq[4], vq[4], pq[4];
  int fq[4] = {0, 0, 0, 0};
  if (full) {
    const float4 g4 = *reinterpret_cast<const float4*>(g + i0);
    const float4 m4 = *reinterpret_cast<const float4*>(m + i0);
    const float4 v4 = *reinterpret_cast<const float4*>(v + i0);
    const float4 p4 = *reinterpret_cast<const float4*>(p + i0);
    gq[0] = g4.x; gq[1] = g4.y; gq[2] = g4.z; gq[3] = g4.w;
    mq[0] = m4.x; mq[1] = m4.y; mq[2] = m4.z; mq[3] = m4.w;
    vq[0] = v4.x; vq[1] = v4.y; vq[2] = v4.z; vq[3] = v4.w;
    pq[0] = p4.x; pq[1] = p4.y; pq[2] = p4.z; pq[3] = p4.w;
    if (first_step) {
      const int4 f4 = *reinterpret_cast<const int4*>(first_step + i0);
      fq[0] = f4.x; fq[1] = f4.y; fq[2] = f4.z; fq[3] = f4.w;
    }
  } else {
    for (int k = 0; k < 4 && i0 + k < n; ++k) {
      gq[k] = g[i0 + k]; mq[k] = m[i0 + k]; vq[k] = v[i0 + k]; pq[k] = p[i0 + k];
      if (first_step) fq[k] = first_step[i0 + k];
    }
  }
  int last_t = -1;
  float step_size = 0.f, bc2s = 1.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (!full && i0 + k >= n) break;
    const int t = first_step ? (fq[k] < 0 ? 0 : step - fq[k]) : step;
    if (t <= 0) continue;   // parameter has never carried a gradient
    if (t != last_t) {
      step_size = lr / (1.f - powf(b1, (float)t));
      bc2s = sqrtf(1.f - powf(b2, (float)t));
      last_t = t;
    }
    const float gj = gq[k];
    mq[k] = b1 * mq[k] + (1.f - b1) * gj;
    vq[k] = b2 * vq[k] + (1.f - b2) * gj * gj;
    pq[k] -= step_size * mq[k] / (sqrtf(vq[k]) / bc2s + eps);
  }
  if (full) {
    *reinterpret_cast<float4*>(m + i0) = make_float4(mq[0], mq[1], mq[2], mq[3]);
    *reinterpret_cast<float4*>(v + i0) = make_float4(vq[0], vq[1], vq[2], vq[3]);
    *reinterpret_cast<float4*>(p + i0) = make_float4(pq[0], pq[1], pq[2], pq[3]);
    // consumed: the gradient buffer is clean for the next step's accumulation
    if (ZERO_G) *reinterpret_cast<float4*>(g + i0) = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    for (int k = 0; k < 4 && i0 + k < n; ++k) {
      m[i0 + k] = mq[k]; v[i0 + k] = vq[k]; p[i0 + k] = pq[k];
      if (ZERO_G) g[i0 + k] = 0.f;
    }
  }
}
__global__ void adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g,
                                 float* __restrict__ m, float* __restrict__ v, int64_t n, float lr,
                                 float b1, float b2, float eps,
                                 const int32_t* __restrict__ step_counter,
                                 const int32_t* __restrict__ first_step) {
  adam_update4<false>(p, const_cast<float*>(g), m, v, n, lr, b1, b2, eps, *step_counter, first_step);
}
// same, with the step increment folded in: all CTAs read the counter before taking a ticket,
// the last ticket holder stores counter + 1 (and re-arms the ticket)
__global__ void adam_flat_step_kernel(float* __restrict__ p, float* __restrict__ g,
                                      float* __restrict__ m, float* __restrict__ v, int64_t n,
                                      float lr, float b1, float b2, float eps,
                                      int32_t* step_counter, const int32_t* __restrict__ first_step,
                                      int32_t* ticket, float* loss_src, float* loss_ring) {
  const int step = *reinterpret_cast<volatile int32_t*>(step_counter) + 1;
  adam_update4<true>(p, g, m, v, n, lr, b1, b2, eps, step, first_step);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    int t = atomicAdd(ticket, 1);
    if (t == (int)gridDim.x - 1) {
      // the step's loss goes straight to (mapped, pinned) host memory: slot = step & (PVB_LOSS_RING - 1)
      if (loss_src) {
        const float L = loss_src[0];
        if (loss_ring) loss_ring[step & (PVB_LOSS_RING - 1)] = L;
        loss_src[1] = L;       // "last loss" slot, read by the host side after the step
        loss_src[0] = 0.f;     // loss accumulator of the next step
      }
      *step_counter = step;
      *ticket = 0;
    }
  }
}
extern "C" int pvb_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, float lr,
                             float beta1, float beta2, float eps, const int32_t* step_counter,
                             const int32_t* first_step, void* stream) {
  PVB_CHECK_ARG(p && g && m && v && step_counter && n >= 0, "pvb_adam_flat: bad argument");
  PVB_CHECK_ARG(((uintptr_t)p % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)m % 16 == 0) &&
                    ((uintptr_t)v % 16 == 0),
                "pvb_adam_flat: buffers must be 16-byte aligned");
  PVB_CHECK_ARG((uintptr_t)first_step % 16 == 0, "pvb_adam_flat: first_step must be 16-byte aligned");
  if (n == 0) return 0;
  int64_t n4 = (n + 3) / 4;
  adam_flat_kernel<<<pvb::cdiv(n4, 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1,
                                                                         beta2, eps, step_counter,
                                                                         first_step); pvb::count_launch();
  return pvb::launch_status();
}

extern "C" int pvb_adam_flat_step(float* p, float* g, float* m, float* v, int64_t n, float lr,
                                  float beta1, float beta2, float eps, int32_t* step_counter,
                                  const int32_t* first_step, int32_t* ticket, float* loss_src,
                                  float* loss_ring, void* stream) {
  PVB_CHECK_ARG(p && g && m && v && step_counter && ticket && n >= 0, "pvb_adam_flat_step: bad argument");
  PVB_CHECK_ARG(!loss_ring || loss_src, "pvb_adam_flat_step: loss_ring needs loss_src");
  PVB_CHECK_ARG(((uintptr_t)p % 16 == 0) && ((uintptr_t)g % 16 == 0) && ((uintptr_t)m % 16 == 0) &&
                    ((uintptr_t)v % 16 == 0),
                "pvb_adam_flat_step: buffers must be 16-byte aligned");
  PVB_CHECK_ARG((uintptr_t)first_step % 16 == 0, "pvb_adam_flat_step: first_step must be 16-byte aligned");
  if (n == 0) return 0;   // (the counter is not advanced for an empty parameter set)
  int64_t n4 = (n + 3) / 4;
  adam_flat_step_kernel<<<pvb::cdiv(n4, 256), 256, 0, (cudaStream_t)stream>>>(
      p, g, m, v, n, lr, beta1, beta2, eps, step_counter, first_step, ticket, loss_src, loss_ring);
  pvb::count_launch();
  return pvb::launch_status();
}

// ---- row gather (on-device shuffle of the GPU-resident loader) -------------------------------------
namespace {
template <int VEC>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx,
                   float* __restrict__ dst, int64_t rows, int64_t row_vecs, int64_t n_src) {
  // one warp walks one row at a time: consecutive lanes read consecutive 4 / 16-byte words
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += nwarps) {
    int64_t s = idx[r];
    if (s < 0 || s >= n_src) continue;          // out-of-range index: row left untouched
    if (VEC == 4) {
      const float4* a = reinterpret_cast<const float4*>(src) + s * row_vecs;
      float4* b = reinterpret_cast<float4*>(dst) + r * row_vecs;
      for (int64_t c = lane; c < row_vecs; c += 32) b[c] = __ldg(a + c);
    } else {
      const float* a = src + s * row_vecs;
      float* b = dst + r * row_vecs;
      for (int64_t c = lane; c < row_vecs; c += 32) b[c] = __ldg(a + c);
    }
  }
}
}  // namespace

extern "C" int pvb_gather_rows(const float* src, const int64_t* idx, float* dst, int64_t rows,
                               int64_t row_floats, int64_t n_src, void* stream) {
  PVB_CHECK_ARG(src && idx && dst && rows >= 0 && row_floats > 0 && n_src > 0,
                "pvb_gather_rows: bad argument");
  if (rows == 0) return 0;
  const bool vec = row_floats % 4 == 0 && (((uintptr_t)src | (uintptr_t)dst) & 15) == 0;
  // 8 warps per CTA, one row per warp per pass; never more CTAs than a few waves of 148 SMs
  int64_t blocks = (rows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (vec)
    gather_rows_kernel<4><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, idx, dst, rows,
                                                                            row_floats / 4, n_src);
  else
    gather_rows_kernel<1><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, idx, dst, rows,
                                                                            row_floats, n_src);
  pvb::count_launch();
  return pvb::launch_status();
}

