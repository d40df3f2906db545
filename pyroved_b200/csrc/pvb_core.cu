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
  int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  int last_t = -1;
  float step_size = 0.f, bc2s = 1.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int64_t j = i0 + k;
    if (j >= n) break;
    int t = first_step ? (first_step[j] < 0 ? 0 : step - first_step[j]) : step;
    if (t <= 0) {           // parameter has never carried a gradient
      if (ZERO_G) g[j] = 0.f;
      continue;
    }
    if (t != last_t) {
      step_size = lr / (1.f - powf(b1, (float)t));
      bc2s = sqrtf(1.f - powf(b2, (float)t));
      last_t = t;
    }
    float gj = g[j];
    if (ZERO_G) g[j] = 0.f;   // consumed: the buffer is clean for the next step's accumulation
    float mj = b1 * m[j] + (1.f - b1) * gj;
    float vj = b2 * v[j] + (1.f - b2) * gj * gj;
    m[j] = mj;
    v[j] = vj;
    p[j] -= step_size * mj / (sqrtf(vj) / bc2s + eps);
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

