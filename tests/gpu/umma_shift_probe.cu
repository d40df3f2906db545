// Probe (GPU only): cycles per MMA (M128 x N64 x K16, SS, K-major no-swizzle row-chunk operands) as a
// function of (a) the A operand's chunk-column stride and (b) a row shift of its start address --
// the two things the tap-reuse convolution kernels vary.  Timing only (operands are zeros).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_shift_probe.bin umma_shift_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../pyroved_b200/csrc/umma.cuh"

template <int N>
__global__ void probe(int a_rows, int shift, int reps, int busy, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (warp == 0) umma::tmem_alloc<256>(&tmem_base);
  if (tid == 0) { umma::mbar_init(&bar, 1); umma::mbar_fence_init(); }
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = tmem_base;
  const int XCS = a_rows * 16;
  if (tid == 0) {
    constexpr uint32_t idesc = umma::idesc_f16(128, N, 0, 0);
    const uint32_t a0 = umma::smem_u32(smem) + shift * 16, b0 = umma::smem_u32(smem) + 128 * 1024;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
      for (int k = 0; k < 4; ++k)
        umma::mma_f16_ss(tm, umma::smem_desc(a0 + k * 2 * XCS, XCS, 128),
                         umma::smem_desc(b0 + k * 2 * N * 16, N * 16, 128), idesc, 1u);
    umma::commit(&bar);
    umma::mbar_wait(&bar, 0);
    cyc[0] = clock64() - t0;
  } else if (busy && warp >= 1) {
    // other warps keep the LSU / shared memory busy with 16-byte stores and loads
    uint32_t addr = umma::smem_u32(smem + 160 * 1024) + tid * 16;
    uint32_t x = tid, y;
    for (int i = 0; i < busy; ++i) {
      asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(addr + (i & 3) * 8192), "r"(x) : "memory");
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(y) : "r"(addr + ((i + 1) & 3) * 8192) : "memory");
      x += y;
    }
    if (x == 0xdeadbeef) cyc[1] = x;
  }
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<256>(tm);
}

// descriptors change with every MMA (9 taps x 4 K-steps per "tile", B operand = resident weights),
// as in conv_tc_pix3_kernel; `mode` 0: computed by additions inside the loop; 1: all 36 pairs of a
// tile precomputed into shared memory and read back as uint64 (one LDS.64 each)
template <int N>
__global__ void probe_vary(int a_rows, int tiles, int mode, int ts, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  __shared__ uint64_t dtab[72];
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 50 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  if (warp == 0) umma::tmem_alloc<256>(&tmem_base);
  if (tid == 0) { umma::mbar_init(&bar, 1); umma::mbar_fence_init(); }
  const int XCS = a_rows * 16;
  const uint64_t da_base = umma::smem_desc(umma::smem_u32(smem), XCS, 128);
  const uint64_t db_base = umma::smem_desc(umma::smem_u32(smem) + 64 * 1024, N * 16, 128);
  const uint64_t da_k = (uint64_t)((2 * XCS) >> 4), db_k = (uint64_t)((2 * N * 16) >> 4);
  if (tid < 36) {
    const int t = tid / 4, k = tid % 4;
    dtab[2 * tid] = da_base + (uint64_t)(33 + (t / 3 - 1) * 33 + (t % 3 - 1)) + k * da_k;
    dtab[2 * tid + 1] = db_base + (uint64_t)((t * N * 128) >> 4) + k * db_k;
  }
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = tmem_base;
  if (tid == 0) {
    constexpr uint32_t idesc = umma::idesc_f16(128, N, 0, 0);
    long long t0 = clock64();
    for (int it = 0; it < tiles; ++it) {
      if (mode == 0) {
        int t = 0;
        for (int th = 0; th < 3; ++th) {
          const int row_shift = 33 + (th - 1) * 33 - 1;
          for (int tw = 0; tw < 3; ++tw, ++t) {
            uint64_t da = da_base + (uint64_t)(row_shift + tw);
            uint64_t db = db_base + (uint64_t)((t * N * 128) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma::mma_f16_ss(tm + (it & 1) * 128, da + k * da_k, db + k * db_k, idesc, 1u);
          }
        }
      } else {
#pragma unroll 4
        for (int i = 0; i < 36; ++i)
          umma::mma_f16_ss(tm + (it & 1) * 128, dtab[2 * i], dtab[2 * i + 1], idesc, 1u);
      }
      if (ts) { umma::commit(&bar); }
    }
    if (!ts) umma::commit(&bar);
    long long t1 = clock64();
    cyc[1] = t1 - t0;                       // issue time
    if (!ts) umma::mbar_wait(&bar, 0);
    cyc[0] = clock64() - t0;                // until complete (ts = 0)
  }
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<256>(tm);
}

template <int N>
void run_vary(const char* name, int mode) {
  long long* dC; cudaMalloc(&dC, 16);
  auto kern = probe_vary<N>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int tiles = 16;
  kern<<<1, 128, 200 * 1024>>>(196, tiles, mode, 0, dC);
  cudaError_t e = cudaDeviceSynchronize();
  long long c[2] = {0, 0}; cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost);
  printf("%-44s : issue %6.1f, complete %6.1f cycles / MMA %s\n", name, (double)c[1] / (tiles * 36),
         (double)c[0] / (tiles * 36), e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(dC);
}

template <int N>
void run(const char* name, int a_rows, int shift, int busy) {
  long long* dC; cudaMalloc(&dC, 16);
  auto kern = probe<N>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int reps = 64;
  kern<<<1, 512, 200 * 1024>>>(a_rows, shift, reps, busy, dC);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0; cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost);
  printf("%-44s rows=%3d shift=%2d busy=%5d : %6.1f cycles / MMA %s\n", name, a_rows, shift, busy,
         (double)c / (reps * 4), e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(dC);
}

int main() {
  run<64>("N=64 aligned stride, no shift", 128, 0, 0);
  run<64>("N=64 aligned stride, shift 1 row", 128, 1, 0);
  run<64>("N=64 aligned stride, shift 4 rows", 128, 4, 0);
  run<64>("N=64 aligned stride, shift 8 rows", 128, 8, 0);
  run<64>("N=64 stride 196 rows (64 mod 128)", 196, 0, 0);
  run<64>("N=64 stride 196 rows, shift 35", 196, 35, 0);
  run<64>("N=64 stride 200 rows (0 mod 128), shift 35", 200, 35, 0);
  run<64>("N=64 stride 200 rows, shift 32", 200, 32, 0);
  run<64>("N=64 aligned, no shift, LSU busy", 128, 0, 20000);
  run<64>("N=64 stride 196, shift 35, LSU busy", 196, 35, 20000);
  run<128>("N=128 aligned stride, no shift", 128, 0, 0);
  run<128>("N=128 aligned stride, shift 1 row", 128, 1, 0);
  run<128>("N=128 aligned, LSU busy", 128, 0, 20000);
  run_vary<64>("N=64 varying descriptors, additions", 0);
  run_vary<64>("N=64 varying descriptors, table", 1);
  run_vary<128>("N=128 varying descriptors, additions", 0);
  return 0;
}
