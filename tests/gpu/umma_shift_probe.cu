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
  return 0;
}
