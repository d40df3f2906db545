// Standalone probe (GPU only): validates the tcgen05 descriptor encodings and
// the shared-memory operand layouts of pyroved_b200/csrc/umma.cuh against a
// host reference before they are relied on by the fused decoder kernel.
//   build:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe.bin umma_probe.cu
//   run  :  ./umma_probe.bin     (prints max |err| per case, exit 0 iff all pass)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../../pyroved_b200/csrc/umma.cuh"

// case: D[M=128][N] = sum_k A(m,k) B(n,k), K = KT
//  a_mn: A operand read MN-major (buffer stored as [K rows][M cols] row-chunk tile)
//  b_mn: B operand read MN-major (buffer stored as [K rows][N cols])
template <int N, int KT, int A_MN, int B_MN>
__global__ void probe_kernel(const __half* __restrict__ A, const __half* __restrict__ B,
                             float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  constexpr int A_ROWS = A_MN ? KT : 128;   // rows of the stored tile
  constexpr int A_COLS = A_MN ? 128 : KT;
  constexpr int B_ROWS = B_MN ? KT : N;
  constexpr int B_COLS = B_MN ? N : KT;
  uint8_t* sA = smem;
  uint8_t* sB = smem + A_ROWS * A_COLS * 2;
  const int tid = threadIdx.x, warp = tid >> 5;
  // A(m,k): stored element (row,col) = A_MN ? (k,m) : (m,k); global A is [128][KT] row-major
  for (int idx = tid; idx < 128 * KT; idx += blockDim.x) {
    int m = idx / KT, k = idx % KT;
    int r = A_MN ? k : m, c = A_MN ? m : k;
    *reinterpret_cast<__half*>(sA + umma::tile_off(A_ROWS, r, c)) = A[idx];
  }
  for (int idx = tid; idx < N * KT; idx += blockDim.x) {
    int n = idx / KT, k = idx % KT;
    int r = B_MN ? k : n, c = B_MN ? n : k;
    *reinterpret_cast<__half*>(sB + umma::tile_off(B_ROWS, r, c)) = B[idx];
  }
  if (warp == 0) umma::tmem_alloc<256>(&tmem_base);
  if (tid == 0) { umma::mbar_init(&bar, 1); umma::mbar_fence_init(); }
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = tmem_base;
  if (tid == 0) {
    constexpr uint32_t idesc = umma::idesc_f16(128, N, A_MN, B_MN);
    for (int k = 0; k < KT / 16; ++k) {
      uint64_t ad, bd;
      if (A_MN) ad = umma::smem_desc(umma::smem_u32(sA) + k * 256, 128, A_ROWS * 16);
      else      ad = umma::smem_desc(umma::smem_u32(sA) + k * 2 * A_ROWS * 16, A_ROWS * 16, 128);
      if (B_MN) bd = umma::smem_desc(umma::smem_u32(sB) + k * 256, 128, B_ROWS * 16);
      else      bd = umma::smem_desc(umma::smem_u32(sB) + k * 2 * B_ROWS * 16, B_ROWS * 16, 128);
      umma::mma_f16_ss(tm, ad, bd, idesc, k > 0);
    }
    umma::commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::fence_after_sync();
  // each of the 4 warps reads its 32 lanes
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    umma::tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
    umma::tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(warp * 32 + (tid & 31)) * N + c0 + j] = v[j];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<256>(tm);
}

// case: A operand in tensor memory (written with tcgen05.st, lane = row, fp16 pairs packed
// per 32-bit column), B in shared memory: D[128][N] = sum_k A(m,k) B(n,k)
template <int N, int KT, int B_MN>
__global__ void probe_ts_kernel(const __half* __restrict__ A, const __half* __restrict__ B,
                                float* __restrict__ D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  constexpr int B_ROWS = B_MN ? KT : N;
  uint8_t* sB = smem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int idx = tid; idx < N * KT; idx += blockDim.x) {
    int n = idx / KT, k = idx % KT;
    int r = B_MN ? k : n, c = B_MN ? n : k;
    *reinterpret_cast<__half*>(sB + umma::tile_off(B_ROWS, r, c)) = B[idx];
  }
  if (warp == 0) umma::tmem_alloc<512>(&tmem_base);
  if (tid == 0) { umma::mbar_init(&bar, 1); umma::mbar_fence_init(); }
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = tmem_base;
  const uint32_t TM_A = 256;   // A tile: columns [256, 256 + KT/2)
  {
    const int row = warp * 32 + lane;
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(A + row * KT);   // fp16 pairs
    for (int c4 = 0; c4 < KT / 2; c4 += 4) {
      uint4 r = make_uint4(arow[c4], arow[c4 + 1], arow[c4 + 2], arow[c4 + 3]);
      umma::tmem_st4(tm + ((uint32_t)(warp * 32) << 16) + TM_A + c4, r);
    }
    umma::tmem_st_wait();
  }
  umma::fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    umma::fence_after_sync();
    constexpr uint32_t idesc = umma::idesc_f16(128, N, 0, B_MN);
    for (int k = 0; k < KT / 16; ++k) {
      uint64_t bd;
      if (B_MN) bd = umma::smem_desc(umma::smem_u32(sB) + k * 256, 128, B_ROWS * 16);
      else      bd = umma::smem_desc(umma::smem_u32(sB) + k * 2 * B_ROWS * 16, B_ROWS * 16, 128);
      umma::mma_f16_ts(tm, tm + TM_A + k * 8, bd, idesc, k > 0);
    }
    umma::commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    umma::tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
    umma::tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(warp * 32 + lane) * N + c0 + j] = v[j];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc<512>(tm);
}

template <int N, int KT, int A_MN, int B_MN>
bool run_case(const char* name) {
  std::vector<__half> hA(128 * KT), hB(N * KT);
  std::vector<float> fA(128 * KT), fB(N * KT), ref(128 * N), out(128 * N);
  srand(7 + N + KT + A_MN * 2 + B_MN);
  for (size_t i = 0; i < hA.size(); ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2half(v); fA[i] = __half2float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2half(v); fB[i] = __half2float(hB[i]); }
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < KT; ++k) s += (double)fA[m * KT + k] * fB[n * KT + k];
      ref[m * N + n] = (float)s;
    }
  __half *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, out.size() * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, out.size() * 4);
  size_t smem = (128 * KT + N * KT) * 2 + 1024;
  if (A_MN == 2) {   // A from tensor memory
    cudaFuncSetAttribute(probe_ts_kernel<N, KT, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_ts_kernel<N, KT, B_MN><<<1, 128, smem>>>(dA, dB, dD);
  } else {
    cudaFuncSetAttribute(probe_kernel<N, KT, (A_MN & 1), B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_kernel<N, KT, (A_MN & 1), B_MN><<<1, 128, smem>>>(dA, dB, dD);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-34s CUDA error: %s\n", name, cudaGetErrorString(e)); return false; }
  cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (size_t i = 0; i < out.size(); ++i) maxerr = fmax(maxerr, fabs((double)out[i] - ref[i]));
  bool ok = maxerr < 2e-3;
  printf("%-34s N=%3d K=%3d  max|err| = %.3e  %s\n", name, N, KT, maxerr, ok ? "PASS" : "FAIL");
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return ok;
}

int main() {
  bool ok = true;
  ok &= run_case<128, 16, 0, 0>("A K-major, B K-major (1 mma)");
  ok &= run_case<128, 128, 0, 0>("A K-major, B K-major");
  ok &= run_case<128, 128, 0, 1>("A K-major, B MN-major");
  ok &= run_case<128, 128, 1, 1>("A MN-major, B MN-major");
  ok &= run_case<144, 128, 1, 1>("A MN-major, B MN-major N=144");
  ok &= run_case<16, 128, 1, 1>("A MN-major, B MN-major N=16");
  ok &= run_case<16, 128, 0, 0>("A K-major, B K-major N=16");
  ok &= run_case<128, 128, 2, 0>("A in TMEM, B K-major");
  ok &= run_case<128, 128, 2, 1>("A in TMEM, B MN-major");
  printf(ok ? "UMMA PROBE: ALL PASS\n" : "UMMA PROBE: FAILURES\n");
  return ok ? 0 : 1;
}
