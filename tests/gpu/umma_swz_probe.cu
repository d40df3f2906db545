// Standalone probe (GPU only): (1) validates the SWIZZLE_128B shared-memory operand layout and its
// descriptors (K-major and MN-major reads of ONE physical tile) against a host reference, and
// (2) measures the tensor core's operand-fetch rate for the no-swizzle row-chunk layout against
// the 128-byte-swizzled one (cycles per M128 x N x K16 MMA, SS form, back-to-back issue).
//   build:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_swz_probe.bin umma_swz_probe.cu
//   run  :  ./umma_swz_probe.bin
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../../pyroved_b200/csrc/umma.cuh"

// physical SW128 tile of ROWS x 64 fp16: atoms of 8 rows x 128 bytes, 16-byte pieces XOR-ed with row % 8
__host__ __device__ inline uint32_t sw128_off(int r, int c) {   // c in [0, 64)
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((c >> 3) ^ (r & 7)) & 7) << 4) + (c & 7) * 2);
}
__device__ inline uint64_t desc_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;     // SWIZZLE_128B
  return d;
}

// MODE 0: A, B K-major.  A(m,k): tile rows = m (128), cols = k; B(n,k): rows = n, cols = k.  K = 64.
// MODE 1: A, B MN-major. A(m,k): tile rows = k, cols = m (two 64-col blocks); B(n,k): rows = k, cols = n.
//         K = 64 rows, M = 128, N = 128: blocks of 64 columns are BLK bytes apart.
// SWZ 0: the same with the row-chunk no-swizzle layout of umma.cuh.
template <int MODE, int SWZ, int N, int REP>
__global__ void probe(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D,
                      long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  constexpr int KT = 64;
  const int tid = threadIdx.x, warp = tid >> 5;
  // tile geometry
  constexpr int A_ROWS = MODE ? KT : 128, A_COLS = MODE ? 128 : KT;
  constexpr int B_ROWS = MODE ? KT : N, B_COLS = MODE ? N : KT;
  constexpr int A_BLK = A_ROWS * 128;      // bytes of one 64-column block (SW128)
  constexpr int B_BLK = B_ROWS * 128;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 32768;
  for (int idx = tid; idx < 128 * KT; idx += blockDim.x) {
    int m = idx / KT, k = idx % KT;
    int r = MODE ? k : m, c = MODE ? m : k;
    uint32_t off = SWZ ? (uint32_t)((c >> 6) * A_BLK) + sw128_off(r, c & 63) : umma::tile_off(A_ROWS, r, c);
    *reinterpret_cast<__half*>(sA + off) = A[idx];
  }
  for (int idx = tid; idx < N * KT; idx += blockDim.x) {
    int n = idx / KT, k = idx % KT;
    int r = MODE ? k : n, c = MODE ? n : k;
    uint32_t off = SWZ ? (uint32_t)((c >> 6) * B_BLK) + sw128_off(r, c & 63) : umma::tile_off(B_ROWS, r, c);
    *reinterpret_cast<__half*>(sB + off) = B[idx];
  }
  (void)A_COLS; (void)B_COLS;
  if (warp == 0) umma::tmem_alloc<256>(&tmem_base);
  if (tid == 0) { umma::mbar_init(&bar, 1); umma::mbar_fence_init(); }
  umma::fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tm = tmem_base;
  if (tid == 0) {
    constexpr uint32_t idesc = umma::idesc_f16(128, N, MODE, MODE);
    const uint32_t a0 = umma::smem_u32(sA), b0 = umma::smem_u32(sB);
    long long t0 = clock64();
    for (int rep = 0; rep < REP; ++rep) {
      for (int k = 0; k < KT / 16; ++k) {
        uint64_t ad, bd;
        if (SWZ) {
          if (MODE == 0) {       // K-major: 32 bytes per K-step inside the 128-byte row; SBO = 8-row group
            ad = desc_sw128(a0 + k * 32, 16, 1024);
            bd = desc_sw128(b0 + k * 32, 16, 1024);
          } else {               // MN-major: 16 rows (2 atoms) per K-step; LBO = 64-column block stride
            ad = desc_sw128(a0 + k * 2048, A_BLK, 1024);
            bd = desc_sw128(b0 + k * 2048, B_BLK, 1024);
          }
        } else {
          if (MODE) {
            ad = umma::smem_desc(a0 + k * 256, 128, A_ROWS * 16);
            bd = umma::smem_desc(b0 + k * 256, 128, B_ROWS * 16);
          } else {
            ad = umma::smem_desc(a0 + k * 2 * A_ROWS * 16, A_ROWS * 16, 128);
            bd = umma::smem_desc(b0 + k * 2 * B_ROWS * 16, B_ROWS * 16, 128);
          }
        }
        umma::mma_f16_ss(tm, ad, bd, idesc, (rep > 0 || k > 0) ? 1u : 0u);
      }
    }
    umma::commit(&bar);
    umma::mbar_wait(&bar, 0);
    long long t1 = clock64();
    cyc[0] = (t1 - t0);
  }
  umma::mbar_wait(&bar, 0);
  umma::fence_after_sync();
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

template <int MODE, int SWZ, int N, int REP>
bool run(const char* name) {
  constexpr int KT = 64;
  std::vector<__half> hA(128 * KT), hB(N * KT);
  std::vector<float> fA(128 * KT), fB(N * KT);
  srand(7 + MODE * 3 + SWZ);
  for (int i = 0; i < 128 * KT; ++i) { float v = (rand() % 17 - 8) / 8.f; hA[i] = __float2half(v); fA[i] = v; }
  for (int i = 0; i < N * KT; ++i) { float v = (rand() % 13 - 6) / 8.f; hB[i] = __float2half(v); fB[i] = v; }
  __half *dA, *dB; float* dD; long long* dC;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * N * 4); cudaMalloc(&dC, 8);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  auto kern = probe<MODE, SWZ, N, REP>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  kern<<<1, 128, 65536>>>(dA, dB, dD, dC);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-40s CUDA error: %s\n", name, cudaGetErrorString(e)); return false; }
  std::vector<float> hD(128 * N); long long cyc = 0;
  cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < KT; ++k) s += (double)fA[m * KT + k] * fB[n * KT + k];
      maxerr = fmax(maxerr, fabs(s * REP - hD[m * N + n]));
    }
  bool ok = maxerr < 1e-2 * REP;
  printf("%-40s N=%3d  max|err| = %.3e %s   %.1f cycles / MMA (M128 N%d K16)\n", name, N, maxerr,
         ok ? "PASS" : "FAIL", (double)cyc / (REP * KT / 16), N);
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  return ok;
}

int main() {
  bool ok = true;
  ok &= run<0, 0, 128, 1>("K-major  no-swizzle (numerics)");
  ok &= run<0, 1, 128, 1>("K-major  SW128      (numerics)");
  ok &= run<1, 0, 128, 1>("MN-major no-swizzle (numerics)");
  ok &= run<1, 1, 128, 1>("MN-major SW128      (numerics)");
  run<0, 0, 128, 64>("K-major  no-swizzle x64");
  run<0, 1, 128, 64>("K-major  SW128      x64");
  run<1, 0, 128, 64>("MN-major no-swizzle x64");
  run<1, 1, 128, 64>("MN-major SW128      x64");
  run<0, 0, 64, 64>("K-major  no-swizzle x64");
  run<0, 1, 64, 64>("K-major  SW128      x64");
  run<0, 1, 256, 32>("K-major  SW128      x32");
  run<0, 0, 256, 32>("K-major  no-swizzle x32");
  printf(ok ? "SWZ PROBE: NUMERICS PASS\n" : "SWZ PROBE: NUMERICS FAILURES\n");
  return 0;
}
