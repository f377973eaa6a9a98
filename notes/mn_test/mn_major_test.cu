// Stand-alone check of the MN-major TF32 operand layout for tcgen05.mma (SWIZZLE_128B_BASE32B, notes/README.md):
// D[m][n] = sum_k A[k][m] * B[k][n] from row-major [K][M] / [K][N] tiles (K = rows) staged by threads in the canonical
// atom layout. Inputs are small multiples of 1/8 (exact in TF32), so the result must equal the fp32 reference exactly.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o mn_major_test mn_major_test.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include "../../yelprecommendation_b200/csrc/tc_common.cuh"

using namespace yr;

constexpr int M = 128, N = 64, K = 64;

__device__ __forceinline__ uint32_t mn_off(int mn, int k, uint32_t lbo, uint32_t sbo) {
  return (uint32_t)(k >> 2) * sbo + (uint32_t)(mn >> 5) * lbo + (uint32_t)(k & 3) * 128u +
         ((((uint32_t)(mn & 31) >> 3) ^ (uint32_t)(k & 3)) << 5) + (uint32_t)(mn & 7) * 4u;
}

__device__ __forceinline__ uint64_t mn_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo, int layout) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1)
mn_test_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int swap_lbo_sbo) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (s2u(smem_raw) & 1023u)) & 1023u);
  unsigned char* As = smem;                 // 32 KB
  unsigned char* Bs = smem + 32768;         // 16 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t lboA = 512, sboA = (M / 32) * 512, lboB = 512, sboB = (N / 32) * 512;
  if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(tmem_slot, 64);
  for (int idx = tid; idx < K * M; idx += 128) { const int k = idx / M, m = idx % M; *reinterpret_cast<float*>(As + mn_off(m, k, lboA, sboA)) = A[idx]; }
  for (int idx = tid; idx < K * N; idx += 128) { const int k = idx / N, n = idx % N; *reinterpret_cast<float*>(Bs + mn_off(n, k, lboB, sboB)) = B[idx]; }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // D=f32 (bit 4), A=B=tf32 (2<<7, 2<<10), a_major=MN (bit 15), b_major=MN (bit 16), N>>3 << 17, M>>4 << 24
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
                         ((uint32_t)(M >> 4) << 24);
  if (tid == 0) {
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint64_t ad = swap_lbo_sbo ? mn_desc(s2u(As) + ks * 2 * sboA, sboA, lboA, 1) : mn_desc(s2u(As) + ks * 2 * sboA, lboA, sboA, 1);
      const uint64_t bd = swap_lbo_sbo ? mn_desc(s2u(Bs) + ks * 2 * sboB, sboB, lboB, 1) : mn_desc(s2u(Bs) + ks * 2 * sboB, lboB, sboB, 1);
      umma_tf32(tmem_base, ad, bd, idesc, ks != 0);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  float v0[32], v1[32];
  const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
  tmem_ld32(trow, v0);
  tmem_ld32(trow + 32, v1);
  const int m = warp * 32 + lane;
  for (int j = 0; j < 32; ++j) { D[m * N + j] = v0[j]; D[m * N + 32 + j] = v1[j]; }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

int main() {
  std::vector<float> A(K * M), B(K * N), D(M * N), R(M * N, 0.f);
  srand(1);
  for (auto& x : A) x = (float)((rand() % 33) - 16) / 8.f;
  for (auto& x : B) x = (float)((rand() % 33) - 16) / 8.f;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float acc = 0.f;
      for (int k = 0; k < K; ++k) acc += A[k * M + m] * B[k * N + n];
      R[m * N + n] = acc;
    }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const int smem = 49152 + 64 + 1024;
  cudaFuncSetAttribute(mn_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(dD, 0, D.size() * 4);
    mn_test_kernel<<<1, 128, smem>>>(dA, dB, dD, mode);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0; double maxd = 0;
    for (int i = 0; i < M * N; ++i) { const double d = fabs((double)D[i] - R[i]); if (d > 1e-4) ++bad; if (d > maxd) maxd = d; }
    printf("mode %d (lbo/sbo %s): err=%s mismatches=%d of %d, max diff %.4f, D[0]=%.3f R[0]=%.3f D[1]=%.3f R[1]=%.3f D[64]=%.3f R[64]=%.3f\n",
           mode, mode ? "swapped" : "as derived", cudaGetErrorString(e), bad, M * N, maxd, D[0], R[0], D[1], R[1], D[64], R[64]);
  }
  return 0;
}
