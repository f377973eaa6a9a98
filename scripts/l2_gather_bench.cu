// Micro-benchmark: the measured L2 -> SM ceiling for random row gathers out of an L2-resident table.
// (the SpMM was compared against a B300 document constant; this measures the ceiling on the
// box the bench runs on.)  Table: R rows of D floats (default 69,716 x 64 = 17.85 MB, the NGCF operand at Yelp
// shape); G gathers per launch (default 3,122,812 = nnz of the Laplacian), grouped LEN per output row like the SpMM.
// Variants:
//   ldg      : half-warp per output row, one float4 per lane, U gathers in flight per group (the SpMM's shape)
//   ldg_sum  : same loads, plain adds instead of the fma chain, no output dependency (pure load ceiling)
//   bulk     : cp.async.bulk (UBLKCP) of whole rows into a shared-memory ring, consumer warps add from shared memory
// Output: one line per variant / configuration: microseconds per launch and TB/s of gathered bytes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o scripts/bin/l2_gather_bench scripts/l2_gather_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static inline uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

template <int D, int U, bool FMA>
__global__ void __launch_bounds__(256) gather_ldg(const float4* __restrict__ X, const int32_t* __restrict__ idx,
                                                  const float* __restrict__ val, float4* __restrict__ Y, int n_rows, int len) {
  constexpr int LPR = D / 4;                     // lanes per row
  constexpr int GPW = 32 / LPR;                  // groups per warp
  const int lane = threadIdx.x & 31, sub = lane / LPR, sl = lane % LPR;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  for (int r = gw * GPW + sub; r < n_rows; r += nw * GPW) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int s = r * len;
    for (int j0 = 0; j0 < len; j0 += U) {
      float4 x[U]; float a[U];
#pragma unroll
      for (int q = 0; q < U; ++q) {
        const int j = j0 + q;
        const int c = (j < len) ? __ldg(idx + s + j) : 0;
        a[q] = (j < len) ? __ldg(val + s + j) : 0.f;
        x[q] = __ldg(X + (int64_t)c * LPR + sl);
      }
#pragma unroll
      for (int q = 0; q < U; ++q) {
        if (FMA) { acc.x = fmaf(a[q], x[q].x, acc.x); acc.y = fmaf(a[q], x[q].y, acc.y); acc.z = fmaf(a[q], x[q].z, acc.z); acc.w = fmaf(a[q], x[q].w, acc.w); }
        else { acc.x += x[q].x; acc.y += x[q].y; acc.z += x[q].z; acc.w += x[q].w; }
      }
    }
    Y[(int64_t)r * LPR + sl] = acc;
  }
}

__device__ __forceinline__ uint32_t s2u(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s2u(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s2u(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s2u(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra Dn;\n\tbra W;\n\tDn:\n\t}" ::"r"(s2u(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(s2u(dst)), "l"(src), "r"(bytes), "r"(s2u(bar)) : "memory");
}

// bulk variant: CTA = 1 producer warp + CW consumer warps. A stage holds 32 rows (one per producer lane).
// Consumers: consumer warp w handles output rows; to keep the benchmark simple every stage is consumed by ONE consumer
// warp (round robin), which adds the 32 rows of the stage into two half-warp accumulators and writes them out.
template <int D, int STAGES, int CW>
__global__ void __launch_bounds__(32 * (CW + 1)) gather_bulk(const float* __restrict__ X, const int32_t* __restrict__ idx,
                                                             float4* __restrict__ Y, int n_stage_total) {
  constexpr int ROWB = D * 4;
  extern __shared__ __align__(128) unsigned char smem[];
  float* ring = reinterpret_cast<float*>(smem);                           // STAGES x 32 x D floats
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * 32 * ROWB);
  uint64_t* empty = full + STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // stages of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int n_mine = (n_stage_total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (warp == CW) {                                // producer
    for (int t = 0; t < n_mine; ++t) {
      const int s = t % STAGES, ph = (t / STAGES) & 1;
      if (t >= STAGES) mbar_wait(empty + s, ph ^ 1);
      const int64_t g = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * 32 + lane;
      const int c = __ldg(idx + g);
      if (lane == 0) mbar_expect_tx(full + s, 32 * ROWB);
      __syncwarp();
      bulk_g2s(ring + ((size_t)s * 32 + lane) * D, X + (int64_t)c * D, ROWB, full + s);
    }
  } else {
    constexpr int LPR = D / 4, GPW = 32 / LPR;
    const int sub = lane / LPR, sl = lane % LPR;
    for (int t = warp; t < n_mine; t += CW) {
      const int s = t % STAGES, ph = (t / STAGES) & 1;
      mbar_wait(full + s, ph);
      const float4* st4 = reinterpret_cast<const float4*>(ring + (size_t)s * 32 * D);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = sub; r < 32; r += GPW) {
        const float4 x = st4[r * LPR + sl];
        acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);
      const int64_t g = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * GPW + sub;
      Y[g * LPR + sl] = acc;
    }
  }
}

template <typename F> static float time_it(F f, int reps) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  float ms = 0; CK(cudaEventElapsedTime(&ms, a, b));
  CK(cudaGetLastError());
  return ms * 1000.f / reps;
}

int main(int argc, char** argv) {
  const int D = 64;
  int R = 69716, len = 45; int64_t G = 3122812; double zipf = 0.0;
  for (int i = 1; i + 1 < argc; i += 2) {
    if (!strcmp(argv[i], "--rows")) R = atoi(argv[i + 1]);
    if (!strcmp(argv[i], "--len")) len = atoi(argv[i + 1]);
    if (!strcmp(argv[i], "--gathers")) G = atoll(argv[i + 1]);
  }
  const int n_rows = (int)(G / len); G = (int64_t)n_rows * len;
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  printf("# table %d rows x %d floats = %.2f MB, %lld gathers of %d B per launch (%.1f MB gathered), %d output rows, %d SMs\n",
         R, D, R * D * 4 / 1e6, (long long)G, D * 4, G * D * 4 / 1e6, n_rows, sms);
  (void)zipf;
  std::vector<int32_t> hidx((size_t)G + 64); std::vector<float> hval((size_t)G + 64, 0.01f);
  for (auto& v : hidx) v = (int32_t)(rnd() % (uint64_t)R);
  float *X, *val; int32_t* idx; float4* Y;
  CK(cudaMalloc(&X, (size_t)R * D * 4)); CK(cudaMalloc(&val, hval.size() * 4)); CK(cudaMalloc(&idx, hidx.size() * 4));
  CK(cudaMalloc(&Y, (size_t)(G / 16 + 1024) * D * 4));
  CK(cudaMemset(X, 0, (size_t)R * D * 4));
  CK(cudaMemcpy(idx, hidx.data(), hidx.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(val, hval.data(), hval.size() * 4, cudaMemcpyHostToDevice));
  const double bytes = (double)G * D * 4;
  const int reps = 20;
  auto report = [&](const char* name, int cfg_a, int cfg_b, float us) {
    printf("%-10s cfg=(%d,%d)  %8.2f us  %6.2f TB/s\n", name, cfg_a, cfg_b, us, bytes / us * 1e-6);
    fflush(stdout);
  };
  const float4* X4 = reinterpret_cast<const float4*>(X);
  for (int cps = 2; cps <= 8; cps += (cps < 4 ? 1 : 2)) {          // CTAs per SM worth of grid (grid-stride)
    const int grid = sms * cps;
    report("ldg_fma_u8", cps, 8, time_it([&] { gather_ldg<D, 8, true><<<grid, 256>>>(X4, idx, val, Y, n_rows, len); }, reps));
    report("ldg_sum_u8", cps, 8, time_it([&] { gather_ldg<D, 8, false><<<grid, 256>>>(X4, idx, val, Y, n_rows, len); }, reps));
    report("ldg_sum_u16", cps, 16, time_it([&] { gather_ldg<D, 16, false><<<grid, 256>>>(X4, idx, val, Y, n_rows, len); }, reps));
  }
  {  // one CTA per ~8 output-row pairs, non-persistent (hardware scheduler balances), like launch_spmm
    const int grid = (n_rows + 15) / 16;
    report("ldg_fma_np", grid, 8, time_it([&] { gather_ldg<D, 8, true><<<grid, 256>>>(X4, idx, val, Y, n_rows, len); }, reps));
  }
  const int n_stage_total = (int)(G / 32);
  {
    constexpr int ST = 16, CW = 4;
    const size_t sm = (size_t)ST * 32 * D * 4 + 2 * ST * 8;
    CK(cudaFuncSetAttribute(gather_bulk<D, ST, CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    for (int cps = 1; cps <= 1; ++cps)
      report("bulk_s16", cps, CW, time_it([&] { gather_bulk<D, ST, CW><<<sms * cps, 32 * (CW + 1), sm>>>(X, idx, Y, n_stage_total); }, reps));
  }
  {
    constexpr int ST = 8, CW = 4;
    const size_t sm = (size_t)ST * 32 * D * 4 + 2 * ST * 8;
    CK(cudaFuncSetAttribute(gather_bulk<D, ST, CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    for (int cps = 1; cps <= 3; ++cps)
      report("bulk_s8", cps, CW, time_it([&] { gather_bulk<D, ST, CW><<<sms * cps, 32 * (CW + 1), sm>>>(X, idx, Y, n_stage_total); }, reps));
  }
  {
    constexpr int ST = 24, CW = 8;
    const size_t sm = (size_t)ST * 32 * D * 4 + 2 * ST * 8;
    CK(cudaFuncSetAttribute(gather_bulk<D, ST, CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    report("bulk_s24", 1, CW, time_it([&] { gather_bulk<D, ST, CW><<<sms, 32 * (CW + 1), sm>>>(X, idx, Y, n_stage_total); }, reps));
  }
  // sequential (coalesced) read of the same number of bytes out of the same L2-resident table: the streaming L2 ceiling
  {
    std::vector<int32_t> seq(hidx.size());
    for (size_t i = 0; i < seq.size(); ++i) seq[i] = (int32_t)(i % (size_t)R);
    CK(cudaMemcpy(idx, seq.data(), seq.size() * 4, cudaMemcpyHostToDevice));
    report("seq_sum_u8", 4, 8, time_it([&] { gather_ldg<D, 8, false><<<sms * 4, 256>>>(X4, idx, val, Y, n_rows, len); }, reps));
    report("seq_sum_u16", 6, 16, time_it([&] { gather_ldg<D, 16, false><<<sms * 6, 256>>>(X4, idx, val, Y, n_rows, len); }, reps));
  }
  return 0;
}
