// Shared device helpers for the sm_100a kernels. No torch headers anywhere in csrc/.
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include "../../include/yelprec_b200.h"

namespace cg = cooperative_groups;

#define YR_CHECK_LAUNCH()                                   \
  do {                                                      \
    cudaError_t e__ = cudaGetLastError();                   \
    if (e__ != cudaSuccess) return (int)e__;                \
  } while (0)

#define YR_CUDA(call)                                       \
  do {                                                      \
    cudaError_t e__ = (call);                               \
    if (e__ != cudaSuccess) return (int)e__;                \
  } while (0)

// library-internal launchers (eval.cu), shared with eval_tc.cu; not part of the C ABI
int yr_eval_exact_launch(const float* Uemb, int64_t nU, const float* Vt, int64_t ldt, int64_t nI, int d,
                         const int64_t* eval_uid, int64_t n_eval, const int32_t* mask_ptr,
                         const int32_t* mask_idx, const int32_t* act_ptr, const int32_t* act_idx,
                         const int32_t* act_nuniq, const double* inv_log2, int K, int64_t* topk_out,
                         float* topk_score, double* user_metrics, int32_t* err, const int32_t* row_list,
                         const int32_t* n_rows_dev, yr_stream stream, void* ws = nullptr, size_t ws_bytes = 0);
int yr_eval_reduce_launch(const double* user_metrics, const int32_t* act_ptr, const int32_t* act_nuniq,
                          int64_t n_eval, double* metric_sums, yr_stream stream);

int yr_ngcf_dense_fwd_tc_launch_d(int d, const float* E, const float* LE, const float* W1, const float* W2, float slope,
                                  int64_t n, float* Eout, cudaStream_t s, const int32_t* row_list = nullptr,
                                  const int32_t* row_count = nullptr, int64_t row_cap = 0, int reserve_sms = 0);
int yr_ngcf_dense_bwd_tc_launch(int d, const float* E, const float* LE, const float* En, const float* Gn, const float* W1,
                                const float* W2, float slope, int64_t n, float* G, float* T, float* ws, int* n_parts,
                                cudaStream_t s, const int32_t* row_list = nullptr, const int32_t* row_count = nullptr,
                                int64_t row_cap = 0, int reserve_sms = 0);
// Y = A X on the rows whose flag is set (all rows if row_flag == NULL); other rows of Y are left untouched
int yr_spmm_csr_rows(const yr_csr* A, int d, const float* X, float* Y, const int32_t* row_flag, cudaStream_t s);

namespace yr {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// Grid barrier with the calling warp re-converged on both sides. cooperative_groups' grid sync lets thread 0 of the CTA
// poll the barrier word between two CTA barriers; when the compiler leaves warp 0 diverged after an `if (threadIdx.x == 0)`
// block in front of it (seen for the d = 32 instantiation of the BPR-MF trainer on sm_100a), lanes 1..31 of that warp ran
// into the next phase before the grid barrier had completed and read rows other CTAs were still writing
// (scripts/probe_mf, notes/README.md). __syncwarp() after the barrier holds them until lane 0 is through.
__device__ __forceinline__ void grid_sync(cg::grid_group& grid) {
  __syncwarp();
  grid.sync();
  __syncwarp();
}

// Per-lane slice of an embedding row of d = 32 * VPL floats. VPL = 1 / 2: lane owns VPL contiguous floats; VPL >= 4: lane
// owns float4 number j * 32 + lane for j < VPL / 4 (every warp access is one contiguous 512-byte run).
template <int VPL> struct Row { float x[VPL]; };

template <int VPL>
__device__ __forceinline__ Row<VPL> ld_row(const float* __restrict__ row, int lane) {
  Row<VPL> r;
  if constexpr (VPL == 1) {
    r.x[0] = row[lane];
  } else if constexpr (VPL == 2) {
    float2 t = reinterpret_cast<const float2*>(row)[lane];
    r.x[0] = t.x; r.x[1] = t.y;
  } else {
#pragma unroll
    for (int j = 0; j < VPL / 4; ++j) {
      float4 t = reinterpret_cast<const float4*>(row)[j * 32 + lane];
      r.x[4 * j + 0] = t.x; r.x[4 * j + 1] = t.y; r.x[4 * j + 2] = t.z; r.x[4 * j + 3] = t.w;
    }
  }
  return r;
}

// same slices through L2 only (ld.global.cg): rows other SMs rewrite every step inside one persistent kernel
template <int VPL>
__device__ __forceinline__ Row<VPL> ld_row_cg(const float* __restrict__ row, int lane) {
  Row<VPL> r;
  if constexpr (VPL == 1) {
    r.x[0] = __ldcg(row + lane);
  } else if constexpr (VPL == 2) {
    float2 t = __ldcg(reinterpret_cast<const float2*>(row) + lane);
    r.x[0] = t.x; r.x[1] = t.y;
  } else {
#pragma unroll
    for (int j = 0; j < VPL / 4; ++j) {
      float4 t = __ldcg(reinterpret_cast<const float4*>(row) + j * 32 + lane);
      r.x[4 * j + 0] = t.x; r.x[4 * j + 1] = t.y; r.x[4 * j + 2] = t.z; r.x[4 * j + 3] = t.w;
    }
  }
  return r;
}

template <int VPL>
__device__ __forceinline__ void st_row(float* __restrict__ row, int lane, const Row<VPL>& r) {
  if constexpr (VPL == 1) {
    row[lane] = r.x[0];
  } else if constexpr (VPL == 2) {
    reinterpret_cast<float2*>(row)[lane] = make_float2(r.x[0], r.x[1]);
  } else {
#pragma unroll
    for (int j = 0; j < VPL / 4; ++j)
      reinterpret_cast<float4*>(row)[j * 32 + lane] =
          make_float4(r.x[4 * j + 0], r.x[4 * j + 1], r.x[4 * j + 2], r.x[4 * j + 3]);
  }
}

// Vector reduction into global memory (sm_90+ has native float2/float4 atomics: RED.E.ADD.F32x2/x4).
template <int VPL>
__device__ __forceinline__ void red_row(float* __restrict__ row, int lane, const Row<VPL>& r) {
  if constexpr (VPL == 1) {
    atomicAdd(row + lane, r.x[0]);
  } else if constexpr (VPL == 2) {
    atomicAdd(reinterpret_cast<float2*>(row) + lane, make_float2(r.x[0], r.x[1]));
  } else {
#pragma unroll
    for (int j = 0; j < VPL / 4; ++j)
      atomicAdd(reinterpret_cast<float4*>(row) + j * 32 + lane,
                make_float4(r.x[4 * j + 0], r.x[4 * j + 1], r.x[4 * j + 2], r.x[4 * j + 3]));
  }
}

template <int VPL>
__device__ __forceinline__ float dot_partial(const Row<VPL>& a, const Row<VPL>& b) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < VPL; ++j) s = fmaf(a.x[j], b.x[j], s);
  return s;
}

// exp / log1p as CORRECTLY ROUNDED fp32 (double-precision evaluation rounded once): the one definition the device and the
// CPU oracle can share bit for bit — libm's expf, CUDA's expf and ATen's Sleef kernels all differ from each other in
// the last ulp, and under Adam a 1-ulp change of the loss gradient is amplified on elements with |grad| ~ eps
// (DESIGN.md section 4). Every implementation above is within 1 ulp of this value.
__device__ __forceinline__ float exp_cr(float x) { return (float)exp((double)x); }
__device__ __forceinline__ float log1p_cr(float x) { return (float)log1p((double)x); }

// -logsigmoid(x) with torch's formulation: logsigmoid(x) = min(x,0) - log1p(exp(-|x|))
// (aten/src/ATen/native/cpu/Activation.cpp log_sigmoid_cpu_kernel); loss.py:25-27.
__device__ __forceinline__ float neg_logsigmoid(float x) {
  return __fsub_rn(log1p_cr(exp_cr(-fabsf(x))), fminf(x, 0.f));
}
// d(-logsigmoid(x))/dx = -sigmoid(-x), in torch's backward formulation.
__device__ __forceinline__ float neg_logsigmoid_grad(float x) {
  const float z = exp_cr(-fabsf(x));
  const float q = __fdiv_rn(z, __fadd_rn(1.f, z));
  const float s = (x < 0.f) ? __fsub_rn(1.f, q) : q;   // sigmoid(-x)
  return -s;
}

// The same two functions on the fp32 pipe (expf / log1pf, <= 1 ulp from the definitions above): for the register-resident SGD
// path of the BPR-MF trainer, whose REDs are not bit-reproducible anyway and whose step (5 us) would pay ~0.9 us for the
// double-precision versions.
__device__ __forceinline__ float neg_logsigmoid_fast(float x) { return log1pf(expf(-fabsf(x))) - fminf(x, 0.f); }
__device__ __forceinline__ float neg_logsigmoid_grad_fast(float x) {
  const float z = expf(-fabsf(x));
  const float q = __fdiv_rn(z, 1.f + z);
  return -((x < 0.f) ? 1.f - q : q);
}

// torch.optim single-tensor update of one element (torch/optim/{sgd,adam,adamw}.py op order).
struct OptScalars {
  int kind;
  float lr, wd;            // SGD: lr, coupled wd; AdamW: lr*wd decoupled
  float w_lerp;            // (float)(1 - beta1)
  float beta2, omb2;       // (float)beta2, (float)(1 - beta2)
  float eps;
  float step_size;         // (float)(lr / (1 - beta1^t))
  float bc2_sqrt;          // (float)sqrt(1 - beta2^t)
  float decay;             // AdamW: (float)(1 - lr*wd)
};

__device__ __forceinline__ void opt_scalars_for_step(OptScalars& s, const yr_opt& o, int t) {
  s.kind = o.kind;
  s.lr = (float)o.lr;
  s.wd = (float)o.weight_decay;
  s.w_lerp = (float)(1.0 - o.beta1);
  s.beta2 = (float)o.beta2;
  s.omb2 = (float)(1.0 - o.beta2);
  s.eps = (float)o.eps;
  const double bc1 = 1.0 - pow(o.beta1, (double)t);
  const double bc2 = 1.0 - pow(o.beta2, (double)t);
  s.step_size = (float)(o.lr / bc1);
  s.bc2_sqrt = (float)sqrt(bc2);
  s.decay = (float)(1.0 - o.lr * o.weight_decay);
}

// Every operation is an explicit IEEE round-to-nearest intrinsic (no reliance on the compiler's fma contraction), so the
// C oracle (oracle/yr_oracle.c, orc_dense_opt_step) reproduces it bit for bit.
__device__ __forceinline__ void opt_update(const OptScalars& s, float& p, float g, float& m, float& v) {
  if (s.kind == YR_OPT_SGD) {
    if (s.wd != 0.f) g = fmaf(p, s.wd, g);            // grad.add(param, alpha=wd)
    p = fmaf(g, -s.lr, p);                            // param.add_(grad, alpha=-lr)
    return;
  }
  if (s.kind == YR_OPT_ADAMW) {
    p = __fmul_rn(p, s.decay);                        // param.mul_(1 - lr*wd)
  } else if (s.wd != 0.f) {
    g = fmaf(p, s.wd, g);                             // grad.add(param, alpha=wd)
  }
  m = fmaf(s.w_lerp, __fsub_rn(g, m), m);             // exp_avg.lerp_(grad, 1-beta1)
  v = __fmul_rn(v, s.beta2);                          // exp_avg_sq.mul_(beta2)
  v = fmaf(__fmul_rn(s.omb2, g), g, v);               //   .addcmul_(grad, grad, value=1-beta2)
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), s.bc2_sqrt), s.eps);   // (sqrt(v)/sqrt(bc2)).add_(eps)
  p = __fadd_rn(p, __fdiv_rn(__fmul_rn(-s.step_size, m), denom));              // param.addcdiv_(m, denom, value=-step_size)
}

inline int yr_sm_count() {
  static int sms[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!sms[dev]) {
    cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    if (sms[dev] <= 0) sms[dev] = 148;
  }
  return sms[dev];
}

// cudaFuncSetAttribute is per DEVICE: remember which devices have been opted in (a process may drive several GPUs)
struct AttrOnce {
  bool done[64] = {};
  template <typename K> int set(K kernel, int bytes) {
    int dev = 0;
    YR_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !done[dev]) {
      YR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
      if (dev >= 0 && dev < 64) done[dev] = true;
    }
    return YR_OK;
  }
};

inline int yr_csr_ok(const yr_csr* A) {
  if (!A || !A->rowptr || !A->col || !A->val || A->n_rows < 0 || A->n_chunks < 0) return YR_ERR_BAD_ARG;
  if (A->n_chunks > 0 && !A->chunk_desc) return YR_ERR_BAD_ARG;
  if (A->n_split_rows > 0 && (!A->split_row || !A->split_ptr || !A->partials || !A->split_count)) return YR_ERR_BAD_ARG;
  return YR_OK;
}

inline int dim_vpl(int d) {
  switch (d) {
    case 32: return 1; case 64: return 2; case 128: return 4; case 256: return 8;
    case 512: return 16; case 1024: return 32;     // BPR-MF only (the reference's mf_sweep_config.yaml goes to 1,024)
    default: return 0;
  }
}

}  // namespace yr
