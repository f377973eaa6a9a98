// Row-sharded BPR-MF (BASELINE config 5): device-side pieces of one step; the NCCL collectives between them are
// issued by the caller (yelprecommendation_b200/trainers/sharded_mf_trainer.py) on the same stream.
//   yr_shard_gather_rows — owner fills the batch's rows it holds (zeros elsewhere) -> SUM all-reduce = exact gather
//   yr_bpr_rows_grad     — forward + BPR loss + per-triple gradient rows for this rank's slice of the batch
//   yr_shard_accumulate  — owner sums the gradient rows of its ids into scratch (duplicates summed, rows listed)
//   yr_shard_step        — owner steps its shard once: listed rows (plain SGD) or every row (dense-semantics Adam / L2)
// Same arithmetic as the single-GPU fused kernel (mf.cu): reference trainers/mf_trainer.py:104-114.
#include "common.cuh"

namespace yr {

template <int VPL>
__global__ void __launch_bounds__(256)
shard_gather_kernel(const float* __restrict__ T, int64_t row0, int64_t row1, int64_t n_rows_global,
                    const int64_t* __restrict__ ids, int64_t n, float* __restrict__ out, int64_t out_ld, int32_t* err) {
  constexpr int D = VPL * 32;
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n; j += nw) {
    const int64_t id = ids[j];
    Row<VPL> r;
#pragma unroll
    for (int q = 0; q < VPL; ++q) r.x[q] = 0.f;
    if (id < 0 || id >= n_rows_global) { if (lane == 0 && err) atomicExch(err, 1); }
    else if (id >= row0 && id < row1) r = ld_row<VPL>(T + (id - row0) * D, lane);
    st_row<VPL>(out + j * out_ld, lane, r);
  }
}

template <int VPL>
__global__ void __launch_bounds__(256)
bpr_rows_grad_kernel(const float* __restrict__ R, int64_t B, int64_t b0, int64_t b1, float* __restrict__ G,
                     double* loss_acc) {
  constexpr int D = VPL * 32;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float inv_b = 1.f / (float)B;
  __shared__ double s_part[8];
  double wl = 0.0;
  for (int64_t b = b0 + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); b < b1; b += nw) {
    const float* rb = R + b * 3 * D;
    const Row<VPL> ur = ld_row<VPL>(rb, lane), pr = ld_row<VPL>(rb + D, lane), nr = ld_row<VPL>(rb + 2 * D, lane);
    const float x = warp_sum(dot_partial<VPL>(ur, pr)) - warp_sum(dot_partial<VPL>(ur, nr));
    wl += (double)neg_logsigmoid(x);
    const float g = neg_logsigmoid_grad(x) * inv_b;
    Row<VPL> gu, gp, gn;
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
      gu.x[q] = __fsub_rn(__fmul_rn(g, pr.x[q]), __fmul_rn(g, nr.x[q]));
      gp.x[q] = g * ur.x[q];
      gn.x[q] = -gp.x[q];
    }
    float* gb = G + b * 3 * D;
    st_row<VPL>(gb, lane, gu);
    st_row<VPL>(gb + D, lane, gp);
    st_row<VPL>(gb + 2 * D, lane, gn);
  }
  if (lane == 0) s_part[wib] = wl;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += s_part[i];
    if (t != 0.0) atomicAdd(loss_acc, t);
  }
}

// accumulate the gradient rows of owned ids into the shard's scratch; first toucher lists the row
template <int VPL>
__global__ void __launch_bounds__(256)
shard_scatter_kernel(yr_shard_state st, const int64_t* __restrict__ ids, int64_t n, const float* __restrict__ G,
                     int64_t g_ld, bool list_rows) {
  constexpr int D = VPL * 32;
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n; j += nw) {
    const int64_t id = ids[j];
    if (id < st.row0 || id >= st.row1) continue;
    const int64_t r = id - st.row0;
    const Row<VPL> g = ld_row<VPL>(G + j * g_ld, lane);
    red_row<VPL>(st.gscratch + r * D, lane, g);
    if (lane == 0) {
      if (list_rows) {
        if (atomicExch(st.flags + r, 1) == 0) st.rows_list[atomicAdd(st.counters, 1)] = (int32_t)r;
      } else {
        st.flags[r] = 1;
      }
    }
  }
}

template <int VPL>
__global__ void __launch_bounds__(256)
shard_update_kernel(yr_shard_state st, yr_opt opt, bool dense) {
  constexpr int D = VPL * 32;
  const int lane = threadIdx.x & 31;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  OptScalars os;
  opt_scalars_for_step(os, opt, opt.step);
  const int64_t n_items = dense ? (st.row1 - st.row0) : (int64_t)__ldcg(st.counters);
  for (int64_t i = w0; i < n_items; i += nw) {
    const int64_t r = dense ? i : (int64_t)st.rows_list[i];
    const bool touched = dense ? (__ldcg(st.flags + r) != 0) : true;
    float* prow = st.T + r * D;
    float* grow = st.gscratch + r * D;
    Row<VPL> pv = ld_row<VPL>(prow, lane), gv, mv, vv;
    if (touched) gv = ld_row<VPL>(grow, lane);
    else {
#pragma unroll
      for (int q = 0; q < VPL; ++q) gv.x[q] = 0.f;
    }
    if (opt.kind != YR_OPT_SGD) {
      mv = ld_row<VPL>(st.m + r * D, lane);
      vv = ld_row<VPL>(st.v + r * D, lane);
#pragma unroll
      for (int q = 0; q < VPL; ++q) opt_update(os, pv.x[q], gv.x[q], mv.x[q], vv.x[q]);
      st_row<VPL>(st.m + r * D, lane, mv);
      st_row<VPL>(st.v + r * D, lane, vv);
    } else {
#pragma unroll
      for (int q = 0; q < VPL; ++q) { float m = 0.f, v = 0.f; opt_update(os, pv.x[q], gv.x[q], m, v); }
    }
    st_row<VPL>(prow, lane, pv);
    if (touched) {
      Row<VPL> z;
#pragma unroll
      for (int q = 0; q < VPL; ++q) z.x[q] = 0.f;
      st_row<VPL>(grow, lane, z);
      if (lane == 0) st.flags[r] = 0;
    }
  }
}

__global__ void shard_reset_kernel(int32_t* counters) { counters[0] = 0; }

static unsigned warp_grid(int64_t n_warps, int per_block_warps = 8) {
  int64_t blocks = (n_warps + per_block_warps - 1) / per_block_warps;
  const int64_t cap = (int64_t)yr_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace yr

using namespace yr;

extern "C" int yr_shard_gather_rows(const float* T_local, int64_t row0, int64_t row1, int64_t n_rows_global, int d,
                                    const int64_t* ids, int64_t n, float* out, int64_t out_ld, int32_t* err,
                                    yr_stream stream) {
  if (!T_local || !ids || !out || n < 0 || row1 < row0 || out_ld < d) return YR_ERR_BAD_ARG;
  if (n == 0) return YR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned g = warp_grid(n);
  switch (dim_vpl(d)) {
    case 1: shard_gather_kernel<1><<<g, 256, 0, s>>>(T_local, row0, row1, n_rows_global, ids, n, out, out_ld, err); break;
    case 2: shard_gather_kernel<2><<<g, 256, 0, s>>>(T_local, row0, row1, n_rows_global, ids, n, out, out_ld, err); break;
    case 4: shard_gather_kernel<4><<<g, 256, 0, s>>>(T_local, row0, row1, n_rows_global, ids, n, out, out_ld, err); break;
    case 8: shard_gather_kernel<8><<<g, 256, 0, s>>>(T_local, row0, row1, n_rows_global, ids, n, out, out_ld, err); break;
    default: return YR_ERR_BAD_DIM;
  }
  YR_CHECK_LAUNCH();
  return YR_OK;
}

extern "C" int yr_bpr_rows_grad(const float* R, int d, int64_t B, int64_t b0, int64_t b1, float* G, double* loss_acc,
                                yr_stream stream) {
  if (!R || !G || !loss_acc || B <= 0 || b0 < 0 || b1 > B || b1 < b0) return YR_ERR_BAD_ARG;
  if (b1 == b0) return YR_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned g = warp_grid(b1 - b0);
  switch (dim_vpl(d)) {
    case 1: bpr_rows_grad_kernel<1><<<g, 256, 0, s>>>(R, B, b0, b1, G, loss_acc); break;
    case 2: bpr_rows_grad_kernel<2><<<g, 256, 0, s>>>(R, B, b0, b1, G, loss_acc); break;
    case 4: bpr_rows_grad_kernel<4><<<g, 256, 0, s>>>(R, B, b0, b1, G, loss_acc); break;
    case 8: bpr_rows_grad_kernel<8><<<g, 256, 0, s>>>(R, B, b0, b1, G, loss_acc); break;
    case 16: bpr_rows_grad_kernel<16><<<g, 256, 0, s>>>(R, B, b0, b1, G, loss_acc); break;      // NGCF: d * (layers + 1) = 512
    case 32: bpr_rows_grad_kernel<32><<<g, 256, 0, s>>>(R, B, b0, b1, G, loss_acc); break;
    default: return YR_ERR_BAD_DIM;
  }
  YR_CHECK_LAUNCH();
  return YR_OK;
}

extern "C" int yr_shard_accumulate(const yr_shard_state* st, const yr_opt* opt, const int64_t* ids, int64_t n,
                                   const float* G, int64_t g_ld, yr_stream stream) {
  if (!st || !opt || !ids || !G || n < 0 || !st->T || !st->gscratch || !st->flags || !st->rows_list || !st->counters)
    return YR_ERR_BAD_ARG;
  if (g_ld < st->d) return YR_ERR_BAD_ARG;
  if (st->row1 <= st->row0 || n == 0) return YR_OK;
  const bool dense = (opt->kind != YR_OPT_SGD) || (opt->weight_decay != 0.0);
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned g = warp_grid(n);
  switch (dim_vpl(st->d)) {
    case 1: shard_scatter_kernel<1><<<g, 256, 0, s>>>(*st, ids, n, G, g_ld, !dense); break;
    case 2: shard_scatter_kernel<2><<<g, 256, 0, s>>>(*st, ids, n, G, g_ld, !dense); break;
    case 4: shard_scatter_kernel<4><<<g, 256, 0, s>>>(*st, ids, n, G, g_ld, !dense); break;
    case 8: shard_scatter_kernel<8><<<g, 256, 0, s>>>(*st, ids, n, G, g_ld, !dense); break;
    default: return YR_ERR_BAD_DIM;
  }
  YR_CHECK_LAUNCH();
  return YR_OK;
}

extern "C" int yr_shard_step(const yr_shard_state* st, const yr_opt* opt, int64_t max_rows, yr_stream stream) {
  if (!st || !opt || !st->T || !st->gscratch || !st->flags || !st->rows_list || !st->counters) return YR_ERR_BAD_ARG;
  if (opt->kind < YR_OPT_SGD || opt->kind > YR_OPT_ADAMW) return YR_ERR_BAD_OPT;
  if (opt->kind != YR_OPT_SGD && (!st->m || !st->v)) return YR_ERR_BAD_ARG;
  if (st->row1 <= st->row0) return YR_OK;
  const bool dense = (opt->kind != YR_OPT_SGD) || (opt->weight_decay != 0.0);
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned g = warp_grid(dense ? (st->row1 - st->row0) : (max_rows > 0 ? max_rows : 1));
  switch (dim_vpl(st->d)) {
    case 1: shard_update_kernel<1><<<g, 256, 0, s>>>(*st, *opt, dense); break;
    case 2: shard_update_kernel<2><<<g, 256, 0, s>>>(*st, *opt, dense); break;
    case 4: shard_update_kernel<4><<<g, 256, 0, s>>>(*st, *opt, dense); break;
    case 8: shard_update_kernel<8><<<g, 256, 0, s>>>(*st, *opt, dense); break;
    default: return YR_ERR_BAD_DIM;
  }
  YR_CHECK_LAUNCH();
  if (!dense) { shard_reset_kernel<<<1, 1, 0, s>>>(st->counters); YR_CHECK_LAUNCH(); }
  return YR_OK;
}
