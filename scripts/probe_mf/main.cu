// temporary probe: one dense-semantics (SGD + weight decay) step of yr_bpr_mf_train at d = D_PROBE, then counts leftover
// gradient rows / flags (both must be zero).
#include <cstdio>
#include <vector>
#include <random>
#include "mf_probe.cu"
int main() {
  const int d = D_PROBE, nU = 3000, nI = 2500, B = 1024;
  std::mt19937_64 rng(d);
  std::vector<float> U((size_t)nU * d), V((size_t)nI * d);
  std::uniform_real_distribution<float> ud(-0.05f, 0.05f);
  for (auto& x : U) x = ud(rng);
  for (auto& x : V) x = ud(rng);
  std::vector<int64_t> u(B), p(B), n(B);
  for (int i = 0; i < B; ++i) { u[i] = rng() % nU; p[i] = rng() % nI; n[i] = rng() % nI; }
  float *dU, *dV, *gU, *gV; int32_t *fU, *fV, *rows, *cnt, *err; int64_t *du, *dp, *dn; double* ls;
  cudaMalloc(&dU, U.size() * 4); cudaMalloc(&dV, V.size() * 4); cudaMalloc(&gU, U.size() * 4); cudaMalloc(&gV, V.size() * 4);
  cudaMalloc(&fU, nU * 4); cudaMalloc(&fV, nI * 4); cudaMalloc(&rows, 3 * B * 4); cudaMalloc(&cnt, 64); cudaMalloc(&err, 4);
  cudaMalloc(&du, B * 8); cudaMalloc(&dp, B * 8); cudaMalloc(&dn, B * 8); cudaMalloc(&ls, 8);
  int bad_total = 0;
  for (int rep = 0; rep < 20; ++rep) {
    cudaMemcpy(dU, U.data(), U.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dV, V.data(), V.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(gU, 0, U.size() * 4); cudaMemset(gV, 0, V.size() * 4); cudaMemset(fU, 0, nU * 4); cudaMemset(fV, 0, nI * 4);
    cudaMemset(cnt, 0, 64); cudaMemset(err, 0, 4); cudaMemset(ls, 0, 8);
    cudaMemcpy(du, u.data(), B * 8, cudaMemcpyHostToDevice); cudaMemcpy(dp, p.data(), B * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dn, n.data(), B * 8, cudaMemcpyHostToDevice);
    yr_mf_state st = {dU, dV, nullptr, nullptr, nullptr, nullptr, gU, gV, fU, fV, rows, cnt, err, nU, nI, d};
    yr_opt opt = {YR_OPT_SGD, 1, 1e-2, 1e-3, 0.9, 0.999, 1e-8};
    int rc = yr_bpr_mf_train(&st, &opt, du, dp, dn, B, B, ls, nullptr, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (rc || e) { printf("rc %d cuda %d\n", rc, (int)e); return 1; }
    std::vector<float> hg(U.size()); std::vector<int32_t> hf(nU);
    cudaMemcpy(hg.data(), gU, U.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hf.data(), fU, nU * 4, cudaMemcpyDeviceToHost);
    int bad = 0, flags = 0, first = -1, col0 = 0;
    for (int r = 0; r < nU; ++r) {
      bool nz = false;
      for (int k = 0; k < d; ++k) nz |= hg[(size_t)r * d + k] != 0.f;
      if (nz) { ++bad; if (first < 0) first = r; col0 += hg[(size_t)r * d] != 0.f; }
      flags += hf[r];
    }
    bad_total += bad;
    if (rep < 3 || bad) printf("d=%d variant %s rep %d: leftover gU rows %d (first %d, with col0 nonzero %d), flags left %d\n", d, VARIANT_NAME, rep, bad, first, col0, flags);
  }
  printf("d=%d variant %s: TOTAL leftover rows over 20 reps = %d\n", d, VARIANT_NAME, bad_total);
  return 0;
}
