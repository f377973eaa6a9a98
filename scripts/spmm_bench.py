"""Profiling aid: time yr_spmm_csr (fwd and accumulate) on the Yelp-shape Laplacian. Usage: python scripts/spmm_bench.py [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from yelprecommendation_b200 import ops
from yelprecommendation_b200.models.ngcf import laplacian_to_csr

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
w = bench.build_workload()
dev = torch.device("cuda", 0)
csr = laplacian_to_csr(w.L, dev)
n = w.inter.num_users + w.inter.num_items
for d in (64, 128):
    X, Y = torch.randn(n, d, device=dev), torch.zeros(n, d, device=dev)
    for acc in (False, True):
        f = (lambda i: ops.spmm_csr(csr.bwd, X, out=Y, accumulate=True)) if acc else (lambda i: ops.spmm_csr(csr.fwd, X, out=Y))
        for _ in range(5):
            f(0)
        ms = bench.timed(f, reps) / reps
        print(f"spmm d={d} acc={acc}: {1e3 * ms:.1f} us  gather {csr.fwd.nnz * d * 4 / (ms * 1e-3) / 1e12:.2f} TB/s", flush=True)
