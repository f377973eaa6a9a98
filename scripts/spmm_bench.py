"""Profiling aid: time yr_spmm_csr (fwd and accumulate) on the Yelp-shape Laplacian for every kernel variant
(YR_SPMM_VARIANT) and both plan orders (YR_SPMM_PLAN_SORT). Usage: python scripts/spmm_bench.py [reps] [variants,comma]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from yelprecommendation_b200 import ops
from yelprecommendation_b200.models.ngcf import laplacian_to_csr

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [-1]
sorts = sys.argv[3].split(",") if len(sys.argv) > 3 else ["0"]
w = bench.build_workload()
dev = torch.device("cuda", 0)
n = w.inter.num_users + w.inter.num_items
ref = {}
for srt in sorts:
    os.environ["YR_SPMM_PLAN_SORT"] = srt
    csr = laplacian_to_csr(w.L, dev)
    for var in variants:
        os.environ["YR_SPMM_VARIANT"] = str(var)
        for d in (64, 128):
            torch.manual_seed(0)
            X, Y = torch.randn(n, d, device=dev), torch.zeros(n, d, device=dev)
            out = ops.spmm_csr(csr.fwd, X)
            key = d
            if key not in ref:
                ref[key] = out.clone()
            same = bool(torch.equal(out, ref[key]))
            for acc in (False, True):
                f = (lambda i: ops.spmm_csr(csr.bwd, X, out=Y, accumulate=True)) if acc else (lambda i: ops.spmm_csr(csr.fwd, X, out=Y))
                for _ in range(5):
                    f(0)
                ms = bench.timed(f, reps) / reps
                print(f"sort={srt} variant={var:2d} d={d} acc={int(acc)}: {1e3 * ms:7.1f} us  gather {csr.fwd.nnz * d * 4 / (ms * 1e-3) / 1e12:5.2f} TB/s  "
                      f"bit_identical={same}", flush=True)
