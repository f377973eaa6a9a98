"""Randomised stress of the tensor-core evaluation against the FP32-pipe kernel (bit-identical outputs expected)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from yelprecommendation_b200 import ops
from yelprecommendation_b200.data.graph import build_eval_csr

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
dev = torch.device("cuda")
bad = 0
for case in range(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    d = int(rng.choice([32, 64, 96, 128, 192, 256]))
    K = int(rng.integers(1, 17))
    nU = int(rng.integers(1, 700))
    nI = int(rng.integers(K + 1, 6000))
    n_eval = int(rng.integers(1, 600))
    scale = float(rng.choice([1e-3, 1.0, 30.0]))
    U = (rng.standard_normal((nU, d)) * scale).astype(np.float32)
    V = (rng.standard_normal((nI, d)) * scale).astype(np.float32)
    if rng.random() < 0.3:                                   # many exact ties: few distinct item rows
        V = V[rng.integers(0, max(2, nI // 50), nI)]
    if rng.random() < 0.2:
        U[rng.integers(0, nU)] = 0.0                         # an all-zero user: every score ties at 0
    uid = rng.integers(0, nU, n_eval)
    pos, mask = [], []
    for _ in range(n_eval):
        pos.append(rng.integers(0, nI, int(rng.integers(0, 30))).tolist())
        m = int(rng.choice([0, 5, 50, max(0, nI - K - 3), nI]))
        mask.append(np.unique(rng.integers(0, nI, m)).tolist() if m else [])
    csr = build_eval_csr(uid, pos, mask, nI)
    ecsr = ops.DeviceEvalCSR(csr, dev, K)
    Ud, Vd = torch.from_numpy(U).to(dev), torch.from_numpy(V).to(dev)
    a = ops.eval_topk_metrics(Ud, Vd, ecsr, mode="exact")
    if not ops._cabi.load().yr_eval_tc_supported(d, K):
        continue
    b = ops.eval_topk_metrics(Ud, Vd, ecsr, mode="tc")
    fb = int(ops.eval_topk_metrics.last_fallback_rows.item())
    ok = torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])
    if not ok:
        bad += 1
    print(f"case {case:3d} d={d:3d} K={K:2d} nU={nU:4d} nI={nI:5d} n_eval={n_eval:4d} scale={scale:g} fallback={fb:4d} {'OK' if ok else 'MISMATCH'}", flush=True)
print("mismatches:", bad)
sys.exit(1 if bad else 0)
