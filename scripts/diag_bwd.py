import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, ctypes as C
from oracle import cport
from util import rel_err
from yelprecommendation_b200 import ops, _cabi
from yelprecommendation_b200.data import synthetic as syn
from yelprecommendation_b200.data.graph import build_laplacian, laplacian_to_csr
lib = _cabi.load()
n_u, n_i = 300, 340
inter = syn.make_interactions(num_users=n_u, num_items=n_i, nnz=6000, seed=12, n_clusters=4)
L = build_laplacian(inter.user, inter.item, inter.rating, n_u, n_i)
csr = laplacian_to_csr(L, "cuda")
rng = np.random.default_rng(0)
n = n_u + n_i
E = rng.standard_normal((n, 64)).astype(np.float32)
W1, W2 = (rng.standard_normal((64, 64)).astype(np.float32) * 0.2 for _ in range(2))
Gn = rng.standard_normal((n, 64)).astype(np.float32)
c = (csr.fwd.rowptr.cpu().numpy(), csr.fwd.col.cpu().numpy(), csr.fwd.val.cpu().numpy())
ct = (csr.bwd.rowptr.cpu().numpy(), csr.bwd.col.cpu().numpy(), csr.bwd.val.cpu().numpy())
En, LE = cport.ngcf_layer_fwd(c, E, W1, W2)
G0 = np.zeros_like(E)
Go, dW1o, dW2o = cport.ngcf_layer_bwd(ct, E, LE, En, Gn, W1, W2, G0)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for mode in (0, 1):
    G = torch.zeros(n, 64, device="cuda")
    dW1, dW2 = ops.ngcf_layer_bwd(csr, cu(E), cu(LE), cu(En), cu(Gn), cu(W1), cu(W2), G, dense_mode=mode)
    torch.cuda.synchronize()
    print("mode", mode, "G", rel_err(G.cpu().numpy(), Go), "dW1", rel_err(dW1.cpu().numpy(), dW1o), "dW2", rel_err(dW2.cpu().numpy(), dW2o))
    if mode == 1:
        d = dW1.cpu().numpy()
        print(" dW1 vs dW1o^T:", rel_err(d, dW1o.T), " ratio sample", d[:2, :4], dW1o[:2, :4])
        # check dense part separately: T = dS + dP*E is inside G via spmm; compare G rows blockwise
        err_rows = np.abs(G.cpu().numpy() - Go).max(axis=1)
        print(" worst rows", np.argsort(-err_rows)[:8], err_rows.max(), "rows>1e-3:", int((err_rows > 1e-3 * np.abs(Go).max()).sum()), "of", n)
