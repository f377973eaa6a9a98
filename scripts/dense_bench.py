"""Micro-benchmark + accuracy check of the NGCF dense transforms (yr_ngcf_dense_fwd / yr_ngcf_dense_bwd) per dense_mode.
Usage: python scripts/dense_bench.py [fwd|bwd|both]   (prints one line per (d, n, mode); float64 torch as the yardstick)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yelprecommendation_b200 import _cabi  # noqa: E402

F32 = torch.float32
SLOPE = 0.2


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "both"
    lib = _cabi.load()
    dev = torch.device("cuda:0")
    p = _cabi.dptr
    st = lambda: _cabi.stream_ptr(dev)
    cases = [(64, 69716), (64, 4099), (128, 69716), (128, 1_500_000)]
    if os.environ.get("YR_DENSE_BIG"):
        cases.append((128, 12_000_000))
    for d, n in cases:
        g = torch.Generator(device=dev).manual_seed(d + n)
        E = torch.randn(n, d, device=dev, generator=g) * 0.1
        LE = torch.randn(n, d, device=dev, generator=g) * 0.1
        W1 = torch.randn(d, d, device=dev, generator=g) * (1.0 / d ** 0.5)
        W2 = torch.randn(d, d, device=dev, generator=g) * (1.0 / d ** 0.5)
        check = n <= 2_000_000
        if what in ("fwd", "both"):
            if check:
                z = ((LE + E).double() @ W1.double().T + (E * LE).double() @ W2.double().T)
                ref = torch.where(z > 0, z, z * SLOPE)
            for mode in (0, 1):
                out = torch.zeros(n, d, device=dev)
                fn = lambda: _cabi.check(lib.yr_ngcf_dense_fwd(d, n, p(E, F32), p(LE, F32), p(W1, F32), p(W2, F32), SLOPE,
                                                               p(out, F32), mode, st()), "fwd")
                ms = timed(fn, 20 if n < 1_000_000 else 5)
                err = float(((out.double() - ref).abs().max() / ref.abs().max())) if check else float("nan")
                gb = 3 * n * d * 4 / 1e9
                print(f"fwd d={d} n={n} mode={mode}: {ms * 1e3:9.1f} us  {gb / ms:7.2f} TB/s(alg)  max_rel_err={err:.2e}", flush=True)
        if what in ("bwd", "both"):
            En = torch.randn(n, d, device=dev, generator=g)
            Gn = torch.randn(n, d, device=dev, generator=g) * 0.01
            G0 = torch.randn(n, d, device=dev, generator=g) * 0.01
            nbytes = lib.yr_ngcf_layer_bwd_ws_bytes(d)
            ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
            if check:
                dZ = torch.where(En > 0, Gn, Gn * SLOPE).double()
                dS, dP = dZ @ W1.double(), dZ @ W2.double()
                T_ref = dS + dP * E.double()
                G_ref = G0.double() + dS + dP * LE.double()
                dW1_ref = dZ.T @ (LE + E).double()
                dW2_ref = dZ.T @ (E * LE).double()
            for mode in (0, 2):
                G = G0.clone()
                T = torch.zeros(n, d, device=dev)
                dW1 = torch.zeros(d, d, device=dev)
                dW2 = torch.zeros(d, d, device=dev)
                fn = lambda: _cabi.check(lib.yr_ngcf_dense_bwd(d, n, p(E, F32), p(LE, F32), p(En, F32), p(Gn, F32), p(W1, F32),
                                                               p(W2, F32), SLOPE, p(G, F32), p(T, F32), p(dW1, F32), p(dW2, F32),
                                                               p(ws), nbytes, mode, st()), "bwd")
                fn()
                torch.cuda.synchronize()
                if check:
                    rel = lambda a, b: float((a.double() - b).abs().max() / b.abs().max())
                    errs = (rel(T, T_ref), rel(G, G_ref), rel(dW1, dW1_ref), rel(dW2, dW2_ref))
                else:
                    errs = (float("nan"),) * 4
                ms = timed(fn, 20 if n < 1_000_000 else 5)
                gb = 7 * n * d * 4 / 1e9
                print(f"bwd d={d} n={n} mode={mode}: {ms * 1e3:9.1f} us  {gb / ms:7.2f} TB/s(alg)  "
                      f"err T/G/dW1/dW2 = {errs[0]:.1e} {errs[1]:.1e} {errs[2]:.1e} {errs[3]:.1e}", flush=True)
        del E, LE
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
