"""Profiling aid: MFTrainer.train(list of host batches) throughput, repeated."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from yelprecommendation_b200.data import synthetic as syn
from yelprecommendation_b200.trainers import MFTrainer
w = bench.build_workload()
tu, tp_, tn = w.tri
torch.manual_seed(42)
mtr = MFTrainer(bench.cfg(optimizer="sgd"), w.inter.num_items, w.inter.num_users)
hb = syn.to_batches(tu[: 2048 * 200], tp_[: 2048 * 200], tn[: 2048 * 200], 2048)
mtr.train(hb[:8])
for rep in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mtr.train(hb)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"rep {rep}: {2048 * 200 / dt / 1e6:.1f} M triples/s  ({1e6 * dt / 200:.1f} us/batch)", flush=True)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); mtr.train(hb); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
