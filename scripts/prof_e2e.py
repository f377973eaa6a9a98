"""Profiling aid: host-side cost of NGCFTrainer.train([batch]) per step (cProfile, 300 steps)."""
import os, sys, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from yelprecommendation_b200.data import synthetic as syn
from yelprecommendation_b200.trainers import NGCFTrainer
w = bench.build_workload()
torch.manual_seed(42)
ntr = NGCFTrainer(bench.cfg(), w.inter.num_items, w.inter.num_users, w.L)
tu, tp_, tn = w.tri
hb = syn.to_batches(tu[: 2048 * 330], tp_[: 2048 * 330], tn[: 2048 * 330], 2048)
for b in hb[:20]:
    ntr.train([b])
torch.cuda.synchronize()
t0 = time.perf_counter()
for b in hb[20:120]:
    ntr.train([b])
torch.cuda.synchronize()
print(f"e2e {1e3 * (time.perf_counter() - t0) / 100:.3f} ms/step")
pr = cProfile.Profile()
pr.enable()
for b in hb[120:320]:
    ntr.train([b])
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
